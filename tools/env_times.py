"""Per-environment time distribution of one step (DJ_PROFILE build): python tools/env_times.py <lib.so> mech B"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import dojo_jl_b200 as dj
from dojo_jl_b200 import solver
solver.LIB_PATH = os.path.abspath(sys.argv[1])
from dojo_jl_b200.solver import BatchedStepper
import bench
name, B = sys.argv[2], int(sys.argv[3])
mech = dj.get_mechanism(name)
Z0, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1)
U = torch.from_numpy(bench.random_inputs(mech, rng, 30, B, bench.SCALE[name])).cuda()
s = BatchedStepper(mech, B)
Za = torch.from_numpy(Z0).cuda(); Zb = torch.empty_like(Za)
it = torch.zeros(B, dtype=torch.int32, device="cuda"); stt = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
out = (C.c_ulonglong * 32)()
s.L.dojo_debug_cycles.argtypes = [C.c_void_p, C.c_void_p]
s.L.dojo_debug_env_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for t in range(30):
    torch.cuda.synchronize(); s.L.dojo_debug_cycles(s.h, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, diters=it.data_ptr(), dstatus=stt.data_ptr(), stream=st); Za, Zb = Zb, Za
    e1.record(); torch.cuda.synchronize()
    if t < 26: continue
    buf = np.zeros(2 * B, dtype=np.uint64)
    s.L.dojo_debug_env_times(s.h, buf.ctypes.data_as(C.c_void_p), B)
    start, dur = buf[0::2].astype(np.float64) * 1e-6, buf[1::2].astype(np.float64) * 1e-6
    iters, stat = it.cpu().numpy(), stt.cpu().numpy()
    grid = min(B, 148 * min(max(1, (227 * 1024) // s.shared_bytes_per_env), 4))  # environments in flight
    print(f"step {t}: wall {e0.elapsed_time(e1):.2f} ms  sum(dur)/grid({grid}) {dur.sum()/grid:.2f} ms  max(start+dur) {np.max(start+dur):.2f}  max dur {dur.max():.2f}  failed {int((stat!=0).sum())}")
    q = np.quantile(dur, [0.5, 0.9, 0.99, 0.999]); print("   dur quantiles 50/90/99/99.9 % (ms):", np.round(q, 2), " iters quantiles:", np.quantile(iters, [0.5, 0.9, 0.99, 0.999]))
    top = np.argsort(-dur)[:8]
    print("   slowest: " + "  ".join(f"[e{e} it{iters[e]} st{stat[e]} start{start[e]:.1f} dur{dur[e]:.1f}]" for e in top))
    print("   ms per iteration (median env):", np.median(dur / np.maximum(iters, 1)))
