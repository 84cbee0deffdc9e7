"""Short single-GPU run for ncu: bring a batch to the state bench.py times, then `steps` more steps
   python tools/prof_one.py mech B steps [fwd|grad] [contact_type]      (the kernels to capture are the LAST launches)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dojo_jl_b200 as dj
from dojo_jl_b200.solver import BatchedStepper
import bench
name, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "fwd"
kw = {"contact_type": sys.argv[5]} if len(sys.argv) > 5 else {}
mech = dj.get_mechanism(name, **kw)
w = bench.WORKLOADS.get(name, dict(rollin=8, episode=0))
pre = w["rollin"] + 3 if not w["episode"] else 4
if name in bench.WORKLOADS:
    Z0, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1, name)
    U = bench.random_inputs(mech, rng, pre + steps, B, bench.SCALE[name])
else:  # widened models (block / sphere with the orthant contact models): dropped from 0.2 .. 0.6 m with random spin
    rng = np.random.default_rng(1)
    Z0 = np.tile(mech.z0, (B, 1)); Z0[:, 2] += rng.uniform(0.2, 0.6, B); Z0[:, 3:6] = rng.normal(0, 1.0, (B, 3)); Z0[:, 10:13] = rng.normal(0, 2.0, (B, 3))
    U = np.zeros((pre + steps, B, mech.nu))
U = torch.from_numpy(U).cuda()
s = BatchedStepper(mech, B)
Za = torch.from_numpy(Z0).cuda(); Zb = torch.empty_like(Za)
st = torch.cuda.current_stream().cuda_stream
ng = 12 * mech.Nb
if mode == "grad":
    Fz = torch.empty((B, ng, ng), dtype=torch.float64, device="cuda"); Fu = torch.empty((B, mech.nu, ng), dtype=torch.float64, device="cuda")
it = torch.zeros(B, dtype=torch.int32, device="cuda"); stt = torch.zeros(B, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(pre + steps):
    if t == pre: torch.cuda.synchronize(); e0.record()
    if mode == "grad" and t >= pre:
        s.step_grad_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), Fz.data_ptr(), Fu.data_ptr(), B, dstatus=stt.data_ptr(), diters=it.data_ptr(), stream=st)
    else:
        s.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, dstatus=stt.data_ptr(), diters=it.data_ptr(), stream=st)
    Za, Zb = Zb, Za
e1.record(); torch.cuda.synchronize()
print("   config:", s.launch_config)
print(f"{name} {kw} B={B} {mode}: {e0.elapsed_time(e1)/steps:.3f} ms/step  {B*steps/e0.elapsed_time(e1)*1e3:.0f} env-steps/s  mean iters {float(it.float().mean()):.2f}  failed {int((stt!=0).sum())}  pre-steps {pre}  smem/env {s.shared_bytes_per_env}")
