"""Dump per-environment Newton iteration counts over a rollout: python tools/dump_iters.py mech B T out.npy"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dojo_jl_b200 as dj
from dojo_jl_b200.solver import BatchedStepper
import bench
name, B, T, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
mech = dj.get_mechanism(name)
Z0, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1)
U = torch.from_numpy(bench.random_inputs(mech, rng, T, B, bench.SCALE[name])).cuda()
s = BatchedStepper(mech, B)
Za = torch.from_numpy(Z0).cuda(); Zb = torch.empty_like(Za)
it = torch.zeros(B, dtype=torch.int32, device="cuda"); stt = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
its, sts, zs, Zs, Us = [], [], [], {}, {}
for t in range(T):
    if t in (30, 45): Zs[t] = Za.cpu().numpy().copy(); Us[t] = U[t].cpu().numpy().copy()
    s.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, diters=it.data_ptr(), dstatus=stt.data_ptr(), stream=st); Za, Zb = Zb, Za
    torch.cuda.synchronize()
    its.append(it.cpu().numpy().copy()); sts.append(stt.cpu().numpy().copy())
    zs.append(Za[:, :3].cpu().numpy().copy())
np.savez(out, iters=np.array(its), status=np.array(sts), torso=np.array(zs), Z30=Zs[30], U30=Us[30], Z45=Zs[45], U45=Us[45])
print("saved", out)
