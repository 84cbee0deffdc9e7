import sys, os
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else '.')
import numpy as np, torch
import dojo_jl_b200 as dj
from dojo_jl_b200.solver import BatchedStepper
import bench
name, B = sys.argv[1], int(sys.argv[2]); kw = {"contact_type": sys.argv[3]} if len(sys.argv) > 3 else {}
mech = dj.get_mechanism(name, **kw)
if name in bench.WORKLOADS:
    Z0, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1, name); U = bench.random_inputs(mech, rng, 30, B, bench.SCALE[name])
else:
    rng = np.random.default_rng(1); Z0 = np.tile(mech.z0, (B, 1)); Z0[:, 2] += rng.uniform(0.2, 0.6, B); Z0[:, 3:6] = rng.normal(0, 1.0, (B, 3)); Z0[:, 10:13] = rng.normal(0, 2.0, (B, 3)); U = np.zeros((30, B, mech.nu))
U = torch.from_numpy(U).cuda(); s = BatchedStepper(mech, B)
Za = torch.from_numpy(Z0).cuda(); Zb = torch.empty_like(Za); ng = 12 * mech.Nb
Fz = torch.empty((B, ng, ng), dtype=torch.float64, device="cuda"); Fu = torch.empty((B, mech.nu, ng), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for t in range(8):
    s.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, stream=st); Za, Zb = Zb, Za
torch.cuda.synchronize(); ts = []
for t in range(8, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.step_grad_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), Fz.data_ptr(), Fu.data_ptr(), B, stream=st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1)); Za, Zb = Zb, Za
print(name, kw, B, "grad step times [ms]:", np.round(ts, 2))
