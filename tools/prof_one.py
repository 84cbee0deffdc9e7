"""Short single-GPU run for ncu: roll a batch in, then a few timed steps (python tools/prof_one.py mech B steps mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200.solver import BatchedStepper
import bench
name, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "fwd"
mech = dj.get_mechanism(name)
Z0, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1)
U = torch.from_numpy(bench.random_inputs(mech, rng, 20 + steps, B, bench.SCALE[name])).cuda()
s = BatchedStepper(mech, B)
Za = torch.from_numpy(Z0).cuda(); Zb = torch.empty_like(Za)
st = torch.cuda.current_stream().cuda_stream
ng = 12 * mech.Nb
if mode == "grad":
    Fz = torch.empty((B, ng, ng), dtype=torch.float64, device="cuda"); Fu = torch.empty((B, mech.nu, ng), dtype=torch.float64, device="cuda")
for t in range(20 + steps):
    if mode == "grad" and t >= 20:
        s.step_grad_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), Fz.data_ptr(), Fu.data_ptr(), B, stream=st)
    else:
        s.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, stream=st)
    Za, Zb = Zb, Za
torch.cuda.synchronize()
print("done", float(Za.abs().max()))
