import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200.solver import BatchedStepper
from oracle.oracle import Oracle
from conftest import jittered_states, random_inputs
name, B, T, scale = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
mech = dj.get_mechanism(name); rng = np.random.default_rng(7)
Z = jittered_states(mech, B, rng); st = BatchedStepper(mech, B); o = Oracle(mech)
for t in range(T):
    U = random_inputs(mech, B, rng, scale)
    Zg, sg, ig, solg = st.step(Z, U, return_sol=True)
    Zo = np.empty_like(Z); so = np.zeros(B, np.int32); io = np.zeros(B, np.int32)
    for e in range(B):
        Zo[e], so[e], io[e] = o.step(Z[e], U[e])
        err = np.abs(Zg[e]-Zo[e]).max()
        if err > 1e-7:
            tr = o.trace()
            print("step", t, "env", e, "err %.3e"%err, "status", so[e], sg[e], "iters", io[e], ig[e], "last trace", tr[-1], "n_trace", len(tr))
    Z = Zo
print("done")
