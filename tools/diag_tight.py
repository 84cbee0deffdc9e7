import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200 import solver
if os.environ.get("DOJO_LIB"): solver.LIB_PATH = os.path.abspath(os.environ["DOJO_LIB"])
from dojo_jl_b200.solver import BatchedStepper
from oracle.oracle import Oracle
from conftest import jittered_states, random_inputs
mech = dj.get_mechanism("ant"); rng = np.random.default_rng(11); B=48
opts = capi.solver_options(rtol=1e-9, btol=1e-9)
Z = jittered_states(mech, B, rng); st = BatchedStepper(mech, B); o = Oracle(mech, opts)
for t in range(12):
    U = random_inputs(mech, B, rng, 1.0)
    Zg, sg, ig = st.step(Z, U, opts=opts)
    Zo = np.empty_like(Z); so = np.zeros(B, np.int32); io = np.zeros(B, np.int32)
    for e in range(B):
        Zo[e], so[e], io[e] = o.step(Z[e], U[e])
    err = np.abs(Zg-Zo).max(axis=1)
    bad = np.where((err > 1e-7) | (sg != so))[0]
    print("step", t, "max err", err.max(), "mismatch iters", int((ig!=io).sum()), "status g/o", int((sg!=0).sum()), int((so!=0).sum()), [(int(e), float(err[e]), int(ig[e]), int(io[e]), int(sg[e]), int(so[e])) for e in bad[:4]])
    Z = Zo
