import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200.solver import BatchedStepper
from oracle.oracle import Oracle
from conftest import jittered_states, random_inputs
for name, T, scale in (("pendulum", 3, 1.0), ("ant", 25, 1.0), ("quadruped", 30, 1.0), ("atlas", 25, 2.0)):
    mech = dj.get_mechanism(name); rng = np.random.default_rng(3)
    B = 8
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    st = BatchedStepper(mech, B); o = Oracle(mech)
    for t in range(T):
        U = random_inputs(mech, B, rng, scale)
        Z, _, _ = st.step(Z, U)
    U = random_inputs(mech, B, rng, scale)
    Zn, Fz, Fu, sg, ig = st.step_grad(Z, U)
    for e in range(min(B, 4)):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e], use_factor=False)
        zo2, Fzo2, Fuo2, _, _ = o.step_grad(Z[e], U[e], use_factor=True)
        ez = np.abs(Fz[e] - Fzo).max() / max(1, np.abs(Fzo).max()); eu = np.abs(Fu[e] - Fuo).max() / max(1, np.abs(Fuo).max())
        ezf = np.abs(Fz[e] - Fzo2).max() / max(1, np.abs(Fzo).max())
        print(name, "env", e, "rel err Fz %.2e (vs oracle-factor %.2e) Fu %.2e" % (ez, ezf, eu), "oracle dense-vs-factor %.2e" % (np.abs(Fzo - Fzo2).max() / max(1, np.abs(Fzo).max())), "z err %.1e" % np.abs(Zn[e] - zo).max(), "iters", io, ig[e], "scale", np.abs(Fzo).max())
