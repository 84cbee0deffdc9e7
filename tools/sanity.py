"""Small invocation of every C-ABI entry point (for compute-sanitizer): python tools/sanity.py [mech]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dojo_jl_b200 as dj
from dojo_jl_b200.solver import BatchedStepper
from conftest import jittered_states, random_inputs
name = sys.argv[1] if len(sys.argv) > 1 else "ant"
# mech may carry builder options after a colon: "block:linear", "sphere:impact" (contact_type), "cartpole:full" (springs, dampers,
# joint limits: translational terms), "raiberthopper" -- these run on the second compilation of the kernels (dojo_b200_cm.cu)
name, _, opt = name.partition(":")
kw = {}
if opt in ("linear", "impact", "nonlinear"):
    kw["contact_type"] = opt
elif opt == "full" and name == "cartpole":
    kw = dict(springs=2.0, dampers=0.3, joint_limits={"cart_joint": (-0.3, 0.25), "pole_joint": (-1.2, 1.4)})
mech = dj.get_mechanism(name, **kw)
rng = np.random.default_rng(0)
B, T = 11, 3
Z = jittered_states(mech, B, rng) if mech.Nb > 2 else np.tile(mech.z0, (B, 1))
U = random_inputs(mech, B, rng)
s = BatchedStepper(mech, 16)
Zn, st, it = s.step(Z, U)
Zg, Fz, Fu, sg, ig = s.step_grad(Z, U)
Zf, sr, traj = s.rollout(Z, np.stack([U] * T), T, record=True)
X = s.maximal_to_minimal(Zn)
Z2 = s.minimal_to_maximal(X)
Xn, _, _ = s.step_minimal(X, U)
M = s.maximal_to_minimal_jacobian(Zn)
N = s.minimal_to_maximal_jacobian(Z2)
Xg, Gx, Gu, _, _ = s.minimal_gradients(X, U)
Zr, sto, diag, _, _ = s.step_record(Z, U)
Zs, trj, sts, dgs, _ = s.simulate_record(Z, np.stack([U] * T), T)
from dojo_jl_b200 import capi, environments as E
cls = {"ant": E.AntARS, "quadruped": E.QuadrupedSampling, "pendulum": E.Pendulum}.get(name)
if cls is not None:
    spec = capi.env_spec(**cls.spec_kwargs)
    ns, na = s.env_sizes(spec)
    S = np.zeros((B, ns))
    S[:, :2 * mech.nu] = X
    Sn, rew, done, _, _ = s.env_step(spec, S, rng.uniform(-1, 1, (B, na)))
    s.env_reset(spec, Sn, S[0], done)
print("widened ok", float(np.abs(M).max()), float(np.abs(N).max()), float(np.abs(Gx).max()), float(np.abs(sto).max()), float(np.abs(dgs).max()))
print("sanity ok", name, float(np.abs(Zn).max()), float(np.abs(Fz).max()), float(np.abs(traj).max()), float(np.abs(Xn).max()), it.tolist())
