"""Generate tests/golden/*.npz with the CPU oracle (oracle/, the C++ restatement of the reference algorithm).

These are ORACLE-generated regression vectors, not reference-generated ones: the reference is Julia and cannot run in the
build image (tools/export_fixtures.jl is the script a Julia-equipped maintainer runs to dump the corresponding results of the real
Dojo.jl).  They pin (a) the oracle against compiler / refactoring drift and (b) the CUDA path
against a fixed set of numbers that does not depend on the oracle being built on the GPU box.

    python tools/make_golden.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dojo_jl_b200 as dj
from oracle.oracle import Oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import jittered_states, random_inputs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"pendulum": (4, 6, 1.0), "ant": (6, 8, 1.0), "quadruped": (4, 8, 1.0), "atlas": (3, 4, 2.0)}  # B, roll-in steps, input scale

for name, (B, T, scale) in CASES.items():
    mech = dj.get_mechanism(name)
    rng = np.random.default_rng(2024)
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    o = Oracle(mech)
    for _ in range(T):  # roll in so that contacts / limits are active
        U = random_inputs(mech, B, rng, scale)
        Z = np.stack([o.step(Z[e], U[e])[0] for e in range(B)])
    U = random_inputs(mech, B, rng, scale)
    Zn, st, it, Fz, Fu = [], [], [], [], []
    ngrad = 0 if name == "atlas" else 1  # gradients of the first environment only (file size)
    for e in range(B):
        if e < ngrad:
            zo, fz, fu, s, i = o.step_grad(Z[e], U[e])
            Fz.append(fz); Fu.append(fu)
        else:
            zo, s, i = o.step(Z[e], U[e])
        Zn.append(zo); st.append(s); it.append(i)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), Z=Z, U=U, Z_next=np.array(Zn), status=np.array(st, np.int32), iters=np.array(it, np.int32),
                        Fz=np.array(Fz), Fu=np.array(Fu), source=np.array("oracle (C++ restatement), tools/make_golden.py"))
    print(name, "B", B, "iters", it, "status", st)
