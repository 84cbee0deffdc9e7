"""Diagnostic (GPU): how far are the fused environment kernels (dojo_env_step) from the composition of the library's own calls
(state_map, input_map, dojo_step_minimal)?  Prints the largest differences of the maps and of the full step."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dojo_jl_b200 as dj  # noqa: E402,F401
from dojo_jl_b200 import environments as E  # noqa: E402
from test_gpu_parity import _random_minimal_batch  # noqa: E402

for name in ("ant_ars", "quadruped_sampling", "pendulum"):
    rng = np.random.default_rng(61)
    B = 24
    env = E.get_environment(name, batch=B)
    mech, spec = env.mechanism, env.spec
    S = np.zeros((B, env.ns))
    S[:, :2 * mech.nu] = _random_minimal_batch(mech, B, rng)
    if mech.Nb > 1:
        S[:, 2] += rng.uniform(0.3, 0.7, B)
    A = rng.uniform(-1, 1, (B, env.na))
    Sn, reward, done, status, iters = env.stepper.env_step(spec, S, A)
    X, U = env.state_map(S), env.input_map(A)
    Xn, st2, it2 = env.stepper.step_minimal(X, U)
    Z = env.stepper.minimal_to_maximal(X)
    Zn, st3, it3 = env.stepper.step(Z, U)
    Xn3 = env.stepper.maximal_to_minimal(Zn)
    d = np.abs(Sn[:, :2 * mech.nu] - Xn).max(axis=1)
    print(name, "env_step vs step_minimal: max", d.max(), "nonzero envs", int((d > 0).sum()), "iters equal", bool(np.array_equal(iters, it2)),
          "| step_minimal vs (min_to_max, step, max_to_min):", np.abs(Xn - Xn3).max(), bool(np.array_equal(it2, it3)))
