"""rtol = btol = 1e-10: the device kernels converge wherever the CPU oracle does.

Round 1 condensed BOTH sides of every joint limit analytically (division by the slack of the active side, s -> 0): terms of size
gamma / s ~ 1e16 on the body rows swamped the dynamics, the linear residual stalled at ~3e-9 and a quarter of the environments in hard
contact ended :failed below rtol ~1e-8 while the oracle (and the reference, which eliminates the bodies BEFORE the joint node)
converged.  The kernels now keep the dual of the limit side nearer to its bound as an explicit row of the joint's node
(dojo_plan.h joint_nq, dojo_kernels.cuh limit_side).  The kernel source runs here on the CPU emulation (tests/hostemu); the same
check on the GPU is tests/test_gpu_parity.py::test_step_parity_tight_tolerances.  The reference's own conservation tests run at 1e-12
(test/momentum.jl:154-218).
"""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from conftest import jittered_states, random_inputs
from hostemu.harness import HostEmu
from oracle.oracle import Oracle, step_batch_threads


def _states_in_contact(name, B, steps, seed, scale):
    mech = dj.get_mechanism(name)
    rng = np.random.default_rng(seed)
    Z = jittered_states(mech, B, rng)
    for _ in range(steps):  # default tolerances: fall onto the ground, joints driven into their limits
        Z, _, _ = step_batch_threads(mech, Z, random_inputs(mech, B, rng, scale), None, 4)
    return mech, Z, random_inputs(mech, B, rng, scale)


@pytest.mark.parametrize("name,B,steps,scale,tol", [("ant", 24, 14, 1.0, 1e-10), ("quadruped", 12, 10, 2.0, 1e-10)])
def test_tight_tolerances_converge_like_the_oracle(name, B, steps, scale, tol):
    mech, Z, U = _states_in_contact(name, B, steps, seed=29, scale=scale)
    opts = capi.solver_options(rtol=tol, btol=tol)
    Zo, so, io = step_batch_threads(mech, Z, U, opts, 4)
    Zg, sg, ig, _ = HostEmu(mech).step(Z, U, opts=opts, slots=2)
    assert (so == 0).sum() >= B // 2                      # the case is meaningful: the oracle converges (in 15 - 25 iterations)
    assert np.array_equal(sg, so), (sg, so)
    conv = so == 0
    assert np.abs(ig[conv] - io[conv]).max() <= 3         # iteration counts within the noise of the last digits
    assert np.abs(Zg - Zo)[conv].max() < 1e-6
    assert np.median(np.abs(Zg - Zo)[conv].max(axis=1)) < 1e-10
