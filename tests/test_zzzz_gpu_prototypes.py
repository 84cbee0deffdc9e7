"""All 16 joint prototypes (springs and dampers on both halves) and limits on 1 - 3 axes of either half THROUGH THE C-ABI ON THE GPU:
the sweeps of tests/test_joint_prototypes.py (reference: src/joints/prototypes.jl:482-499, test/damper.jl:1-17, joints/limits.jl), which
round 1 ran on the kernel emulation only.  Step (status, Newton-iteration count, next state, full solution vector in the reference
ordering) and IFT gradients against the oracle."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from test_joint_prototypes import LIMITED, PROTOTYPES, snake

pytestmark = pytest.mark.gpu


def _compare(m, Z, U, steps, tol_z=1e-9, tol_sol=1e-7):
    from dojo_jl_b200.solver import BatchedStepper
    o, s = Oracle(m), BatchedStepper(m, Z.shape[0])
    B = Z.shape[0]
    for _ in range(steps):
        Zn, st, it, sol = s.step(Z, U, return_sol=True)
        for e in range(B):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            assert (st[e], it[e]) == (so, io)
            assert np.abs(Zn[e] - zo).max() < tol_z and np.abs(sol[e] - solo).max() < tol_sol
        Z = Zn
    Zn, Fz, Fu, st, it = s.step_grad(Z, U)
    for e in range(B):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        assert (st[e], it[e]) == (so, io)
        assert np.abs(Fz[e] - Fzo).max() < 1e-7 * max(1.0, np.abs(Fzo).max()) and np.abs(Fu[e] - Fuo).max() < 1e-7 * max(1.0, np.abs(Fuo).max())
    if hasattr(s, "close"):
        s.close()
    return sol


@pytest.mark.parametrize("joint_type", list(PROTOTYPES))
def test_prototype_with_springs_and_dampers(joint_type):
    m = snake(joint_type)
    rng = np.random.default_rng(14)
    B = 3
    _compare(m, np.tile(m.z0, (B, 1)), 0.3 * rng.normal(size=(B, m.nu)), steps=12)


@pytest.mark.parametrize("joint_type,half", LIMITED)
def test_limits_on_every_number_of_free_axes(joint_type, half):
    m = snake(joint_type, spring=0.0, damper=0.1)
    j = m.joints[1]
    el = j.rot if half == "rot" else j.tra
    c0 = 0.1 * np.arange(1, j.input_dimension + 1)
    c0 = c0[j.tra.nfree:] if half == "rot" else c0[:j.tra.nfree]
    el.limit_lo, el.limit_hi = c0 - 0.03, c0 + 0.04  # tight box around the initial coordinates: the inputs push into it
    rng = np.random.default_rng(15)
    B = 3
    U = np.zeros((B, m.nu))
    U[:, 6:] = 3.0 * rng.normal(size=(B, m.nu - 6))
    sol = _compare(m, np.tile(m.z0, (B, 1)), U, steps=25)
    s0, nb = m.joint_sol_offset(1) + (0 if half == "tra" else j.tra.nimpulses), 2 * el.nfree  # [s (nb) | gamma (nb) | eq] inside the limited half
    assert sol[:, s0 + nb: s0 + 2 * nb].max() > 0.01  # a limit is active at the end (duals of the last step)
