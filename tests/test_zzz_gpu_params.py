"""dojo_update_params (system-identification loop: same topology, new numbers) on the GPU: a handle whose parameters were
swapped in place gives bit-identical results to a freshly created handle; a different topology is refused."""
import copy

import numpy as np
import pytest

import dojo_jl_b200 as dj
from conftest import jittered_states, random_inputs

pytestmark = pytest.mark.gpu


def test_update_params_equals_fresh_handle():
    from dojo_jl_b200.solver import BatchedStepper
    mech = dj.get_mechanism("ant")
    rng = np.random.default_rng(91)
    B = 32
    Z = jittered_states(mech, B, rng)
    U = random_inputs(mech, B, rng)
    stepper = BatchedStepper(mech, B)
    for _ in range(6):
        Z, _, _ = stepper.step(Z, U)
    Z0, s0, _ = stepper.step(Z, U)
    mech2 = copy.deepcopy(mech)
    for c in mech2.contacts:
        c.friction = 0.9
        c.radius = c.radius * 1.1
    for b in mech2.bodies:
        b.mass = b.mass * 1.3
        b.inertia = np.asarray(b.inertia) * 1.3
    mech2.gravity = np.array([0.0, 0.0, -5.0])
    stepper.update_params(mech2)
    Z1, s1, i1 = stepper.step(Z, U)
    Zg, Fz1, Fu1, _, _ = stepper.step_grad(Z, U)
    fresh = BatchedStepper(mech2, B)
    Z2, s2, i2 = fresh.step(Z, U)
    _, Fz2, Fu2, _, _ = fresh.step_grad(Z, U)
    assert np.array_equal(Z1, Z2) and np.array_equal(s1, s2) and np.array_equal(i1, i2)
    assert np.array_equal(Fz1, Fz2) and np.array_equal(Fu1, Fu2)
    assert np.abs(Z1 - Z0).max() > 1e-6  # the parameters did change the dynamics
    stepper.update_params(mech)           # and back
    Z3, _, _ = stepper.step(Z, U)
    assert np.array_equal(Z3, Z0)


def test_update_params_refuses_other_topology():
    from dojo_jl_b200.solver import BatchedStepper
    stepper = BatchedStepper(dj.get_mechanism("ant"), 4)
    with pytest.raises(RuntimeError, match="topology"):
        stepper.update_params(dj.get_mechanism("quadruped"))


def test_contact_gradients_parity_and_sysid_finite_differences():
    """dojo_step_grad_contact vs the oracle's get_contact_gradients, and -- the system-identification loop itself -- vs central
    differences of dojo_step through dojo_update_params(contact radius +- eps) on a standing quadruped"""
    import copy
    from dojo_jl_b200 import capi
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    mech = dj.get_mechanism("quadruped")
    opts = capi.solver_options(rtol=1e-9, btol=1e-9)
    B = 8
    rng = np.random.default_rng(93)
    Z = np.tile(mech.z0, (B, 1))
    Z[:, 2] += rng.uniform(0.0, 0.02, B)
    stepper = BatchedStepper(mech, B)
    U = np.zeros((B, mech.nu))
    for _ in range(40):
        Z, _, _ = stepper.step(Z, U)
    # parity at the benchmark tolerances (1e-6 / 1e-6): at 1e-9 the contact blocks are so close to complementarity that the
    # KKT matrix has a condition number ~1e12 and two exact-arithmetic-equivalent solves (dense LU in the oracle, condensed
    # block LDU on the device) differ by ~1e-4 relative in *all* gradients (reproduced by tests/hostemu)
    popts = capi.solver_options(rtol=1e-6, btol=1e-6)
    Zn, Fz, Fu, Fc, st, it = stepper.step_grad_contact(Z, U, popts)
    Zn2, Fz2, Fu2, _, _ = stepper.step_grad(Z, U, popts)
    assert np.array_equal(Zn, Zn2) and np.array_equal(Fz, Fz2) and np.array_equal(Fu, Fu2)
    o = Oracle(mech, popts)
    errs = []
    for e in range(B):
        _, _, _, so, io = o.step_grad(Z[e], U[e])
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        Fco = o.contact_gradients()
        errs.append(np.abs(Fc[e] - Fco).max() / max(1.0, np.abs(Fco).max()))
    errs = np.array(errs)
    assert len(errs) >= B // 2 and np.median(errs) < 1e-6 and errs.max() < 1e-2, errs
    # tight solves for the finite differences below
    Zn, Fz, Fu, Fc, st, it = stepper.step_grad_contact(Z, U, opts)
    # finite differences through dojo_update_params: radius of contact 0
    eps = 1e-6
    out = []
    for sgn in (1.0, -1.0):
        m2 = copy.deepcopy(mech)
        m2.contacts[0].radius += sgn * eps
        stepper.update_params(m2)
        out.append(stepper.step(Z, U, opts)[0])
    stepper.update_params(mech)
    fd = (out[0] - out[1]) / (2 * eps)
    zr = fd.reshape(B, mech.Nb, 13)
    col = Fc[:, :, 1].reshape(B, mech.Nb, 12)
    # velocities and positions (the attitude rows need the quaternion reduction: compared through x, v, w only)
    scale = max(1.0, np.abs(col).max())
    assert np.abs(zr[:, :, 0:3] - col[:, :, 0:3]).max() < 1e-3 * scale
    assert np.abs(zr[:, :, 3:6] - col[:, :, 3:6]).max() < 1e-3 * scale
    assert np.abs(zr[:, :, 10:13] - col[:, :, 9:12]).max() < 1e-3 * scale
