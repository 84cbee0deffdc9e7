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
