"""GPU tests of trajectory recording and the momentum / energy diagnostics through the C-ABI:
dojo_step_record / dojo_simulate_record against the oracle's restatement of save_to_storage! and mechanics/{momentum,energy}.jl."""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import api
from conftest import jittered_states, random_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pendulum", "ant", "quadruped", "atlas"])
def test_step_record_matches_oracle(name):
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    mech = dj.get_mechanism(name)
    rng = np.random.default_rng(81)
    B = 12
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    stepper, o = BatchedStepper(mech, B), Oracle(mech)
    for _ in range(8):
        Z, _, _ = stepper.step(Z, random_inputs(mech, B, rng))
    U = random_inputs(mech, B, rng)
    Zn, sto, diag, st, it = stepper.step_record(Z, U)
    Zf, sf, itf = stepper.step(Z, U)
    assert np.array_equal(Zn, Zf) and np.array_equal(st, sf) and np.array_equal(it, itf)  # recording does not change the step
    compared = 0
    for e in range(B):
        zo, so, io = o.step(Z[e], U[e])
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        body, d = o.storage_record()
        # joint impulses of ill-conditioned contact steps agree to ~1e-6 between the two factorisation orders (see
        # test_gradient_parity); the formulas themselves are checked to 1e-11 on the CPU (tests/test_storage_host.py)
        assert np.abs(sto[e] - body).max() < 1e-5 * max(1.0, np.abs(body).max())
        assert np.abs(diag[e] - d).max() < 1e-5 * max(1.0, np.abs(d).max())
        compared += 1
    assert compared >= B // 2


def test_simulate_record_storage_semantics():
    """Storage of simulate!(...; record=true): x/q/v/w of step k = the state before the k-th solve; the momenta and energies
    equal those of step-by-step dojo_step_record calls (bit-identical); device-pointer and host-pointer paths agree."""
    import torch
    from dojo_jl_b200.solver import BatchedStepper
    mech = dj.get_mechanism("ant")
    rng = np.random.default_rng(83)
    B, T = 16, 6
    Z0 = jittered_states(mech, B, rng)
    U = np.stack([random_inputs(mech, B, rng) for _ in range(T)])
    stepper = BatchedStepper(mech, B)
    Zf, traj, sto, diag, st_any = stepper.simulate_record(Z0, U, T)
    Z = Z0
    for k in range(T):
        assert np.array_equal(traj[k], Z)
        Z, s_k, d_k, _, _ = stepper.step_record(Z, U[k])
        assert np.array_equal(sto[k], s_k) and np.array_equal(diag[k], d_k)
    assert np.array_equal(Zf, Z)
    # device pointers
    dZ0, dU = torch.from_numpy(Z0).cuda(), torch.from_numpy(U).cuda()
    dZf = torch.empty_like(dZ0)
    dtraj = torch.empty((T, B, mech.nz), dtype=torch.float64, device="cuda")
    dsto = torch.empty((T, B, mech.Nb, 12), dtype=torch.float64, device="cuda")
    ddiag = torch.empty((T, B, 8), dtype=torch.float64, device="cuda")
    dst = torch.empty(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    import ctypes as C
    from dojo_jl_b200 import capi
    o = capi.solver_options()
    rc = stepper.L.dojo_simulate_record(stepper.h, C.byref(o), B, T, C.c_void_p(dZ0.data_ptr()), C.c_void_p(dU.data_ptr()), C.c_void_p(dZf.data_ptr()),
                                        C.c_void_p(dtraj.data_ptr()), C.c_void_p(dsto.data_ptr()), C.c_void_p(ddiag.data_ptr()), C.c_void_p(dst.data_ptr()))
    assert rc == 0
    assert np.array_equal(dZf.cpu().numpy(), Zf) and np.array_equal(dtraj.cpu().numpy(), traj)
    assert np.array_equal(dsto.cpu().numpy(), sto) and np.array_equal(ddiag.cpu().numpy(), diag) and np.array_equal(dst.cpu().numpy(), st_any)


def test_momentum_conservation_on_device():
    """test/momentum.jl:154-219 on the device: without gravity, contacts out of reach and no inputs the total linear and
    angular momentum recorded by dojo_simulate_record stay constant (internal joint / spring / damper impulses cancel)."""
    mech = dj.get_mechanism("quadruped", gravity=0.0)
    rng = np.random.default_rng(85)
    B, T = 8, 40
    Z0 = jittered_states(mech, B, rng, base_z=(1.0, 2.0))
    Zr = Z0.reshape(B, mech.Nb, 13)
    Zr[:, 0, 3:6] = rng.normal(0, 0.2, (B, 3))      # kick the trunk
    Zr[:, 0, 10:13] = rng.normal(0, 0.5, (B, 3))
    storage = api.simulate_record(mech, T, Z0.reshape(B, -1), opts=api.SolverOptions(rtol=1e-10, btol=1e-10))
    p = api.momentum(mech, storage)
    assert len(storage) == T and p.shape == (T, B, 6)
    drift = np.abs(p[2:] - p[2]).max()   # the first steps absorb the initial joint-constraint violation of the kicked state
    assert drift < 1e-6, drift
    ke = api.kinetic_energy(mech, storage)
    assert (ke >= 0).all() and np.isfinite(api.mechanical_energy(mech, storage)).all()
