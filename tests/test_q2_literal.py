"""DOJO_FLAG_Q2_LITERAL_GRADIENTS: what get_maximal_gradients!(mechanism, z, u) literally returns.

Reference: gradients/state.jl:69-76 runs step! (which ends with update_state!, bodies/set.jl:22-36) and THEN builds the data
Jacobian and the integrator chain rule at the shifted state, while full_matrix(mechanism.system) still holds the KKT entries of
the unshifted final iterate (solver/mehrotra.jl:66-69).  The oracle restates that sequence literally; the device kernels
(run here through the CPU emulation of the kernel source, on the GPU in tests/test_gpu_parity.py) shift the body states in the
arena between the assembly and the gradient pass.
"""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from conftest import jittered_states, random_inputs
from hostemu.harness import HostEmu
from oracle.oracle import Oracle

Q2 = 2  # DOJO_FLAG_Q2_LITERAL_GRADIENTS


def _setup(name, B, steps, seed=5):
    mech = dj.get_mechanism(name)
    rng = np.random.default_rng(seed)
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    U = random_inputs(mech, B, rng, 1.0)
    o = Oracle(mech)
    for _ in range(steps):  # into contact
        Z = np.stack([o.step(Z[e], U[e])[0] for e in range(B)])
    return mech, o, Z, U


def test_oracle_literal_is_the_shifted_data_jacobian_against_the_unshifted_matrix():
    """The literal variant = dense solve of (KKT matrix at the final iterate, unshifted) against (data Jacobian at the state after
    update_state!), chained with the integrator Jacobians of the shifted state: rebuilt here from the oracle's pieces."""
    mech, o, Z, U = _setup("pendulum", 1, 3)
    zn, Fz_lit, Fu_lit, st, _ = o.step_grad(Z[0], U[0], flags=Q2)
    zn2, Fz_con, Fu_con, _, _ = o.step_grad(Z[0], U[0], flags=0)
    assert st == 0 and np.array_equal(zn, zn2)          # the step itself is not affected by the flag
    assert np.abs(Fz_lit - Fz_con).max() > 1e-4         # ... the gradients are
    # pieces: matrix at the unshifted final iterate, data Jacobian at the shifted state
    o.step(Z[0], U[0])
    A, _ = o.assemble(0.0)
    sol = o.get_solution()
    o.set_state(zn, np.zeros(mech.nu))                  # (x2, q2, v15, w15) <- (x3, q3, v25, w25), inputs cleared
    o.set_solution(sol, 0.0)
    D = o.data_jacobian()
    X = np.linalg.solve(A, D)
    off = mech.node_offsets()[mech.Ne]
    dv = X[off:off + 3]                                 # d v25 / d theta of the single body
    np.testing.assert_allclose(Fz_lit[3:6, :], dv[:, :12], rtol=0, atol=1e-9)
    np.testing.assert_allclose(Fu_lit[3:6, :], dv[:, 12:], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name,steps", [("pendulum", 3), ("ant", 6), ("quadruped", 6)])
def test_device_kernels_match_the_literal_oracle(name, steps):
    mech, o, Z, U = _setup(name, 3, steps)
    emu = HostEmu(mech)
    Zn, Fz, Fu, st, it = emu.step_grad(Z, U, flags=Q2)
    Zc, Fzc, Fuc, _, _ = emu.step_grad(Z, U, flags=0)
    assert np.array_equal(Zn, Zc)
    for e in range(Z.shape[0]):
        _, Fzo, Fuo, sto, _ = o.step_grad(Z[e], U[e], flags=Q2)
        assert sto == st[e]
        sz, su = max(1.0, np.abs(Fzo).max()), max(1.0, np.abs(Fuo).max())
        assert np.abs(Fz[e] - Fzo).max() / sz < 1e-7 and np.abs(Fu[e] - Fuo).max() / su < 1e-7
        assert np.abs(Fz[e] - Fzc[e]).max() / sz > 1e-4  # literal != consistent
