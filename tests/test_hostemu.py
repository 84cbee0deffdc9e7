"""The CUDA step / gradient kernels THEMSELVES in the CPU suite.

tests/hostemu runs the product's kernel source (dojo.jl_b200/csrc/dojo_kernels.cuh, dojo_grad.cuh, dojo_linalg.cuh and the
kernel body + plan-table builder cut out of dojo_b200.cu by marker comments) on CPU fibers: one fiber per CUDA thread, named
barriers, warp shuffles, several environments per CTA with the CTA-wide alignment barrier, the atomic work queue and the
completion-ordered hand-over from the forward to the gradient launch.  It is test infrastructure (not a fallback: the product
never includes it) and it is deterministic, so a missing synchronisation shows up as a reproducible difference.

What is checked here, against the oracle and the golden vectors, is therefore the device arithmetic and control flow of the
real kernels; what only a GPU can check (the C-ABI glue, streams, the programmatic dependent launch, performance) stays in
the `-m gpu` tests."""
import os

import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from oracle.oracle import Oracle

from conftest import jittered_states, random_inputs
from hostemu.harness import HostEmu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,B,T,slots", [("pendulum", 5, 40, 4), ("ant", 6, 4, 4), ("quadruped", 4, 4, 2), ("atlas", 2, 3, 2)])
def test_fused_rollout_matches_oracle(name, B, T, slots):
    """dojo_rollout semantics: T steps in one launch, several environments per CTA; every recorded state against the oracle
    (identical status and Newton-iteration counts, |dz| <= 1e-9 on these well-conditioned steps)."""
    mech = dj.get_mechanism(name)
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(17)
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    U = np.stack([random_inputs(mech, B, rng) for _ in range(T)])
    Zf, st, it, sol, traj = em.step(Z, U, T=T, slots=slots, record=True)
    Zo = Z.copy()
    for k in range(T):
        for e in range(B):
            Zo[e], so, io = o.step(Zo[e], U[k, e])
            assert so == 0
        assert np.abs(traj[k] - Zo).max() < 1e-9
    assert np.array_equal(Zf, traj[-1]) and (st == 0).all()


@pytest.mark.parametrize("name", ["ant", "quadruped"])
def test_single_step_status_iterations_solution(name):
    """one step! per launch: status, iteration counts and the full solution vector (joint impulses, velocities, contact
    slacks / duals in the reference ordering) against the oracle; results do not depend on the slot count, on the plan tables
    living in shared or global memory, or on the number of CTAs"""
    mech = dj.get_mechanism(name)
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(19)
    B = 5
    Z = jittered_states(mech, B, rng)
    for _ in range(3):
        Z = np.stack([o.step(Z[e], random_inputs(mech, 1, rng)[0])[0] for e in range(B)])
    U = random_inputs(mech, B, rng)
    Zn, st, it, sol = em.step(Z, U, slots=1)
    for e in range(B):
        zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
        assert (st[e], it[e]) == (so, io)
        assert np.abs(Zn[e] - zo).max() < 1e-9 and np.abs(sol[e] - solo).max() < 1e-7
    for kw in (dict(slots=4), dict(slots=2, smem_plan=False), dict(slots=2, grid=3)):
        Z2, st2, it2, sol2 = em.step(Z, U, **kw)
        assert np.array_equal(Z2, Zn) and np.array_equal(it2, it) and np.array_equal(sol2, sol), kw


@pytest.mark.parametrize("name,slots_grad", [("pendulum", 2), ("ant", 2), ("quadruped", 1)])
def test_gradient_kernel_matches_oracle(name, slots_grad):
    """dojo_step_grad: forward launch + gradient launch consuming the completion-ordered list, against get_maximal_gradients
    of the oracle (tolerances of tests/test_gpu_parity.py::test_gradient_parity)"""
    mech = dj.get_mechanism(name)
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(23)
    B = 4
    Z = jittered_states(mech, B, rng) if mech.Nb > 1 else np.tile(mech.z0, (B, 1))
    for _ in range(3):
        Z = np.stack([o.step(Z[e], random_inputs(mech, 1, rng)[0])[0] for e in range(B)])
    U = random_inputs(mech, B, rng)
    Zn, Fz, Fu, st, it = em.step_grad(Z, U, slots=2, slots_grad=slots_grad)
    Z1, st1, it1, _ = em.step(Z, U)
    assert np.array_equal(Zn, Z1) and np.array_equal(it, it1)
    errs = []
    for e in range(B):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        assert so == st[e] == 0 and io == it[e]
        errs.append(max(np.abs(Fz[e] - Fzo).max() / max(1.0, np.abs(Fzo).max()), np.abs(Fu[e] - Fuo).max() / max(1.0, np.abs(Fuo).max())))
    assert np.median(errs) < 1e-7 and max(errs) < 1e-4, errs
    # without the completion-ordered list (DOJO_B200_NO_GRAD_OVERLAP path) the gradients are bit-identical
    _, Fz2, Fu2, _, _ = em.step_grad(Z, U, slots=2, slots_grad=slots_grad, publish_order=False)
    assert np.array_equal(Fz2, Fz) and np.array_equal(Fu2, Fu)


@pytest.mark.parametrize("name", ["pendulum", "ant", "quadruped", "atlas"])
def test_kernels_reproduce_golden_vectors(name):
    """tests/golden/*.npz through the kernel source on the CPU (the -m gpu twin is tests/test_golden.py)"""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    mech = dj.get_mechanism(name)
    em = HostEmu(mech)
    n = min(g["Z"].shape[0], 4)
    Zn, st, it, _ = em.step(g["Z"][:n], g["U"][:n], slots=2)
    assert np.array_equal(st, g["status"][:n])
    same = it == g["iters"][:n]
    assert same.mean() >= 0.75
    assert np.abs(Zn - g["Z_next"][:n])[same].max(initial=0.0) < 1e-6 and np.abs(Zn - g["Z_next"][:n]).max() < 5e-3


def test_options_flags_and_failure_status():
    """max_iter exhaustion is reported per environment (:failed), the Q1-literal flag advances the configuration twice, external
    forces enter the step -- all through the kernel code"""
    mech = dj.get_mechanism("ant")
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(29)
    Z = jittered_states(mech, 2, rng)
    U = random_inputs(mech, 2, rng)
    tight = capi.solver_options(max_iter=2)
    _, st, it, _ = em.step(Z, U, opts=tight)
    assert list(st) == [1, 1] and list(it) == [2, 2]
    Zn, _, _, _ = em.step(Z, U)
    Zq, _, _, _ = em.step(Z, U, flags=1)
    zn = Zn.reshape(2, mech.Nb, 13)
    zq = Zq.reshape(2, mech.Nb, 13)
    assert np.abs(zq[:, :, 0:3] - (zn[:, :, 0:3] + mech.timestep * zn[:, :, 3:6])).max() < 1e-12
    F = rng.normal(0.0, 1.0, (2, 6 * mech.Nb))
    Zf, _, _, _ = em.step(Z, U, fext=F)
    for e in range(2):
        zo, so, io = o.step(Z[e], U[e], fext=F[e])
        assert np.abs(Zf[e] - zo).max() < 1e-9


def test_contact_gradient_columns_match_oracle():
    """dojo_step_grad_contact: the 5 Ni contact-data columns are solved against the same factor in the gradient kernel;
    against get_contact_gradients of the oracle (ant after a roll-in: feet on the ground); the state / control gradients are
    bit-identical with and without the extra columns"""
    mech = dj.get_mechanism("ant")
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(31)
    B = 4
    Z = jittered_states(mech, B, rng)
    for _ in range(12):
        Z = np.stack([o.step(Z[e], random_inputs(mech, 1, rng)[0])[0] for e in range(B)])
    U = random_inputs(mech, B, rng)
    Zn, Fz, Fu, Fc, st, it = em.step_grad(Z, U, slots=2, slots_grad=2, contact=True)
    _, Fz0, Fu0, _, _ = em.step_grad(Z, U, slots=2, slots_grad=2)
    assert np.array_equal(Fz, Fz0) and np.array_equal(Fu, Fu0) and Fc.shape == (B, 12 * mech.Nb, 5 * mech.Ni)
    errs = []
    for e in range(B):
        _, _, _, so, io = o.step_grad(Z[e], U[e])
        Fco = o.contact_gradients()
        assert so == st[e] == 0 and io == it[e]
        errs.append(np.abs(Fc[e] - Fco).max() / max(1.0, np.abs(Fco).max()))
        assert np.abs(Fco).max() > 1.0  # contacts are active: the columns are not trivially zero
    assert np.median(errs) < 1e-7 and max(errs) < 1e-4, errs


@pytest.mark.parametrize("name", ["ant", "atlas"])
def test_kinjac_kernel_with_real_ctas(name):
    """dojo_kinjac_kernel as launched on the GPU (128 threads per CTA, persistent grid over the environments, workspace slice
    per CTA): map Jacobians against the oracle, minimal gradients against M Fz N / M Fu"""
    from test_oracle_properties import _random_minimal
    mech = dj.get_mechanism(name)
    em, o = HostEmu(mech), Oracle(mech)
    rng = np.random.default_rng(37)
    B = 5
    X = np.stack([_random_minimal(mech, rng) for _ in range(B)])
    Z = np.stack([o.minimal_to_maximal(x) for x in X])
    M = em.kinjac(0, Z, grid=2)
    N = em.kinjac(1, Z, grid=3)
    for e in range(B):
        Mo, No = o.maximal_to_minimal_jacobian(Z[e]), o.minimal_to_maximal_jacobian(Z[e])
        assert np.abs(M[e] - Mo).max() < 1e-10 * max(1.0, np.abs(Mo).max())
        assert np.abs(N[e] - No).max() < 1e-10 * max(1.0, np.abs(No).max())
    Fz = rng.normal(size=(B, 12 * mech.Nb, 12 * mech.Nb))
    Fu = rng.normal(size=(B, 12 * mech.Nb, mech.nu))
    Zn = Z[::-1].copy()
    Gx, Gu = em.kinjac(2, Z, Zn, Fz, Fu, grid=2)
    Mn = em.kinjac(0, Zn)
    ref_x, ref_u = Mn @ Fz @ N, Mn @ Fu
    assert np.abs(Gx - ref_x).max() < 1e-10 * max(1.0, np.abs(ref_x).max())
    assert np.abs(Gu - ref_u).max() < 1e-10 * max(1.0, np.abs(ref_u).max())
