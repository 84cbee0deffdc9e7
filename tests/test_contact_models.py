"""ImpactContact / LinearContact (reference src/contacts/impact.jl, linear.jl) -- CPU suite.

Oracle pinned by the reference's own property tests for the three contact models:
  test/jacobian.jl:88-93, :109-114  sphere with contact_type = :nonlinear / :linear / :impact: full_matrix == -d(rhs)/d(solution)
  test/data.jl:28-39                jacobian_data! == finite differences with contact_type = :nonlinear / :linear / :impact
  (ours)                            IFT gradients == finite differences of the step; physical behaviour of the three models
Device code (dojo_contact_orthant.cuh inside the DJ_ANY_CONTACT compilation of the kernels) checked against the oracle through
tests/hostemu (the kernel source on CPU fibers); the GPU run of the same comparison is tests/test_zzzz_gpu_contact_models.py.
"""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from oracle.oracle import Oracle

from test_oracle_properties import _perturb_state, _reduce

MODELS = ["nonlinear", "linear", "impact"]


def _thrown(mech, B, rng, steps, stepper):
    """B bodies thrown at the ground with random spin, advanced `steps` steps (in contact afterwards)."""
    Z = np.tile(mech.z0, (B, 1))
    Z[:, 2] += rng.uniform(-0.4 if mech.name == "sphere" else -0.9, 0.0, B)
    Z[:, 3:6] = rng.normal(size=(B, 3)) * [1.0, 1.0, 0.3]
    Z[:, 10:13] = rng.normal(size=(B, 3))
    U = 0.1 * rng.normal(size=(B, mech.nu))
    for _ in range(steps):
        Z = stepper(Z, U)
    return Z, U


def _oracle_stepper(o):
    return lambda Z, U: np.stack([o.step(Z[e], U[e])[0] for e in range(Z.shape[0])])


def test_descriptor_sizes():
    """N = 2 / 12 / 8 per contact (impact.jl:38, linear.jl:46, nonlinear.jl:47); solution = joints | bodies | contacts"""
    for name, Ni in (("sphere", 1), ("block", 8)):
        for ct, n in (("impact", 2), ("linear", 12), ("nonlinear", 8)):
            mech = dj.get_mechanism(name, contact_type=ct)
            assert mech.Ni == Ni and mech.nres == 6 + n * Ni and all(c.type == capi_type for c, capi_type in zip(mech.contacts, [dict(impact=0, linear=1, nonlinear=2)[ct]] * Ni))
            desc, keep = capi.flatten(mech)
            assert desc.contacts[0].type == mech.contacts[0].type
    with pytest.raises(ValueError):
        dj.get_mechanism("sphere", contact_type="sticky")


@pytest.mark.parametrize("name", ["sphere", "block"])
@pytest.mark.parametrize("ct", MODELS)
def test_solution_matrix_matches_finite_difference(name, ct):
    """test/jacobian.jl:88-93: the assembled KKT matrix equals -d(rhs)/d(solution) for every contact model"""
    mech = dj.get_mechanism(name, contact_type=ct)
    o = Oracle(mech, capi.solver_options(rtol=1e-7, btol=1e-7))
    u = 0.1 * np.ones(mech.nu)
    z = mech.z0.copy()
    if name == "block":
        z[3:6] = [0.8, 0.3, 0.0]
        z[10:13] = [0.5, -0.3, 0.2]
    for _ in range(100 if name == "block" else 60):
        z, _, _ = o.step(z, u)
    _, _, _, sol = o.step(z, u, return_sol=True)
    mu = o.trace()[-1, 3]
    mu = 0.0 if mu != mu else mu
    o.set_state(z, u)
    o.set_solution(sol, mu)
    A, _ = o.assemble(mu)
    fd = np.zeros_like(A)
    d = 1e-5
    for i in range(mech.nres):
        sp, sm = sol.copy(), sol.copy()
        sp[i] += d
        sm[i] -= d
        fd[:, i] = (o.evaluate_rhs(sp, mu) - o.evaluate_rhs(sm, mu)) / (2 * d)
    assert np.abs(fd + A).max() < 1e-6
    n = mech.contacts[0].dim
    assert np.abs(sol[mech.contact_sol_offset(0) + n // 2]) > 1e-3  # the contact carries load: the blocks are exercised


@pytest.mark.parametrize("name", ["sphere", "block"])
@pytest.mark.parametrize("ct", MODELS)
def test_data_jacobian_matches_finite_difference(name, ct):
    """test/data.jl:28-39: state / control columns of jacobian_data! for every contact model"""
    mech = dj.get_mechanism(name, contact_type=ct)
    o = Oracle(mech, capi.solver_options(rtol=1e-8, btol=1e-8))
    rng = np.random.default_rng(3)
    Z, _ = _thrown(mech, 1, rng, 40, _oracle_stepper(Oracle(mech)))
    z, u = Z[0], np.zeros(mech.nu)
    _, _, _, sol = o.step(z, u, return_sol=True)
    mu = o.trace()[-1, 3]
    mu = 0.0 if mu != mu else mu
    o.set_state(z, u)
    o.set_solution(sol, mu)
    o.assemble(mu)
    D = o.data_jacobian()
    ns, eps = 12 * mech.Nb, 1e-6
    worst = 0.0
    for i in range(ns + mech.nu):
        if i < ns:
            o.set_state(_perturb_state(z, i, eps), u)
            rp = o.evaluate_rhs(sol, mu)
            o.set_state(_perturb_state(z, i, -eps), u)
            rm = o.evaluate_rhs(sol, mu)
        else:
            up, um = u.copy(), u.copy()
            up[i - ns] += eps
            um[i - ns] -= eps
            o.set_state(z, up)
            rp = o.evaluate_rhs(sol, mu)
            o.set_state(z, um)
            rm = o.evaluate_rhs(sol, mu)
        worst = max(worst, np.abs((rp - rm) / (2 * eps) - D[:, i]).max())
    assert worst < 1e-6


@pytest.mark.parametrize("ct", MODELS)
def test_ift_gradients_match_finite_difference(ct):
    """get_maximal_gradients of a sphere resting / rolling on the ground vs central differences of the step"""
    mech = dj.get_mechanism("sphere", contact_type=ct)
    o = Oracle(mech, capi.solver_options(rtol=1e-10, btol=1e-10))
    u = np.zeros(mech.nu)
    z = mech.z0.copy()
    for _ in range(50):
        z, _, _ = o.step(z, u)
    zn, Fz, Fu, st, it0 = o.step_grad(z, u)
    assert st == 0
    eps, checked = 1e-6, 0
    for i in range(12):
        zp, _, ip = o.step(_perturb_state(z, i, eps), u)
        zm, _, im = o.step(_perturb_state(z, i, -eps), u)
        if ip != it0 or im != it0:
            continue
        col = (_reduce(zp, zn, 1) - _reduce(zm, zn, 1)) / (2 * eps)
        assert np.abs(col - Fz[:, i]).max() < 5e-5 * max(1.0, np.abs(Fz).max())
        checked += 1
    assert checked >= 6


def test_friction_models_behave_differently():
    """impact: no friction, the tangential velocity is untouched by the ground; nonlinear / linear: a sliding block stops.
    The pyramid (linear) and the cone (nonlinear) agree along a pyramid axis (friction_cone_comparison.jl)."""
    out = {}
    for ct in MODELS:
        mech = dj.get_mechanism("block", contact_type=ct)
        o = Oracle(mech)
        z = mech.z0.copy()
        z[2] = 0.25 + 1e-3
        z[3:6] = [1.0, 0.0, 0.0]
        for _ in range(80):
            z, st, _ = o.step(z, np.zeros(mech.nu))
            assert st == 0
        out[ct] = z.copy()
        assert z[2] > 0.25 - 1e-6  # no penetration (corner contacts of radius 0)
    assert abs(out["impact"][3] - 1.0) < 1e-6
    assert abs(out["nonlinear"][3]) < 1e-4 and abs(out["linear"][3]) < 1e-4
    assert abs(out["nonlinear"][0] - out["linear"][0]) < 2e-3


# ----------------------------------------------------------------------------------------------------------------
# device code through the kernel emulation
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,ct", [("sphere", "impact"), ("sphere", "linear"), ("block", "impact"), ("block", "linear"), ("block", "nonlinear")])
def test_kernel_emulation_matches_oracle(name, ct):
    """dojo_step_kernel / gradient kernel of the DJ_ANY_CONTACT compilation: status, iteration counts, full solution vector
    (reference ordering [s(N½); gamma(N½)] per contact), next state and IFT gradients against the oracle"""
    from hostemu.harness import HostEmu
    mech = dj.get_mechanism(name, contact_type=ct)
    o, em = Oracle(mech), HostEmu(mech)
    rng = np.random.default_rng(11)
    B = 3
    Z, U = _thrown(mech, B, rng, 0, None)
    for t in range(30):
        Zn, st, it, sol = em.step(Z, U, slots=2)
        for e in range(B):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            assert (st[e], it[e]) == (so, io), (t, e)
            assert np.abs(Zn[e] - zo).max() < 1e-9 and np.abs(sol[e] - solo).max() < 1e-7
        Z = Zn
    Z1, st1, it1, sol1 = em.step(Z, U, slots=1, smem_plan=False)
    Z2, st2, it2, sol2 = em.step(Z, U, slots=4, grid=2)
    assert np.array_equal(Z1, Z2) and np.array_equal(it1, it2) and np.array_equal(sol1, sol2)
    Zn, Fz, Fu, st, it = em.step_grad(Z, U, slots=2, slots_grad=2)
    for e in range(B):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        assert (st[e], it[e]) == (so, io)
        assert np.abs(Fz[e] - Fzo).max() < 1e-7 * max(1.0, np.abs(Fzo).max())
        assert np.abs(Fu[e] - Fuo).max() < 1e-7 * max(1.0, np.abs(Fuo).max())


def test_kernel_emulation_rollout_with_linear_contacts():
    """fused rollout (dojo_rollout) of the block with the friction pyramid == step by step"""
    from hostemu.harness import HostEmu
    mech = dj.get_mechanism("block", contact_type="linear")
    em = HostEmu(mech)
    rng = np.random.default_rng(13)
    B, T = 3, 12
    Z, _ = _thrown(mech, B, rng, 0, None)
    U = 0.1 * rng.normal(size=(T, B, mech.nu))
    Zf = em.step(Z, U, T=T, slots=2)[0]
    Zs = Z
    for t in range(T):
        Zs = em.step(Zs, U[t], slots=2)[0]
    assert np.array_equal(Zf, Zs)
