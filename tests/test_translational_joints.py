"""Translational springs, dampers and limits (reference src/joints/translational/{springs,dampers}.jl,
src/joints/limits.jl) -- CPU suite.

Oracle pinned by the reference's property tests on a mechanism that exercises them (a floating body, a Prismatic joint with a
parent BODY and offsets, a Revolute joint; the reference runs test/jacobian.jl and test/data.jl with springs = dampers = 1 on
mechanisms with Prismatic / Planar joints such as :slider, :nslider, :cartpole):
  full_matrix == -d(rhs)/d(solution), jacobian_data! == finite differences, IFT gradients == finite differences of the step.
Device code (dojo_joint_tra.cuh in the DJ_ANY_CONTACT compilation of the kernels) against the oracle through tests/hostemu.
"""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi, quat as Q
from dojo_jl_b200.mechanism import Body, Joint, JointElement, Mechanism
from oracle.oracle import Oracle

from test_oracle_properties import _perturb_state, _reduce


def _element(nlambda, axis=None, damper=0.0, spring=0.0, limits=None, offset=None):
    V1, V2, V3 = Q.orthogonal_rows(np.zeros(3) if axis is None else np.asarray(axis, float))
    e = JointElement(nlambda=nlambda, axis_mask=np.stack([V1, V2, V3]), spring=spring, damper=damper,
                     spring_offset=np.zeros(3 - nlambda) if offset is None else np.asarray(offset, float))
    if limits is not None:
        e.limit_lo, e.limit_hi = np.atleast_1d(limits[0]).astype(float), np.atleast_1d(limits[1]).astype(float)
    return e


def chain(kt=0.0, dt=0.0, kr=0.0, dr=0.0, lim=None, planar=False):
    """origin -Floating- b0 -Prismatic (or Planar) with vertices and an orientation offset- b1 -Revolute- b2"""
    bodies = [Body(f"b{i}", 1.0 + 0.3 * i, np.diag([0.1, 0.2, 0.15]) * (1 + 0.2 * i)) for i in range(3)]
    j0 = Joint("float", -1, 0, _element(0), _element(0))
    tra = _element(1, axis=[0.3, 1.0, 0.2], spring=kt, damper=dt, limits=lim, offset=[0.05, -0.02]) if planar else \
        _element(2, axis=[0.3, 1.0, 0.2], spring=kt, damper=dt, limits=lim, offset=[0.05])
    j1 = Joint("slide", 0, 1, tra, _element(3), vertex_parent=np.array([0.1, 0.2, -0.1]), vertex_child=np.array([-0.05, 0.1, 0.2]),
               orientation_offset=Q.rpy_to_quat([0.2, -0.1, 0.3]))
    j2 = Joint("rev", 1, 2, _element(3), _element(2, axis=[1.0, 0.2, 0.0], spring=kr, damper=dr), vertex_parent=np.array([0.0, 0.1, -0.3]),
               vertex_child=np.array([0.0, 0.0, 0.25]))
    m = Mechanism("chain", bodies, [j0, j1, j2], [], timestep=0.01)
    m.z0 = m.forward_kinematics({"float": [0.1, 0.2, 1.0, 0.2, -0.1, 0.3], "slide": [0.15, 0.05][:2 if planar else 1], "rev": [0.4]})
    return m


def cartpole():
    return dj.get_mechanism("cartpole", springs=2.0, dampers=0.3, joint_limits={"cart_joint": (-0.3, 0.25), "pole_joint": (-1.2, 1.4)})


CASES = {"spring": dict(kt=30.0, kr=0.5, dr=0.2), "damper": dict(dt=2.0, kr=0.5, dr=0.2), "limits": dict(kr=0.5, lim=([-0.05], [0.2])),
         "all": dict(kt=30.0, dt=2.0, kr=0.5, dr=0.2, lim=([-0.05], [0.2])), "planar": dict(kt=10.0, dt=1.0, planar=True),
         "planar_limits": dict(kt=5.0, dt=0.5, planar=True, lim=([-0.05, -0.1], [0.2, 0.12]))}  # two limited translational axes


def _state_in_motion(m, steps, seed=3):
    o = Oracle(m)
    rng = np.random.default_rng(seed)
    z, u = m.z0.copy(), 0.5 * rng.normal(size=m.nu)
    for _ in range(steps):
        z, st, _ = o.step(z, u)
        assert st == 0
    return z, u


def test_cartpole_options_mirror_the_reference_builder():
    m = cartpole()
    cart, pole = m.joints
    assert (cart.tra.spring, cart.tra.damper, cart.tra.nlimits, cart.rot.nlimits) == (2.0, 0.3, 1, 0)
    assert (pole.rot.spring, pole.rot.damper, pole.rot.nlimits, pole.tra.nlimits) == (2.0, 0.3, 1, 0)
    assert m.nres == 30 and m.nu == 2  # cart: 2 + 3 eq + 4 limit entries, pole: 3 + 2 eq + 4, bodies 12
    with pytest.raises(ValueError):
        dj.get_mechanism("sphere", joint_limits={"floating_joint": (0.0, 1.0)})


@pytest.mark.parametrize("case", list(CASES))
def test_solution_matrix_and_data_jacobian_match_finite_differences(case):
    """test/jacobian.jl + test/data.jl on a mechanism with translational springs / dampers / limits"""
    m = chain(**CASES[case])
    z, u = _state_in_motion(m, 40 if "lim" in CASES[case] else 10)
    o = Oracle(m, capi.solver_options(rtol=1e-8, btol=1e-8))
    u0 = np.zeros(m.nu)
    _, _, _, sol = o.step(z, u0, return_sol=True)
    mu = o.trace()[-1, 3]
    mu = 0.0 if mu != mu else mu
    o.set_state(z, u0)
    o.set_solution(sol, mu)
    A, _ = o.assemble(mu)
    D = o.data_jacobian()
    d = 1e-6
    for i in range(m.nres):
        sp, sm = sol.copy(), sol.copy()
        sp[i] += d
        sm[i] -= d
        assert np.abs((o.evaluate_rhs(sp, mu) - o.evaluate_rhs(sm, mu)) / (2 * d) + A[:, i]).max() < 1e-6
    ns = 12 * m.Nb
    for i in range(ns + m.nu):
        if i < ns:
            o.set_state(_perturb_state(z, i, d), u0)
            rp = o.evaluate_rhs(sol, mu)
            o.set_state(_perturb_state(z, i, -d), u0)
            rm = o.evaluate_rhs(sol, mu)
        else:
            up, um = u0.copy(), u0.copy()
            up[i - ns] += d
            um[i - ns] -= d
            o.set_state(z, up)
            rp = o.evaluate_rhs(sol, mu)
            o.set_state(z, um)
            rm = o.evaluate_rhs(sol, mu)
        assert np.abs((rp - rm) / (2 * d) - D[:, i]).max() < 2e-6


def test_ift_gradients_match_finite_difference():
    m = chain(**CASES["all"])
    z, _ = _state_in_motion(m, 40)
    o = Oracle(m, capi.solver_options(rtol=1e-10, btol=1e-10))
    u = np.zeros(m.nu)
    zn, Fz, Fu, st, it0 = o.step_grad(z, u)
    assert st == 0
    eps, checked = 1e-6, 0
    for i in range(12 * m.Nb):
        zp, _, ip = o.step(_perturb_state(z, i, eps), u)
        zm, _, im = o.step(_perturb_state(z, i, -eps), u)
        if ip != it0 or im != it0:
            continue
        col = (_reduce(zp, zn, m.Nb) - _reduce(zm, zn, m.Nb)) / (2 * eps)
        assert np.abs(col - Fz[:, i]).max() < 5e-5 * max(1.0, np.abs(Fz).max())
        checked += 1
    assert checked >= 18


def test_cart_stops_at_its_limits_and_dampers_dissipate():
    m = cartpole()
    o = Oracle(m)
    for push, bound in ((6.0, 0.25), (-6.0, -0.3)):
        z = m.z0.copy()
        for _ in range(200):
            z, st, _ = o.step(z, np.array([push, 0.0]))
            assert st == 0
            assert -0.3 - 1e-4 <= z[1] <= 0.25 + 1e-4  # joints/limits.jl: the slider coordinate stays inside [lo, hi]
        assert abs(z[1] - bound) < 2e-2
    free = dj.get_mechanism("cartpole", gravity=0.0)
    damped = dj.get_mechanism("cartpole", gravity=0.0, dampers=2.0)
    v = []
    for mech in (free, damped):
        oo, z = Oracle(mech), mech.z0.copy()
        z[4] = 1.0  # cart velocity along y
        z[13 + 4] = 1.0
        for _ in range(100):
            z, _, _ = oo.step(z, np.zeros(2))
        v.append(z[4])
    assert abs(v[0] - 1.0) < 1e-6 and abs(v[1] - np.exp(-1.0)) < 1e-2  # m dv/dt = -d v with m = 2, d = 2, t = 1


# ----------------------------------------------------------------------------------------------------------------
# device code through the kernel emulation
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["all", "planar", "planar_limits", "cartpole"])
def test_kernel_emulation_matches_oracle(case):
    from hostemu.harness import HostEmu
    m = cartpole() if case == "cartpole" else chain(**CASES[case])
    o, em = Oracle(m), HostEmu(m)
    rng = np.random.default_rng(7)
    B = 3
    Z = np.tile(m.z0, (B, 1))
    U = (2.0 if case == "planar_limits" else 0.5) * rng.normal(size=(B, m.nu))
    if case == "cartpole":
        U[:, 0] = [8.0, -8.0, 12.0]
    T = 60 if case == "cartpole" else 50 if case == "planar_limits" else 30
    gmax = 0.0
    for t in range(T):
        Zn, st, it, sol = em.step(Z, U, slots=4 if t % 8 == 0 else 1)
        for e in range(B):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            assert (st[e], it[e]) == (so, io), (t, e)
            assert np.abs(Zn[e] - zo).max() < 1e-10 and np.abs(sol[e] - solo).max() < 1e-8
            if case == "cartpole":
                gmax = max(gmax, solo[2], solo[3])
        Z = Zn
    if case == "cartpole":
        assert gmax > 0.5  # the limit duals of the slider are active: the condensed limit rows are exercised
    Zf = em.step(Z, np.tile(U, (6, 1, 1)), T=6, slots=2)[0]
    Zs = Z
    for _ in range(6):
        Zs = em.step(Zs, U, slots=1, smem_plan=False)[0]
    assert np.array_equal(Zf, Zs)
    Zn, Fz, Fu, st, it = em.step_grad(Z, U, slots=2, slots_grad=2)
    for e in range(B):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        assert (st[e], it[e]) == (so, io)
        assert np.abs(Fz[e] - Fzo).max() < 1e-8 * max(1.0, np.abs(Fzo).max())
        assert np.abs(Fu[e] - Fuo).max() < 1e-8 * max(1.0, np.abs(Fuo).max())


@pytest.mark.parametrize("case", ["all", "planar", "cartpole"])
def test_coordinate_maps_with_translational_free_axes(case):
    """minimal <-> maximal maps and their Jacobians (the lane-independent device headers compiled for the host) on Prismatic / Planar
    joints between bodies: what step_minimal_coordinates! and get_minimal_gradients! need around the step"""
    from hostcheck.harness import HostCheck
    from test_oracle_properties import _random_minimal
    m = cartpole() if case == "cartpole" else chain(**CASES[case])
    hc, o = HostCheck(m), Oracle(m)
    rng = np.random.default_rng(1)
    X = np.stack([_random_minimal(m, rng) for _ in range(6)])
    Zo = np.stack([o.minimal_to_maximal(x) for x in X])
    assert np.abs(hc.minimal_to_maximal(X) - Zo).max() < 1e-12 and np.abs(hc.maximal_to_minimal(Zo) - X).max() < 1e-12
    Mo = np.stack([o.maximal_to_minimal_jacobian(z) for z in Zo])
    No = np.stack([o.minimal_to_maximal_jacobian(z) for z in Zo])
    assert np.abs(hc.maximal_to_minimal_jacobian(Zo) - Mo).max() < 1e-11 and np.abs(hc.minimal_to_maximal_jacobian(Zo) - No).max() < 1e-11


def test_cartpole_dqn_mirror_logic(monkeypatch):
    """environments/cartpole_dqn.jl mirror: input_map(a) = [a; 0], state_map = identity, step! = step_minimal_coordinates!.  The
    host logic is exercised with a stand-in for the GPU stepper that answers from the oracle (test-only; on the GPU the same calls go
    to libdojo_b200.so, tests/test_zzzz_gpu_translational.py)."""
    from dojo_jl_b200 import environments as E

    class StubStepper:
        def __init__(self, mech, batch, device=0):
            self.mech, self.o = mech, Oracle(mech)

        def env_sizes(self, spec):
            return 2 * self.mech.nu, self.mech.nu

        def maximal_to_minimal(self, Z):
            return np.stack([self.o.maximal_to_minimal(z) for z in np.atleast_2d(Z)])

        def env_reset(self, spec, S, s0, mask=None):
            S[:] = s0
            return S

        def step_minimal(self, X, U, opts=None):
            out = [self.o.minimal_gradients(x, u) for x, u in zip(X, U)]
            return np.stack([r[0] for r in out]), np.array([r[3] for r in out], dtype=np.int32), np.array([r[4] for r in out], dtype=np.int32)

    monkeypatch.setattr(E, "BatchedStepper", StubStepper)
    wp = E.get_environment("quadruped_waypoint", batch=2)  # quadruped_waypoint.jl:8, :25: timestep 0.001, foot contacts only
    assert wp.mechanism.timestep == 0.001 and wp.mechanism.input_scaling == 0.001 and wp.mechanism.Ni == 4 and wp.ns == 36
    env = E.get_environment("cartpole_dqn", batch=3, dampers=0.5, joint_limits={"cart_joint": (-0.2, 0.2)})
    assert (env.ns, env.na) == (4, 1) and np.allclose(env.get_state()[:, 2], np.pi / 4)  # per joint [coordinates; velocities]: [y, ydot, theta, thetadot]
    assert np.array_equal(env.input_map([1.0, -2.0, 0.5]), [[1.0, 0.0], [-2.0, 0.0], [0.5, 0.0]])
    for _ in range(300):
        reward, done = env.step(action=[12.0, -12.0, 0.0])
    S = env.get_state()
    assert not done.any() and (np.abs(S[:, 0]) < 0.2 + 1e-4).all()  # the slider stays inside its limits ...
    assert abs(S[0, 0] - 0.2) < 3e-2 and abs(S[1, 0] + 0.2) < 3e-2   # ... and is pushed against them


def _hopper_batch(m, B, rng):
    Z = np.tile(m.z0, (B, 1))
    dz = rng.uniform(0.0, 0.3, B)
    Z[:, 2] += dz
    Z[:, 13 + 2] += dz
    U = np.zeros((B, m.nu))
    U[:, 6] = rng.uniform(-10.0, 20.0, B)  # leg force (the Prismatic joint's input)
    return Z, U


def test_raiberthopper_kernel_emulation_matches_oracle():
    """get_raiberthopper defaults (raiberthopper/mechanism.jl:1-86): Floating body, Prismatic leg with a translational damper 0.1,
    foot and body contacts -- the translational damper together with the contact solve"""
    from hostemu.harness import HostEmu
    m = dj.get_mechanism("raiberthopper")
    assert m.joints[1].tra.damper == 0.1 and m.joints[1].tra.nfree == 1 and m.Ni == 2
    o, em = Oracle(m), HostEmu(m)
    Z, U = _hopper_batch(m, 4, np.random.default_rng(2))
    landed = False
    for t in range(25):
        Zn, st, it, sol = em.step(Z, U, slots=4 if t % 8 == 0 else 2)
        for e in range(4):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            assert (st[e], it[e]) == (so, io), (t, e)
            assert np.abs(Zn[e] - zo).max() < 1e-9 and np.abs(sol[e] - solo).max() < 1e-7
            landed = landed or solo[m.contact_sol_offset(0) + 4] > 1.0  # normal impulse of the foot contact
        Z = Zn
    assert landed
    Zn, Fz, Fu, st, it = em.step_grad(Z, U, slots=2, slots_grad=2)
    for e in range(4):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        assert (st[e], it[e]) == (so, io)
        assert np.abs(Fz[e] - Fzo).max() < 1e-7 * max(1.0, np.abs(Fzo).max()) and np.abs(Fu[e] - Fuo).max() < 1e-7 * max(1.0, np.abs(Fuo).max())


@pytest.mark.parametrize("case", ["all", "planar", "cartpole", "raiberthopper"])
def test_storage_with_translational_impulses(case):
    """save_to_storage! / momentum / energies (dojo_storage.cuh compiled for the host) with translational spring, damper and limit
    impulses, against the oracle's restatement (simulation/storage.jl:50-67, mechanics/momentum.jl:44-52, energy.jl:69-90)"""
    from hostcheck.harness import HostCheck
    m = cartpole() if case == "cartpole" else dj.get_mechanism("raiberthopper") if case == "raiberthopper" else chain(**CASES[case])
    o, hc = Oracle(m), HostCheck(m)
    rng = np.random.default_rng(5)
    z, u = m.z0.copy(), 0.5 * rng.normal(size=m.nu)
    if case == "cartpole":
        u[0] = 6.0  # against the upper limit of the slider
    for t in range(120 if case == "cartpole" else 20):
        zn, st, _, sol = o.step(z, u, return_sol=True)
        body, diag = o.storage_record()
        bh, dh = hc.storage(z, zn, u, sol)
        assert np.abs(bh[0] - body).max() < 1e-11 * max(1.0, np.abs(body).max())
        assert np.abs(dh[0] - diag).max() < 1e-11 * max(1.0, np.abs(diag).max())
        z = zn


@pytest.mark.parametrize("case", ["spring_against_limit", "planar_limits"])
def test_momentum_conservation_with_translational_springs_dampers_limits(case):
    """test/momentum.jl (run there on :slider / :nslider / :npendulum-like mechanisms with springs = dampers > 0): in zero gravity
    the spring, damper and limit impulses of a joint between two bodies (and rotational inputs) are internal -- linear and angular momentum of
    the floating chain stay constant.  Pins the impulse transforms of the translational terms (equal and opposite forces AND the
    torques that go with them) independently of the finite-difference tests."""
    # spring_against_limit: the spring (rest position 0.05) pulls the slider against its lower limit 0.1
    m = chain(kt=30.0, dt=2.0, kr=0.5, dr=0.2, lim=([0.1], [0.2])) if case == "spring_against_limit" else chain(**CASES[case])
    m.gravity = np.zeros(3)
    o = Oracle(m, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = m.z0.copy()
    zz = z.reshape(-1, 13)
    zz[0, 3:6] = [0.3, -0.2, 0.1]   # the floating base moves and spins; the other bodies follow through the joints
    zz[0, 10:13] = [0.4, 0.3, -0.5]
    for _ in range(3):              # a few steps make the velocities of the chain consistent with its joints
        z, st, _ = o.step(z, np.zeros(m.nu))
    P, active = [], False
    for k in range(150):
        u = np.zeros(m.nu)
        u[-1] = 1.5 if k < 60 else 0.0   # the Revolute joint's input (internal).  NOT the Prismatic / Planar joint's: the reference halves
        # the torque part of a translational input (translational/input.jl:21, :23: `Jτ2 += Jτaa / 2`), so a translational input with a
        # lever arm does not conserve angular momentum -- replicated on the device as in the reference
        z, st, _, sol = o.step(z, u, return_sol=True)
        assert st == 0
        P.append(o.momentum())
        jo = m.joint_sol_offset(1)
        nb = 2 * m.joints[1].tra.nlimits
        active = active or (nb > 0 and sol[jo + nb: jo + 2 * nb].max() > 0.05)
    P = np.array(P)
    assert np.abs(P - P[0]).max() < 1e-8
    assert active  # the limit duals were at work during the run


def test_slider_known_answers_of_the_reference():
    """test/energy.jl "Slider 1" (:188-231): a box on a Prismatic joint with a translational spring k = 10 in zero gravity,
    released 0.5 from the spring's rest position, is a harmonic oscillator -- the reference asserts the ANALYTIC amplitude and
    peak velocity (z0, z0 sqrt(k / m), both to 1e-4) and a mechanical-energy band of 1e-3 over 5 s.  One of the few known-answer
    tests of the reference on this path; run here on the oracle and on the device kernels (emulation)."""
    from hostemu.harness import HostEmu
    m = dj.get_mechanism("slider", gravity=0.0, springs=10.0)
    z0, k, mass = 0.5, 10.0, m.bodies[0].mass
    o, em = Oracle(m, capi.solver_options(rtol=1e-10, btol=1e-10)), HostEmu(m)
    z = m.forward_kinematics({"joint": [-z0]})  # initialize!(mech, :slider; position = z0) (slider/mechanism.jl:41-53)
    zd = z.copy()[None]
    xs, vs, me, xd = [], [], [], []
    for step in range(500):
        z, st, _ = o.step(z, np.zeros(1))
        assert st == 0
        body, diag = o.storage_record()
        xs.append(z[2])
        vs.append(body[0][8])
        me.append(diag[6] + diag[7])
        zd = em.step(zd, np.zeros((1, 1)), capi.solver_options(rtol=1e-10, btol=1e-10))[0]
        xd.append(zd[0, 2])
    me = np.array(me[100:])
    assert abs(max(xs) - z0 + 0.5) < 1e-4                      # maximum amplitude (the body centre sits 0.5 below the joint)
    assert abs(max(vs) - z0 * np.sqrt(k / mass)) < 1e-4        # maximum velocity
    assert np.abs((me - me[0]) / me.mean()).max() < 1e-3       # no energy drift with the variational integrator
    assert np.abs(np.array(xd) - np.array(xs)).max() < 1e-10   # the device kernels follow the same trajectory


@pytest.mark.parametrize("gravity,spring,z0,seconds,band", [(-9.81, 0.0, 0.5, 1.5, 1e-6), (-9.81, 1.0, 0.1, 10.0, 1e-3)])
def test_slider_energy_conservation_of_the_reference(gravity, spring, z0, seconds, band):
    """test/energy.jl "Slider 2" (:243-277: gravity only, mechanical energy constant to 1e-6) and "Slider 3" (:288-321: gravity and a
    translational spring, band 1e-3), energies as the reference records them (save_to_storage!, mechanics/energy.jl)"""
    m = dj.get_mechanism("slider", gravity=gravity, springs=spring)
    o = Oracle(m, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = m.forward_kinematics({"joint": [-z0]})
    me = []
    for _ in range(int(seconds / m.timestep)):
        z, st, _ = o.step(z, np.zeros(1))
        assert st == 0
        _, diag = o.storage_record()
        me.append(diag[6] + diag[7])
    me = np.array(me[100:])
    assert np.abs((me - me[0]) / me.mean()).max() < band
