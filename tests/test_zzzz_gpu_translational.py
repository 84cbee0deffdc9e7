"""GPU tests of translational springs / dampers / limits through the C-ABI against the oracle; the same
comparison runs on the CPU through the kernel emulation (tests/test_translational_joints.py)."""
import numpy as np
import pytest

from test_translational_joints import CASES, cartpole, chain

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["all", "planar", "cartpole"])
def test_step_rollout_and_gradient_parity(case):
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    m = cartpole() if case == "cartpole" else chain(**CASES[case])
    rng = np.random.default_rng(23)
    B = 32
    stepper, o = BatchedStepper(m, B), Oracle(m)
    Z = np.tile(m.z0, (B, 1))
    U = 0.5 * rng.normal(size=(B, m.nu))
    if case == "cartpole":
        U[:, 0] = rng.uniform(-6.0, 6.0, B)
    T = 100 if case == "cartpole" else 40
    same = total = 0
    for t in range(T):
        Zn, st, it, sol = stepper.step(Z, U, return_sol=True)
        for e in range(0, B, 4):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            total += 1
            assert st[e] == so == 0
            if it[e] != io:
                assert np.abs(Zn[e] - zo).max() < 1e-4
                continue
            same += 1
            assert np.abs(Zn[e] - zo).max() < 1e-8 and np.abs(sol[e] - solo).max() < 1e-6
        Z = Zn
    assert same >= 0.95 * total
    if case == "cartpole":
        assert (Z[:, 1] > -0.3 - 1e-4).all() and (Z[:, 1] < 0.25 + 1e-4).all() and (np.abs(Z[:, 1]) > 0.2).sum() >= B // 4
    Zf, _ = stepper.rollout(Z, np.tile(U, (5, 1, 1)), T=5)
    Zs = Z
    for _ in range(5):
        Zs, _, _ = stepper.step(Zs, U)
    assert np.array_equal(Zf, Zs)
    Zn, Fz, Fu, st, it = stepper.step_grad(Z, U)
    errs = []
    for e in range(0, B, 4):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        errs.append(max(np.abs(Fz[e] - Fzo).max() / max(1.0, np.abs(Fzo).max()), np.abs(Fu[e] - Fuo).max() / max(1.0, np.abs(Fuo).max())))
    errs = np.array(errs)
    assert len(errs) >= 6 and np.median(errs) < 1e-8 and errs.max() < 1e-4, errs


def test_recording_with_translational_impulses():
    """dojo_step_record on the cartpole with springs, dampers and limits: momentum / energies against the oracle"""
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    m = cartpole()
    B = 8
    stepper, o = BatchedStepper(m, B), Oracle(m)
    rng = np.random.default_rng(37)
    Z = np.tile(m.z0, (B, 1))
    U = np.zeros((B, m.nu))
    U[:, 0] = rng.uniform(-6.0, 6.0, B)
    for _ in range(40):
        Z, _, _ = stepper.step(Z, U)
    out = stepper.step_record(Z, U)
    Zn, storage, diag, st, it = out
    compared = 0
    for e in range(B):
        zo, so, io = o.step(Z[e], U[e])
        body, dg = o.storage_record()
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        compared += 1
        assert np.abs(Zn[e] - zo).max() < 1e-8
        assert np.abs(np.asarray(storage[e]).reshape(-1) - body.reshape(-1)).max() < 1e-6 * max(1.0, np.abs(body).max())
        assert np.abs(np.asarray(diag[e]) - dg).max() < 1e-6 * max(1.0, np.abs(dg).max())
    assert compared >= B - 2


def test_cartpole_environment_and_minimal_gradients():
    """cartpole_dqn mirror (state_map, input_map [a; 0], step_minimal_coordinates!) and get_minimal_gradients! on a mechanism with
    a Prismatic joint, translational spring / damper / limits, against the oracle"""
    from dojo_jl_b200 import environments as E
    from oracle.oracle import Oracle
    B = 16
    env = E.get_environment("cartpole_dqn", batch=B, springs=2.0, dampers=0.3, joint_limits={"cart_joint": (-0.3, 0.25)})
    m = env.mechanism
    o = Oracle(m)
    rng = np.random.default_rng(29)
    a = rng.uniform(-4.0, 4.0, B)
    S = env.get_state()
    for _ in range(60):
        env.step(action=a)
    Sn = env.get_state()
    assert np.isfinite(Sn).all() and (Sn[:, 0] > -0.3 - 1e-4).all() and (Sn[:, 0] < 0.25 + 1e-4).all()
    for e in range(0, B, 4):
        x = S[e]
        for _ in range(60):
            x = o.minimal_gradients(x, np.array([a[e], 0.0]))[0]
        assert np.abs(x - Sn[e]).max() < 1e-6
    Xn, Gx, Gu, st, it = env.stepper.minimal_gradients(Sn, env.input_map(a))
    for e in range(0, B, 4):
        xo, Gxo, Guo, so, io = o.minimal_gradients(Sn[e], np.array([a[e], 0.0]))
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        assert np.abs(Xn[e] - xo).max() < 1e-8
        assert np.abs(Gx[e] - Gxo).max() < 1e-6 * max(1.0, np.abs(Gxo).max()) and np.abs(Gu[e] - Guo).max() < 1e-6 * max(1.0, np.abs(Guo).max())


def test_raiberthopper_parity():
    """get_raiberthopper defaults: translational damper on the Prismatic leg + foot / body contacts"""
    import dojo_jl_b200 as dj
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    from test_translational_joints import _hopper_batch
    m = dj.get_mechanism("raiberthopper")
    B = 64
    stepper, o = BatchedStepper(m, B), Oracle(m)
    Z, U = _hopper_batch(m, B, np.random.default_rng(31))
    same = conv = total = status_diff = 0
    for t in range(40):
        Zn, st, it, sol = stepper.step(Z, U, return_sol=True)
        for e in range(0, B, 4):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            total += 1
            if st[e] != so:
                status_diff += 1
                continue
            if so != 0:
                continue
            conv += 1
            if it[e] != io:
                assert np.abs(Zn[e] - zo).max() < 2e-2
                continue
            same += 1
            assert np.abs(Zn[e] - zo).max() < 1e-6 and np.abs(sol[e] - solo).max() < 1e-5
        Z = Zn
    assert status_diff <= 0.03 * total and conv >= 0.7 * total and same >= 0.9 * conv, (same, conv, status_diff, total)
    Zn, Fz, Fu, st, it = stepper.step_grad(Z, U)
    errs = []
    for e in range(0, B, 4):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        errs.append(max(np.abs(Fz[e] - Fzo).max() / max(1.0, np.abs(Fzo).max()), np.abs(Fu[e] - Fuo).max() / max(1.0, np.abs(Fuo).max())))
    errs = np.array(errs)
    assert len(errs) >= 8 and np.median(errs) < 1e-7 and np.quantile(errs, 0.9) < 1e-4 and errs.max() < 1e-2, errs


def test_quadruped_waypoint_environment():
    """environments/quadruped_waypoint.jl (timestep 0.001, foot contacts only) on the fused environment kernels against the oracle"""
    from dojo_jl_b200 import environments as E
    from oracle.oracle import Oracle
    from oracle.oracle_env import env_step
    from test_gpu_parity import _random_minimal_batch
    rng = np.random.default_rng(43)
    B = 16
    env = E.get_environment("quadruped_waypoint", batch=B)
    mech, spec = env.mechanism, env.spec
    assert mech.timestep == 0.001 and mech.Ni == 4
    o = Oracle(mech)
    S = np.zeros((B, env.ns))
    S[:, :2 * mech.nu] = _random_minimal_batch(mech, B, rng)
    S[:, 2] += rng.uniform(0.3, 0.5, B)
    A = rng.uniform(-0.2, 0.2, (B, env.na))
    for _ in range(10):  # a few steps away from the random start (feet inside the ground)
        S = env.stepper.env_step(spec, S, A)[0]
    Sn, reward, done, status, iters = env.stepper.env_step(spec, S, A)
    compared = 0
    for e in range(B):
        sn, r, d, so, io = env_step(o, spec, S[e], A[e])
        if so != 0 or status[e] != 0 or io != iters[e]:
            continue
        assert np.abs(Sn[e] - sn).max() < 1e-6 * max(1.0, np.abs(sn).max()) and done[e] == d
        compared += 1
    assert compared >= B // 2
