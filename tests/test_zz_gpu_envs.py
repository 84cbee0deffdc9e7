"""GPU tests of the batched environment layer through the C-ABI: dojo_env_step / dojo_env_reset against the
literal restatement of DojoEnvironments in oracle/oracle_env.py and against the composition of the library's own calls."""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200 import environments as E
from test_gpu_parity import _random_minimal_batch

pytestmark = pytest.mark.gpu


def _batch_states(env, rng):
    mech = env.mechanism
    B = env.batch
    S = np.zeros((B, env.ns))
    S[:, :2 * mech.nu] = _random_minimal_batch(mech, B, rng)
    if mech.Nb > 1:
        S[:, 2] += rng.uniform(0.3, 0.7, B)
    return S


@pytest.mark.parametrize("name", ["ant_ars", "quadruped_sampling", "pendulum"])
def test_env_step_matches_reference_semantics(name):
    from oracle.oracle import Oracle
    from oracle.oracle_env import env_step
    rng = np.random.default_rng(61)
    B = 24
    env = E.get_environment(name, batch=B)
    mech, spec = env.mechanism, env.spec
    o = Oracle(mech)
    S = _batch_states(env, rng)
    A = rng.uniform(-1, 1, (B, env.na))
    Sn, reward, done, status, iters = env.stepper.env_step(spec, S, A)
    # composition of the library's own calls: state_map, input_map, step_minimal
    Xn, st2, it2 = env.stepper.step_minimal(env.state_map(S), env.input_map(A))
    # The fused pre / post kernels inline the same map functions as the stand-alone map kernels, but the two compilations need
    # not round identically (FMA contraction depends on the inlining context); a last-bit difference of z is amplified by a
    # contact solve like any other rounding difference: typical 1e-13, single environments up to solver
    # tolerance.  Measured on B200 (round 1): not bit-identical for ant / quadruped, bit-identical for the pendulum.
    same = (status == st2) & (iters == it2)
    assert same.mean() >= 0.9
    cerr = np.abs(Sn[:, :2 * mech.nu] - Xn).max(axis=1) / np.maximum(1.0, np.abs(Xn).max(axis=1))
    assert np.median(cerr[same]) < 1e-10 and np.quantile(cerr[same], 0.9) < 1e-6 and cerr[same].max() < 5e-3, cerr
    compared = 0
    for e in range(B):
        sn, r, d, so, io = env_step(o, spec, S[e], A[e])
        if so != 0 or status[e] != 0 or io != iters[e]:
            continue
        scale = max(1.0, np.abs(sn).max())
        assert np.abs(Sn[e] - sn).max() < 1e-6 * scale
        assert abs(reward[e] - r) < 1e-5 * max(1.0, abs(r)) and done[e] == d
        compared += 1
    assert compared >= B // 2
    if spec.contact_obs:
        assert np.abs(Sn[:, 2 * mech.nu:]).max() <= 1.0


def test_environment_mirror_rollout_and_reset():
    """AntARS: hold the state, step with zero actions from initialize!, reset the failed environments"""
    rng = np.random.default_rng(67)
    env = E.get_environment("ant_ars", batch=32)
    s0 = env.initial_state()
    assert np.array_equal(env.get_state(), np.tile(s0, (32, 1)))
    total = np.zeros(32)
    for k in range(5):
        reward, done = env.step(action=rng.uniform(-1, 1, (32, env.na)))
        total += reward
    assert np.isfinite(total).all() and (env.status <= 1).all()  # :success or, rarely, :failed (max_iter); never NaN / excessive velocity
    mask = np.zeros(32, dtype=np.int32)
    mask[::2] = 1
    before = env.get_state()
    env.initialize(mask)
    after = env.get_state()
    assert np.array_equal(after[::2], np.tile(s0, (16, 1))) and np.array_equal(after[1::2], before[1::2])


def test_env_step_device_pointers_and_reset_on_device():
    import torch
    rng = np.random.default_rng(71)
    B = 128
    env = E.get_environment("ant_ars", batch=B)
    spec = env.spec
    S = _batch_states(env, rng)
    A = rng.uniform(-1, 1, (B, env.na))
    Sn, reward, done, status, iters = env.stepper.env_step(spec, S, A)
    dS, dA = torch.from_numpy(S).cuda(), torch.from_numpy(A).cuda()
    dSn = torch.empty_like(dS)
    dR = torch.empty(B, dtype=torch.float64, device="cuda")
    dD = torch.empty(B, dtype=torch.int32, device="cuda")
    dSt = torch.empty(B, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    env.stepper.env_step_device(spec, dS.data_ptr(), dA.data_ptr(), dSn.data_ptr(), B, dreward=dR.data_ptr(), ddone=dD.data_ptr(), dstatus=dSt.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert np.array_equal(dSn.cpu().numpy(), Sn) and np.array_equal(dR.cpu().numpy(), reward)
    assert np.array_equal(dD.cpu().numpy(), done) and np.array_equal(dSt.cpu().numpy(), status)
    mask = torch.zeros(B, dtype=torch.int32, device="cuda")
    mask[: B // 2] = 1
    env.stepper.env_reset(spec, (dSn.data_ptr(), B), env.initial_state(), mask.data_ptr())
    out = dSn.cpu().numpy()
    assert np.array_equal(out[: B // 2], np.tile(env.initial_state(), (B // 2, 1))) and np.array_equal(out[B // 2:], Sn[B // 2:])


def test_env_rollout_equals_stepwise_accumulation():
    """dojo_env_rollout == T calls of dojo_env_step with the return accumulated on the host until the failure test fires"""
    rng = np.random.default_rng(73)
    B, T = 48, 12
    env = E.get_environment("ant_ars", batch=B)
    spec = env.spec
    S0 = _batch_states(env, rng)
    S0[:4, 2] = 0.25  # close to the lower end of the healthy range: some of these fail during the rollout
    A = rng.uniform(-1, 1, (T, B, env.na))
    Sf, ret, failed = env.stepper.env_rollout(spec, S0, A, T)
    S, acc, dead = S0, np.zeros(B), np.zeros(B, dtype=bool)
    for k in range(T):
        S, r, d, _, _ = env.stepper.env_step(spec, S, A[k])
        acc += np.where(dead, 0.0, r)
        dead |= d.astype(bool)
    assert np.array_equal(Sf, S) and np.array_equal(failed.astype(bool), dead)
    assert np.abs(ret - acc).max() < 1e-9 * max(1.0, np.abs(acc).max())
    env2 = E.get_environment("ant_ars", batch=B)
    env2.state = S0.copy()
    ret2, failed2 = env2.rollout(A)
    assert np.array_equal(ret2, ret) and np.array_equal(env2.get_state(), Sf)


def test_env_policy_rollout_equals_host_policy_loop():
    """dojo_env_policy_rollout == a host loop that evaluates a = Theta_e ((s - mean) / std) and calls dojo_env_step"""
    rng = np.random.default_rng(79)
    B, T = 32, 10
    env = E.get_environment("ant_ars", batch=B)
    spec, ns, na = env.spec, env.ns, env.na
    S0 = np.tile(env.initial_state(), (B, 1))
    Theta = 0.2 * rng.normal(size=(B, na, ns))
    mean, std = 0.1 * rng.normal(size=ns), np.sqrt(rng.uniform(0.01, 1.0, ns))
    Sf, ret, failed, traj = env.stepper.env_policy_rollout(spec, S0, Theta, T, mean, std, record_states=True)
    S, acc, dead = S0, np.zeros(B), np.zeros(B, dtype=bool)
    for k in range(T):
        # traj[k] = device policy + device step of traj[k-1]; S = host policy + dojo_env_step of traj[k-1]
        kerr = np.abs(traj[k] - S).max(axis=1) / max(1.0, np.abs(S).max())
        assert np.quantile(kerr, 0.9) < 1e-6 and kerr.max() < 5e-3, (k, kerr)
        A = np.einsum("bki,bi->bk", Theta, (traj[k] - mean) / std)   # same observations as the device (summation order may differ)
        S, r, d, _, _ = env.stepper.env_step(spec, traj[k], A)
        acc += np.where(dead, 0.0, r)
        dead |= d.astype(bool)
    # the host evaluates the policy with a different summation order: actions agree to rounding; a rounding-level flip of a
    # line-search comparison may move single environments to solver tolerance
    scale = max(1.0, np.abs(S).max())
    err = np.abs(Sf - S).max(axis=1) / scale
    assert np.quantile(err, 0.9) < 1e-6 and err.max() < 5e-3, err
    assert (failed.astype(bool) == dead).mean() >= 0.9
    rerr = np.abs(ret - acc) / max(1.0, np.abs(acc).max())
    assert np.quantile(rerr, 0.9) < 1e-6, rerr
