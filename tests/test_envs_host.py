"""Environment layer -- CPU tests: the device code of dojo_envs.cuh compiled for the host (tests/hostcheck)
around the ORACLE's step, against the literal restatement of DojoEnvironments in oracle/oracle_env.py."""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import environments as E
from oracle.oracle import Oracle
from oracle.oracle_env import env_step

from hostcheck.harness import HostCheck
from test_oracle_properties import _random_minimal


def _spec(cls):
    from dojo_jl_b200 import capi
    return capi.env_spec(**cls.spec_kwargs)


@pytest.mark.parametrize("cls", [E.AntARS, E.QuadrupedSampling, E.Pendulum])
def test_env_pre_post_device_code_on_host(cls):
    mech = dj.get_mechanism(cls.mechanism_name)
    spec = _spec(cls)
    o, hc = Oracle(mech), HostCheck(mech)
    rng = np.random.default_rng(3)
    B = 6
    ns = 2 * mech.nu + (mech.Ni if spec.contact_obs else 0)
    na = mech.nu - spec.n_unactuated
    S = np.zeros((B, ns))
    for e in range(B):
        S[e, :2 * mech.nu] = _random_minimal(mech, rng, 0.2, 0.3)
        if mech.Nb > 1:
            S[e, 2] += rng.uniform(0.2, 0.6)  # some start in contact, some above the ground, one below the healthy range
    A = rng.uniform(-1, 1, (B, na))
    Z, U = hc.env_pre(spec, S, A)
    Zn, sol = np.empty_like(Z), np.empty((B, mech.nres))
    for e in range(B):
        assert np.abs(Z[e] - o.minimal_to_maximal(S[e, :2 * mech.nu])).max() < 1e-12
        assert np.array_equal(U[e], np.concatenate([np.zeros(spec.n_unactuated), A[e]]))
        Zn[e], _, _, sol[e] = o.step(Z[e], U[e], return_sol=True)
    Sn, reward, done = hc.env_post(spec, S, A, Zn, sol)
    for e in range(B):
        sn, r, d, _, _ = env_step(o, spec, S[e], A[e])
        assert np.abs(Sn[e] - sn).max() < 1e-9 * max(1.0, np.abs(sn).max())  # finite-difference velocities: rounding / timestep
        assert abs(reward[e] - r) < 1e-9 * max(1.0, abs(r))
        assert done[e] == d
    if cls is E.AntARS:
        assert np.abs(Sn[:, 2 * mech.nu:]).max() <= 1.0  # clamped contact observations


def test_failure_test_flags_nonfinite_and_out_of_range():
    mech = dj.get_mechanism("ant")
    spec = _spec(E.AntARS)
    hc, o = HostCheck(mech), Oracle(mech)
    x = o.maximal_to_minimal(mech.z0)
    S = np.tile(np.concatenate([x, np.zeros(mech.Ni)]), (3, 1))
    Zn = np.tile(mech.z0, (3, 1))
    Zn[1, 2] = 1.5      # torso above healthy_max
    Zn[2, 0] = np.nan   # non-finite state
    sol = np.zeros((3, mech.nres))
    _, reward, done = hc.env_post(spec, S, np.zeros((3, 8)), Zn, sol)
    assert list(done) == [0, 1, 1]
    assert abs(reward[0] - 0.05) < 1e-9  # no motion, no control, no contact: survive reward only


def test_rollout_return_accumulation_stops_after_failure():
    """rollout_policy (examples/learning/ant_ars.jl:106-115): the reward of the step at which the failure test fires still
    counts, later steps do not."""
    mech = dj.get_mechanism("ant")
    spec = _spec(E.AntARS)
    hc, o = HostCheck(mech), Oracle(mech)
    x = o.maximal_to_minimal(mech.z0)
    S = np.tile(np.concatenate([x, np.zeros(mech.Ni)]), (2, 1))
    sol = np.zeros((2, mech.nres))
    ret, dead = np.zeros(2), np.zeros(2, dtype=np.int32)
    A = np.zeros((2, 8))
    Zok, Zbad = np.tile(mech.z0, (2, 1)), np.tile(mech.z0, (2, 1))
    Zbad[1, 2] = 1.5  # environment 1 leaves the healthy range at the second step
    for Zn in (Zok, Zbad, Zok, Zok):
        hc.env_post(spec, S, A, Zn, sol, ret, dead)
    assert list(dead) == [0, 1]
    assert abs(ret[0] - 4 * 0.05) < 1e-12 and abs(ret[1] - (2 * 0.05 + 100 * (Zbad[1, 0] - x[0]) / mech.timestep)) < 1e-9


def test_linear_policy_device_code_on_host():
    """a = theta * normalize(state)  (examples/learning/ant_ars.jl:47-50, :88)"""
    mech = dj.get_mechanism("ant")
    hc = HostCheck(mech)
    rng = np.random.default_rng(4)
    B, ns, na = 5, 37, 8
    S = rng.normal(size=(B, ns))
    Theta = rng.normal(size=(B, na, ns))
    mean, var = rng.normal(size=ns), rng.uniform(0.01, 2.0, ns)
    A = hc.env_policy(S, Theta, mean, np.sqrt(var))
    ref = np.einsum("bki,bi->bk", Theta, (S - mean) / np.sqrt(var))
    assert np.abs(A - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    assert np.abs(hc.env_policy(S, Theta) - np.einsum("bki,bi->bk", Theta, S)).max() < 1e-12 * 50
