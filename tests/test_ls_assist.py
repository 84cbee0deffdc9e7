"""Line-search assist (round 2): when one slot of a CTA still has an environment and the others have drained -- the tail of a launch --
the drained slots evaluate further line-search trials of that environment's Newton iteration at the same time (ls_assist_loop,
dojo.jl_b200/csrc/dojo_kernels.cuh).  The trials are the ones the owner would have evaluated itself in later passes, so status,
iteration counts, states and the whole solution vector must be BIT-identical with and without helpers.  Run on the kernel emulation
(tests/hostemu): the CTA-wide request / result protocol through the alignment barrier is exactly the part a missing synchronisation
would break, and the emulation makes that a reproducible difference (thread orders: tests/test_hostemu_orders.py)."""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from oracle.oracle import Oracle

from conftest import jittered_states, random_inputs
from hostemu.harness import HostEmu


def _emulators(mech, monkeypatch):
    monkeypatch.delenv("DOJO_B200_NO_LS_ASSIST", raising=False)
    on = HostEmu(mech)
    monkeypatch.setenv("DOJO_B200_NO_LS_ASSIST", "1")
    off = HostEmu(mech)
    monkeypatch.delenv("DOJO_B200_NO_LS_ASSIST", raising=False)
    return on, off


@pytest.mark.parametrize("name,B,slots,tight", [("ant", 1, 4, False), ("ant", 3, 4, True), ("ant", 5, 4, False), ("quadruped", 2, 4, True), ("atlas", 1, 2, False)])
def test_assisted_line_search_is_bit_identical(name, B, slots, tight, monkeypatch):
    mech = dj.get_mechanism(name)
    on, off = _emulators(mech, monkeypatch)
    o = Oracle(mech)
    rng = np.random.default_rng(11)
    Z = jittered_states(mech, B, rng)
    for _ in range(5):
        Z = np.stack([o.step(Z[e], random_inputs(mech, 1, rng)[0])[0] for e in range(B)])
    U = random_inputs(mech, B, rng)
    opts = capi.solver_options(rtol=1e-9, btol=1e-9) if tight else None  # tight tolerances: more line-search trials per iteration
    a, b = on.step(Z, U, opts, slots=slots), off.step(Z, U, opts, slots=slots)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    T = 3
    U3 = np.stack([random_inputs(mech, B, rng) for _ in range(T)])
    a, b = on.step(Z, U3, opts, T=T, slots=slots, record=True), off.step(Z, U3, opts, T=T, slots=slots, record=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_stalled_environment_with_three_helpers(monkeypatch):
    """An environment that rejects every line-search trial (all ten evaluated in two passes by four slots instead of five passes by
    one) and runs into max_iter: identical iterate after 50 iterations, identical status, equal to the oracle's verdict."""
    mech = dj.get_mechanism("ant")
    on, off = _emulators(mech, monkeypatch)
    o = Oracle(mech, capi.solver_options(rtol=1e-14, btol=1e-14, max_iter=12))
    rng = np.random.default_rng(3)
    Z = jittered_states(mech, 1, rng)
    oo = Oracle(mech)
    for _ in range(8):
        Z = np.stack([oo.step(Z[0], random_inputs(mech, 1, rng)[0])[0]])
    U = random_inputs(mech, 1, rng)
    opts = capi.solver_options(rtol=1e-14, btol=1e-14, max_iter=12)  # unreachable tolerances: the solve stalls on its rounding floor
    a, b = on.step(Z, U, opts, slots=4), off.step(Z, U, opts, slots=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    zo, so, io = o.step(Z[0], U[0])
    assert a[1][0] == so and a[2][0] == io
