"""Race detection with the kernel emulation: between two synchronisation points the result of a kernel must not depend on the order in
which its threads run.  tests/hostemu runs the threads of a scheduling round in ascending order; HOSTEMU_ORDER=reverse / random
(read once per process, hence the subprocesses) runs them in descending / pseudo-random order.  The results must be BIT-IDENTICAL:
a difference means a thread read something another thread writes without a barrier in between (in the step kernel, the gradient
kernel, the paired line search, the slot hand-over ...)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import dojo_jl_b200 as dj
from hostemu.harness import HostEmu
from conftest import jittered_states, random_inputs
out = {}
for name, kw in (("ant", {}), ("block", {"contact_type": "linear"}), ("raiberthopper", {})):
    m = dj.get_mechanism(name, **kw)
    em = HostEmu(m)
    rng = np.random.default_rng(41)
    B = 3
    Z = jittered_states(m, B, rng) if m.Nb > 2 else np.tile(m.z0, (B, 1))
    U = random_inputs(m, B, rng)
    for t in range(3):
        Z = em.step(Z, U, slots=4)[0]
    Zf = em.step(Z, np.tile(U, (3, 1, 1)), T=3, slots=2, grid=2)[0]
    Zn, Fz, Fu, st, it = em.step_grad(Z, U, slots=2, slots_grad=2 if name != "quadruped" else 1)
    out[name + "_Z"], out[name + "_Zf"], out[name + "_Fz"], out[name + "_Fu"], out[name + "_it"] = Z, Zf, Fz, Fu, it
np.savez(sys.argv[1], **out)
"""


def _run(order, path):
    env = dict(os.environ)
    env.pop("HOSTEMU_ORDER", None)
    if order:
        env["HOSTEMU_ORDER"] = order
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, path], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    if order:
        assert "thread order of a round = " + order in r.stderr
    return np.load(path)


def test_kernels_do_not_depend_on_the_thread_order(tmp_path):
    ref = _run(None, str(tmp_path / "asc.npz"))
    for order in ("reverse", "random"):
        got = _run(order, str(tmp_path / (order + ".npz")))
        for k in ref.files:
            assert np.array_equal(ref[k], got[k]), (order, k)


RACY = r"""
#include "%(shim)s"
// thread 0 publishes a value, every other thread reads it -- `with_barrier` selects the correct version
static void kernel(bool with_barrier, double* out) {
  double* smem = hostemu_smem();
  const int t = threadIdx.x;
  if (t == 0) smem[0] = 42.0;
  if (with_barrier) __syncthreads();
  out[t] = smem[0];
}
extern "C" void run(int with_barrier, double* out) {
  emu::run_cta(0, 1, 64, 64, [=] { kernel(with_barrier != 0, out); });
}
"""


def test_thread_order_modes_detect_a_missing_barrier(tmp_path):
    """negative control of the detector: a kernel without its barrier gives different answers under different thread orders, the
    same kernel with the barrier does not"""
    import ctypes
    src, lib = tmp_path / "racy.cpp", tmp_path / "libracy.so"
    src.write_text(RACY % {"shim": os.path.join(ROOT, "tests", "hostemu", "cuda_shim.h")})
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(lib), str(src)])
    prog = ("import ctypes, numpy as np, sys\nL = ctypes.CDLL(sys.argv[1])\nfor wb in (0, 1):\n    o = np.zeros(64)\n"
            "    L.run(wb, ctypes.c_void_p(o.ctypes.data))\n    print(int((o == 42.0).sum()))\n")
    res = {}
    for order in ("", "reverse"):
        env = dict(os.environ)
        env.pop("HOSTEMU_ORDER", None)
        if order:
            env["HOSTEMU_ORDER"] = order
        r = subprocess.run([sys.executable, "-c", prog, str(lib)], env=env, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        res[order] = [int(x) for x in r.stdout.split()]
    assert res[""][1] == 64 and res["reverse"][1] == 64      # with the barrier: every thread sees the value, in any order
    assert res[""][0] == 64 and res["reverse"][0] == 1       # without it: the ascending order hides the race, the reverse order exposes it
