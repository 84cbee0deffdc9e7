"""Host-side checks that run without a GPU: the C-ABI library loads and exports every symbol that
include/dojo_b200.h declares; descriptor flattening; the package refuses to run without its CUDA library/device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dojo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dojo_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    L = C.CDLL(solver.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/dojo_b200.h but not exported"
    assert sorted(solver.EXPORTS) == syms


def test_default_options_match_reference():
    """SolverOptions defaults, src/solver/options.jl:16-26."""
    L = solver.load_library()
    o = capi.DojoSolverOptions()
    L.dojo_default_options(C.byref(o))
    assert (o.rtol, o.btol, o.ls_scale, o.max_iter, o.max_ls) == (1e-6, 1e-4, 0.5, 50, 10)
    assert o.undercut == float("inf") and o.no_progress_max == 3 and o.no_progress_undercut == 10.0
    p = capi.solver_options()
    assert (p.rtol, p.btol, p.max_iter, p.max_ls, p.undercut) == (o.rtol, o.btol, o.max_iter, o.max_ls, o.undercut)


@pytest.mark.parametrize("name,Nb,Ne,Ni,nres,nu", [("pendulum", 1, 1, 0, 11, 1), ("ant", 13, 13, 9, 246, 14),
                                                    ("quadruped", 13, 13, 12, 282, 18), ("atlas", 31, 31, 20, 496, 36)])
def test_mechanism_sizes(name, Nb, Ne, Ni, nres, nu):
    """SURVEY.md Appendix B problem sizes."""
    m = dj.get_mechanism(name)
    assert (m.Nb, m.Ne, m.Ni, m.nres, m.nu, m.nz) == (Nb, Ne, Ni, nres, nu, 13 * Nb)
    d, keep = capi.flatten(m)
    assert d.num_bodies == Nb and d.num_joints == Ne and d.num_contacts == Ni
    assert abs(d.timestep - m.timestep) == 0


def test_forward_kinematics_satisfies_joint_constraints():
    from oracle.oracle import Oracle
    for name in ("ant", "quadruped", "atlas"):
        m = dj.get_mechanism(name)
        o = Oracle(m)
        o.set_state(m.z0, np.zeros(m.nu))
        o.reset_solution()
        sol = o.get_solution()
        # zero velocities: next configuration = current configuration, so the joint equality rows vanish
        for b in range(m.Nb):
            sol[m.body_sol_offset(b): m.body_sol_offset(b) + 6] = 0
        o.set_solution(sol, 0.0)
        rhs = o.evaluate_rhs(sol, 0.0)
        for ji, j in enumerate(m.joints):
            off = m.joint_sol_offset(ji)
            eq = list(range(off, off + j.tra.nlambda)) + list(range(off + j.tra.nimpulses + 4 * j.rot.nlimits, off + j.nimpulses))
            assert np.abs(rhs[eq]).max(initial=0.0) < 1e-12


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly, never fall back to the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    m = dj.get_mechanism("pendulum")
    with pytest.raises(RuntimeError, match="no usable CUDA device|CPU fallback"):
        solver.BatchedStepper(m, 4)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dojo.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", "").lower() or f in ("capi.py",), f


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """every struct of include/dojo_b200.h: size and the offset of each field as the C compiler lays them out == the ctypes mirror in
    dojo.jl_b200/capi.py (what the Python host, the tests and the bench pass through the C-ABI)"""
    import ctypes as C
    import subprocess
    from dojo_jl_b200 import capi
    names = ["DojoBodyDesc", "DojoJointElementDesc", "DojoJointDesc", "DojoContactDesc", "DojoMechanismDesc", "DojoSolverOptions", "DojoEnvSpec"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "dojo_b200.h")}"', "int main(void) {"]
    for n in names:
        st = getattr(capi, n)
        lines.append(f'  printf("{n} %zu\\n", sizeof({n}));')
        for f, _ in st._fields_:
            lines.append(f'  printf("{n}.{f} %zu\\n", offsetof({n}, {f}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        st = getattr(capi, n)
        assert int(out[n]) == C.sizeof(st), n
        for f, _ in st._fields_:
            assert int(out[f"{n}.{f}"]) == getattr(st, f).offset, (n, f)
