"""GPU tests of the ImpactContact / LinearContact kernels through the C-ABI: the DJ_ANY_CONTACT compilation of the
step / gradient kernels (dojo_b200_cm.cu), selected by dojo_create for mechanisms that contain such contacts, against the oracle.
The same comparison runs on the CPU through the kernel emulation (tests/test_contact_models.py)."""
import numpy as np
import pytest

import dojo_jl_b200 as dj

from test_contact_models import _thrown

pytestmark = pytest.mark.gpu

CASES = [("sphere", "impact"), ("sphere", "linear"), ("block", "impact"), ("block", "linear")]


@pytest.mark.parametrize("name,ct", CASES)
def test_step_and_gradient_parity(name, ct):
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    mech = dj.get_mechanism(name, contact_type=ct)
    rng = np.random.default_rng(17)
    B, T = 64, 30
    stepper, o = BatchedStepper(mech, B), Oracle(mech)
    Z, U = _thrown(mech, B, rng, 0, None)
    same = total = status_diff = conv = 0
    for t in range(T):
        Zn, st, it, sol = stepper.step(Z, U, return_sol=True)
        for e in range(0, B, 4):
            zo, so, io, solo = o.step(Z[e], U[e], return_sol=True)
            total += 1
            if st[e] != so:      # the block on its four lower corners with the friction pyramid is a degenerate program: single steps
                status_diff += 1  # end at max_iter (:failed) in the reference too, and a rounding difference can decide which
                continue
            if so != 0:
                continue
            conv += 1
            if it[e] != io:
                assert np.abs(Zn[e] - zo).max() < 2e-2  # a rounding-level flip of a line-search comparison: solver tolerance (test_gpu_parity.TOL_SOLVER)
                continue
            same += 1
            assert np.abs(Zn[e] - zo).max() < 1e-6 and np.abs(sol[e] - solo).max() < 1e-5
        Z = Zn
    assert status_diff <= 0.03 * total and conv >= 0.7 * total and same >= 0.9 * conv, (same, conv, status_diff, total)
    assert (Z[:, 2] > (0.5 if name == "sphere" else 0.25) - 1e-4).all()  # resting on the ground, no penetration
    # fused rollout == step by step (bit-identical)
    Zf, _ = stepper.rollout(Z, np.tile(U, (5, 1, 1)), T=5)
    Zs = Z
    for _ in range(5):
        Zs, _, _ = stepper.step(Zs, U)
    assert np.array_equal(Zf, Zs)
    # IFT gradients
    Zn, Fz, Fu, st, it = stepper.step_grad(Z, U)
    errs = []
    for e in range(0, B, 4):
        zo, Fzo, Fuo, so, io = o.step_grad(Z[e], U[e])
        if so != 0 or st[e] != 0 or io != it[e]:
            continue
        errs.append(max(np.abs(Fz[e] - Fzo).max() / max(1.0, np.abs(Fzo).max()), np.abs(Fu[e] - Fuo).max() / max(1.0, np.abs(Fuo).max())))
    errs = np.array(errs)
    assert len(errs) >= 8 and np.median(errs) < 1e-7 and np.quantile(errs, 0.9) < 1e-4 and errs.max() < 1e-2, errs


def test_contact_data_gradients_are_nonlinear_only():
    """the reference defines the contact-data blocks for NonlinearContact only (gradients/data.jl:152, :173)"""
    from dojo_jl_b200.solver import BatchedStepper
    mech = dj.get_mechanism("sphere", contact_type="linear")
    stepper = BatchedStepper(mech, 4)
    with pytest.raises(RuntimeError, match="NonlinearContact"):
        stepper.step_grad_contact(np.tile(mech.z0, (4, 1)))


def test_nonlinear_mechanisms_keep_their_kernels():
    """sphere / block with the default NonlinearContact run on the first compilation (the benchmarked kernels) and agree with
    the oracle like the BASELINE models do"""
    from dojo_jl_b200.solver import BatchedStepper
    from oracle.oracle import Oracle
    mech = dj.get_mechanism("block")
    rng = np.random.default_rng(19)
    B = 32
    stepper, o = BatchedStepper(mech, B), Oracle(mech)
    Z, U = _thrown(mech, B, rng, 0, None)
    for t in range(20):
        Zn, st, it = stepper.step(Z, U)
        for e in range(0, B, 8):
            zo, so, io = o.step(Z[e], U[e])
            assert st[e] == so
            if it[e] == io:
                assert np.abs(Zn[e] - zo).max() < 1e-6
        Z = Zn
