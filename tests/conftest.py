import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dojo_jl_b200  # noqa: E402,F401  (import shim for the dojo.jl_b200 package directory)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(1234)


def random_inputs(mech, B, rng, scale=1.0):
    """U(-scale, scale) inputs on the actuated joints, zeros on a floating base."""
    U = rng.uniform(-scale, scale, (B, mech.nu))
    off = 0
    for j in mech.joints:
        if j.nimpulses == 0:
            U[:, off:off + j.input_dimension] = 0.0
        off += j.input_dimension
    return U


def jittered_states(mech, B, rng, base_z=(0.0, 0.3), joint_jitter=0.2):
    """B initial maximal states: forward kinematics at jittered joint coordinates (inside the limits),
    random base height, zero velocities."""
    from dojo_jl_b200.mechanism import unpack_maximal_state
    base = mech.minimal_coordinates(mech.z0)
    Z = np.zeros((B, mech.nz))
    for e in range(B):
        coords = {}
        for j in mech.joints:
            c = np.array(base[j.name], dtype=float)
            if j.nimpulses == 0:  # floating base: [x y z, rotation vector]
                c[2] += rng.uniform(*base_z)
                c[3:6] += rng.normal(0.0, 0.1, 3)
            elif j.input_dimension > 0:
                c = c + rng.uniform(-joint_jitter, joint_jitter, c.shape)
                for el, sl in ((j.tra, slice(0, j.tra.nfree)), (j.rot, slice(j.tra.nfree, None))):
                    if el.nlimits:
                        lo, hi = el.limit_lo, el.limit_hi
                        c[sl] = np.clip(c[sl], lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo))
            coords[j.name] = c
        Z[e] = mech.forward_kinematics(coords)
    return Z
