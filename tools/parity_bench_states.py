"""Parity on the BENCHMARKED states: the batch bench.py times (after its roll-in) is stepped once more on the GPU and a
random subset is compared with the CPU oracle: status, Newton-iteration count, contact-mode bitmap, z_next.

    python tools/parity_bench_states.py [mech] [B] [sample] [rollin]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi


def contact_modes(mech, sol):
    off = mech.contact_sol_offset(0) if mech.Ni else 0
    s = sol[:, off:].reshape(sol.shape[0], mech.Ni, 8)
    return s[:, :, 4] > s[:, :, 0]


def compare(mech, Z, U, stepper, sample, seed=0, opts=None):
    """Returns a dict of mismatch counts / errors for `sample` random environments of the batch (Z, U)."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(seed)
    B = Z.shape[0]
    idx = np.sort(rng.choice(B, size=min(sample, B), replace=False))
    Zg, sg, ig, solg = stepper.step(Z, U, opts=opts, return_sol=True)
    o = Oracle(mech, opts)
    n = len(idx)
    Zo, so, io, solo = np.empty((n, mech.nz)), np.zeros(n, np.int32), np.zeros(n, np.int32), np.empty((n, mech.nres))
    for k, e in enumerate(idx):
        Zo[k], so[k], io[k], solo[k] = o.step(Z[e], U[e], return_sol=True)
    Zg, sg, ig, solg = Zg[idx], sg[idx], ig[idx], solg[idx]
    conv = (so == 0) & (sg == 0)
    same = conv & (ig == io)
    err = np.abs(Zg - Zo).max(axis=1)
    modes = (contact_modes(mech, solg) != contact_modes(mech, solo)).any(axis=1) if mech.Ni else np.zeros(n, bool)
    return {"sample": int(n), "status_mismatch": int((sg != so).sum()), "failed_gpu": int((sg != 0).sum()), "failed_oracle": int((so != 0).sum()),
            "iters_mismatch": int((conv & (ig != io)).sum()), "contact_mode_mismatch_same_iters": int((modes & same).sum()),
            "contact_mode_mismatch_all_converged": int((modes & conv).sum()),
            "max_abs_dz_same_iters": float(err[same].max(initial=0.0)), "median_abs_dz_same_iters": float(np.median(err[same])) if same.any() else 0.0,
            "max_abs_dz_converged": float(err[conv].max(initial=0.0)), "mean_iters_gpu": float(ig.mean()), "mean_iters_oracle": float(io.mean())}


if __name__ == "__main__":
    import json
    from dojo_jl_b200.solver import BatchedStepper
    name = sys.argv[1] if len(sys.argv) > 1 else "ant"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    sample = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    rollin = int(sys.argv[4]) if len(sys.argv) > 4 else 23
    mech = dj.get_mechanism(name)
    Z, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1)
    U = bench.random_inputs(mech, rng, rollin + 4, B, bench.SCALE[name])
    st = BatchedStepper(mech, B)
    for t in range(rollin):
        Z, _, _ = st.step(Z, U[t])
    for t in range(rollin, rollin + 3):
        r = compare(mech, Z, U[t], st, sample, seed=t)
        print(json.dumps({"mech": name, "B": B, "step": t, **r}))
        Z, _, _ = st.step(Z, U[t])
