"""Can the environments that will run into max_iter be predicted BEFORE the solve (to start them first)?  CPU study with the oracle.

    python tools/tail_predictor.py            (about one minute on 8 threads)

The per-step launch of B = 4096 ant environments ends with a tail: ~0.5 % of the environments need all 50 Newton iterations with ten
line-search trials each (5.3 ms on their two warps against 0.87 ms for the median environment), and whichever of them is dequeued
late finishes alone (profiles/README.md).  A longest-first order needs a predictor.  This script rolls the benchmark batch with the CPU
oracle (same seeds as bench.py), then scores, for steps 22..29, how well quantities known at the START of a step separate the
environments that will take >= 30 iterations: AUC per feature, and recall of the long environments inside the first wave (592 of 4096
= 14.5 % of the queue) for single features and for a gradient-boosted classifier trained on the first half of the steps.

Result (round 2): the previous step's iteration count has AUC 0.54 (1 - 3 of ~35 long environments were long one step earlier);
the best physical features are "nearly at rest" ones -- kinetic energy, largest contact normal velocity: a sticking contact with zero
tangential velocity is the degenerate point of the friction cone -- at AUC ~0.73, recall inside the first wave 0.26 - 0.35 (chance:
0.145); the classifier reaches 0.33 on held-out steps.  Nobody can start the long environments FIRST with that.  But the order only has
to keep them from starting LAST: the second part of the script list-schedules the measured iteration counts on 592 slots -- index /
random / previous-iterations order end at 10.3 - 10.4 (arbitrary units, ideal 6.3, longest environment 4.9), ascending kinetic energy
(the environments least likely to stall last) at 8.8, an oracle longest-first order at 6.8.  That key is what dojo_risk_key_kernel
computes (dojo.jl_b200/csrc/dojo_b200.cu); measured on B200: ant B = 4096 per-step 9.2 -> 8.1 ms.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from oracle import oracle as orc


def qrot(q, v):
    s, u = q[:, 0:1], q[:, 1:4]
    v = np.broadcast_to(v, (q.shape[0], 3))
    return v + 2 * np.cross(u, np.cross(u, v) + s * v)


def features(mech, z, u):
    B = z.shape[0]
    zz = z.reshape(B, mech.Nb, 13)
    phis, vns, vts = [], [], []
    for c in mech.contacts:
        x, q, v, w = zz[:, c.body, 0:3], zz[:, c.body, 6:10], zz[:, c.body, 3:6], zz[:, c.body, 10:13]
        ow = qrot(q, np.asarray(c.origin))
        n = np.asarray(c.normal)
        phis.append((x + ow - np.asarray(c.offset)) @ n - c.radius)
        vc = v + np.cross(qrot(q, w), ow - c.radius * n)
        vns.append(vc @ n)
        vts.append(np.linalg.norm(vc - np.outer(vc @ n, n), axis=1))
    phis, vns, vts = np.stack(phis, 1), np.stack(vns, 1), np.stack(vts, 1)
    return {"contacts_active": (phis < 1e-3).sum(1), "min_clearance": phis.min(1), "max_tangential_speed_active": np.where(phis < 1e-3, vts, 0).max(1),
            "max_abs_normal_velocity": np.abs(vns).max(1), "kinetic": (zz[:, :, 3:6] ** 2).sum((1, 2)) + (zz[:, :, 10:13] ** 2).sum((1, 2)),
            "max_angular_velocity": np.abs(zz[:, :, 10:13]).max((1, 2)), "max_input": np.abs(u).max(1)}


def auc(x, y):
    o = np.argsort(x)
    r = np.empty(len(x))
    r[o] = np.arange(len(x))
    n1, n0 = y.sum(), len(y) - y.sum()
    return (r[y].sum() - n1 * (n1 - 1) / 2) / (n1 * n0)


def recall_first_wave(score, y, B, frac):
    rs = []
    for s in range(0, len(y), B):
        top = np.argsort(-score[s:s + B])[:int(frac * B)]
        rs.append(y[s:s + B][top].sum() / max(1, y[s:s + B].sum()))
    return float(np.mean(rs))


def main():
    mech, B, T = dj.get_mechanism("ant"), 4096, 30
    opts = capi.solver_options()
    Z, rng = bench.synthetic_batch(mech, B, 0xD0D0 + 1, "ant")
    U = bench.random_inputs(mech, rng, T, B, 1.0)
    threads = bench.host_threads()[0]
    rec, t0 = [], time.time()
    for t in range(T):
        Zn, st, it = orc.step_batch_threads(mech, Z, U[t], opts, threads)
        rec.append((Z, U[t], st, it))
        Z = Zn
    print(f"rolled {T} steps of {B} environments on {threads} threads in {time.time() - t0:.0f} s; failed per step (last 8): {[int((r[2] != 0).sum()) for r in rec[-8:]]}")
    F, ys = {}, []
    for t in range(22, T):
        f = features(mech, rec[t][0], rec[t][1])
        f["previous_iterations"] = rec[t - 1][3].astype(float)
        for k, v in f.items():
            F.setdefault(k, []).append(np.asarray(v, float))
        ys.append(rec[t][3] >= 30)
    y = np.concatenate(ys)
    frac = 592 / B
    print(f"long environments (>= 30 iterations): {y.mean():.4f} of the batch")
    for k in F:
        v = np.concatenate(F[k])
        a = auc(v, y)
        print(f"  {k:30s} AUC {max(a, 1 - a):.3f} ({'low' if a < 0.5 else 'high'} first)   recall inside the first wave {recall_first_wave(v if a >= 0.5 else -v, y, B, frac):.3f}")
    # list scheduling of the measured work on 592 slots under different orders of the queue (duration ~ iterations; stalled
    # environments a little more for their extra line-search passes)
    import heapq

    def makespan(dur, order, m=592):
        h = [0.0] * m
        end = 0.0
        for e in order:
            t = heapq.heappop(h) + dur[e]
            end = max(end, t)
            heapq.heappush(h, t)
        return end
    sims = {}
    for si, t in enumerate(range(22, T)):
        dur = 0.087 * rec[t][3].astype(float)
        dur[rec[t][2] != 0] *= 1.12
        kin = F["kinetic"][si]
        key = np.clip(np.floor(2 * np.log2(np.maximum(kin, 1e-300))) + 88, 0, 127)
        orders = {"index order": np.arange(B), "random": np.random.default_rng(t).permutation(B), "previous iterations, descending": np.argsort(-rec[t - 1][3], kind="stable"),
                  "kinetic energy, ascending (half-octave buckets)": np.argsort(key, kind="stable"),
                  "energy key minus previous iterations / 2 (dojo_risk_key_kernel)": np.argsort(2 * key - rec[t - 1][3], kind="stable"),
                  "oracle: longest first": np.argsort(-dur, kind="stable")}
        for k, o in orders.items():
            sims.setdefault(k, []).append(makespan(dur, o))
        sims.setdefault("(ideal: total work / 592)", []).append(dur.sum() / 592)
    print("simulated makespan of one step [ms], mean over steps 22..29:")
    for k, v in sims.items():
        print(f"  {k:66s} {np.mean(v):6.2f}")
    try:
        from sklearn.ensemble import GradientBoostingClassifier
        X = np.stack([np.concatenate(F[k]) for k in F], 1)
        X = np.concatenate([X, np.log10(np.abs(X) + 1e-12)], 1)
        tr = np.arange(len(y)) < 4 * B
        gb = GradientBoostingClassifier(n_estimators=150, max_depth=3, subsample=0.8, random_state=0).fit(X[tr], y[tr])
        p = gb.predict_proba(X)[:, 1]
        print(f"  gradient-boosted classifier: recall inside the first wave, training steps {recall_first_wave(p[tr], y[tr], B, frac):.3f}, held-out steps {recall_first_wave(p[~tr], y[~tr], B, frac):.3f} (chance {frac:.3f})")
    except ImportError:
        pass


if __name__ == "__main__":
    main()
