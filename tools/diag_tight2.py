import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dojo_jl_b200 as dj
from dojo_jl_b200 import capi
from dojo_jl_b200.solver import BatchedStepper
from oracle.oracle import Oracle
from conftest import jittered_states, random_inputs
mech = dj.get_mechanism("ant"); rng = np.random.default_rng(11); B=48
opts = capi.solver_options(rtol=1e-9, btol=1e-9)
Z = jittered_states(mech, B, rng); st = BatchedStepper(mech, B); o = Oracle(mech, opts)
offs = mech.node_offsets()
for t in range(8):
    U = random_inputs(mech, B, rng, 1.0)
    Zg, sg, ig, solg = st.step(Z, U, opts=opts, return_sol=True)
    Zo = np.empty_like(Z)
    for e in range(B): Zo[e], _, _ = o.step(Z[e], U[e])
    for e in np.where(sg != 0)[0][:2]:
        for tol in (1e-6, 1e-7, 1e-8, 1e-9):
            _, s2, i2 = st.step(Z[e:e+1], U[e:e+1], opts=capi.solver_options(rtol=tol, btol=tol))
            print("  tol", tol, "gpu status", s2[0], "iters", i2[0])
        o.set_state(Z[e], U[e]); o.set_solution(solg[e], 0.0)
        r = o.evaluate_rhs(solg[e], 0.0)
        rv, bv = o.violations()
        k = np.argsort(-np.abs(r))[:6]
        def node(i):
            n = np.searchsorted(offs, i, side='right') - 1
            return ("J%d" % n if n < mech.Ne else ("B%d" % (n - mech.Ne) if n < mech.Ne + mech.Nb else "C%d" % (n - mech.Ne - mech.Nb)), int(i - offs[n]))
        print("step", t, "env", e, "oracle-evaluated rvio %.2e bvio %.2e" % (rv, bv), "largest rhs rows:", [(node(i), float("%.2e" % r[i])) for i in k])
        c0 = mech.contact_sol_offset(0)
        sc = solg[e][c0:].reshape(-1, 8)
        print("   contacts s1,g1,s2,g2:", np.array2string(sc[:, [0, 4, 1, 5]], precision=2, max_line_width=200).replace("\n", ""))
    Z = Zo
# dump the first stalled case for offline analysis
rng = np.random.default_rng(11); Z = jittered_states(mech, B, rng)
for t in range(8):
    U = random_inputs(mech, B, rng, 1.0)
    Zg, sg, ig, solg = st.step(Z, U, opts=opts, return_sol=True)
    bad = np.where(sg != 0)[0]
    if len(bad):
        e = bad[0]
        # also the GPU iterate after fewer iterations (where it should still be making progress)
        sols = {}
        for mi in (8, 10, 12, 14, 16, 20):
            _, s2, i2, so2 = st.step(Z[e:e+1], U[e:e+1], opts=capi.solver_options(rtol=1e-9, btol=1e-9, max_iter=mi), return_sol=True)
            sols[mi] = so2[0]
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/stall_case.npz", z=Z[e], u=U[e], sol=solg[e], **{f"sol_{k}": v for k, v in sols.items()})
        print("dumped env", e, "step", t)
        break
    Zo = np.empty_like(Z)
    for e in range(B): Zo[e], _, _ = o.step(Z[e], U[e])
    Z = Zo
