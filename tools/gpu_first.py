import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
import dojo_jl_b200 as d
from dojo_jl_b200 import capi
from dojo_jl_b200.solver import BatchedStepper
from oracle.oracle import Oracle
for name, T in (("pendulum", 5), ("ant", 30), ("quadruped", 40), ("atlas", 30)):
    m = d.get_mechanism(name)
    o = Oracle(m); s = BatchedStepper(m, 64)
    print(name, "arena bytes", s.shared_bytes_per_env)
    rng = np.random.default_rng(0)
    B = 8
    Z = np.tile(m.z0, (B,1)); 
    U = rng.uniform(-1,1,(B,m.nu)); 
    if m.joints[0].nimpulses==0: U[:, :6] = 0
    for k in range(T):
        Zo, so, io = o.step_batch(Z, U)
        Zg, sg, ig = s.step(Z, U)
        err = np.abs(Zo-Zg).max()
        if k % 5 == 0 or err > 1e-7 or (io!=ig).any():
            print(" step", k, "err", err, "iters oracle", io[:4], "gpu", ig[:4], "status", so[:4], sg[:4])
        if not np.isfinite(err) or err > 1e-3: break
        Z = Zo
