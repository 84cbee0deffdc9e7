"""CPU restatement of the DojoEnvironments layer around step_minimal_coordinates! -- TEST INFRASTRUCTURE (like oracle.py).

    state_map / input_map      DojoEnvironments/src/environments/ant_ars.jl:53-61, quadruped_sampling.jl:51-58, pendulum.jl:41-47
    step!(environment, s, a)   DojoEnvironments/src/environments.jl:77-84
    get_state                  ant_ars.jl:72-79 (minimal state + max(-1, min(1, contact.impulses[2][1])) per contact)
    reward, failure test       examples/learning/ant_ars.jl:98-112 ; examples/learning/quadruped_sampling.jl:69
"""
import numpy as np


def env_step(o, spec, s, a, opts=None):
    """One environment: returns (s_next, reward, done, status, iters)."""
    mech = o.mech
    nu, Ni = mech.nu, mech.Ni
    s = np.asarray(s, dtype=float)
    na = nu - spec.n_unactuated
    a = np.zeros(na) if a is None else np.asarray(a, dtype=float)
    x = s[:2 * nu]                                            # state_map
    u = np.concatenate([np.zeros(spec.n_unactuated), a])      # input_map
    zn, status, iters, sol = o.step(o.minimal_to_maximal(x), u, opts=opts, return_sol=True)   # step_minimal_coordinates!
    xn = o.maximal_to_minimal(zn)
    gam = np.array([max(-1.0, min(1.0, sol[mech.contact_sol_offset(c) + 4])) for c in range(Ni)])  # contact.impulses[2][1]
    sn = np.concatenate([xn, gam]) if spec.contact_obs else xn
    reward = spec.survive_reward - spec.w_control * float(a @ a) - spec.w_contact * float(gam @ gam)
    if spec.forward_index >= 0:
        reward += spec.w_forward * (sn[spec.forward_index] - s[spec.forward_index]) / mech.timestep
    ok = bool(np.all(np.isfinite(sn)))
    if spec.healthy_index >= 0:
        ok = ok and spec.healthy_min <= sn[spec.healthy_index] <= spec.healthy_max
    if spec.bound_index >= 0:
        ok = ok and abs(sn[spec.bound_index]) <= spec.bound_abs
    return sn, reward, int(not ok), status, iters
