"""Host-side mirror of the reference interface for the hot path (same names, argument meaning and error behaviour;
Python has no `!`, so `step!` is `step`):

    reference (Julia)                                              here
    -------------------------------------------------------------  ------------------------------------------
    SolverOptions{T}(; rtol, btol, ...)   solver/options.jl:16-26   SolverOptions(rtol=..., btol=..., ...)
    step!(mechanism, z, u; opts)          simulation/step.jl:11     step(mechanism, z, u, opts=None)
    simulate!(mechanism, steps, storage, control!; opts)            simulate(mechanism, steps, control=None, record=False, opts=None)
                                          simulation/simulate.jl:16
    get_maximal_gradients!(mechanism, z, u; opts)                   get_maximal_gradients(mechanism, z, u, opts=None)
                                          gradients/state.jl:69
    get_contact_gradients(mechanism)      gradients/contact.jl:1    get_contact_gradients(mechanism, z, u, opts=None)
    mehrotra!(mechanism; opts) -> :success / :failed                status codes returned next to the states (STATUS)
    minimal_to_maximal(mechanism, x)      mechanism/state.jl:9      minimal_to_maximal(mechanism, x)
    maximal_to_minimal(mechanism, z)      mechanism/state.jl:44     maximal_to_minimal(mechanism, z)
    step_minimal_coordinates!(mechanism, x, u; opts)                step_minimal_coordinates(mechanism, x, u, opts=None)
                                          simulation/step.jl:42
    maximal_to_minimal_jacobian(mechanism, z)  gradients/state.jl:9   maximal_to_minimal_jacobian(mechanism, z)
    minimal_to_maximal_jacobian(mechanism, x)  gradients/state.jl:136 minimal_to_maximal_jacobian(mechanism, x)
    get_minimal_gradients!(mechanism, x, u; opts)                   get_minimal_gradients(mechanism, x, u, opts=None)
                                          gradients/state.jl:182
    simulate!(...; record=true) -> Storage  simulation/simulate.jl:16, storage.jl:15-67      simulate_record(mechanism, steps, ...) -> Storage
    momentum / kinetic_energy / potential_energy / mechanical_energy(mechanism, storage)   same names (mechanics/{momentum,energy}.jl)

NEW relative to the reference: every function also accepts a batch -- z of shape [B, 13 Nb], u of shape [B, nu] --
and then returns batched results.  All compute happens in libdojo_b200.so on the GPU (solver.BatchedStepper).

Deviations (documented switches):
  * step returns the TRUE next state (x3, v25, q3, w25) -- the mechanism's internal state after step! -- unless
    literal_q1=True (the reference's return value advances the configuration twice, Q1);
  * gradients are the consistent IFT gradients at the solution unless literal_q2=True (get_maximal_gradients! builds the data
    Jacobian AFTER update_state! but solves against the KKT matrix assembled before it, Q2);
  * "Excessive angular velocity" (solver/line_search.jl:18-20): status code 2 is reserved for it, but the reference's test can
    never fire -- candidate_step! clips |w|^2 to (3.9/h^2)^2 / |w|^2 < 3.9/h^2 before line_search! compares it with 3.91/h^2
     -- so no path produces it; _check_single keeps the mapping to the reference's error().
"""
from typing import Callable, Optional

import numpy as np

from . import capi
from .mechanism import Mechanism
from .solver import DOJO_FLAG_Q1_LITERAL_RETURN, DOJO_FLAG_Q2_LITERAL_GRADIENTS, STATUS, BatchedStepper

SolverOptions = capi.solver_options

_steppers = {}


def _stepper(mech: Mechanism, B: int, device: int = 0) -> BatchedStepper:
    key = (id(mech), device)
    s = _steppers.get(key)
    if s is None or s.max_batch < B:
        if s is not None:
            s.close()
        s = BatchedStepper(mech, max(B, 64), device)
        _steppers[key] = s
    return s


def _check_single(status):
    if int(status[0]) == 2:
        raise RuntimeError("Excessive angular velocity.")  # reference: error(...) in line_search!


def step(mechanism: Mechanism, z, u, opts=None, literal_q1: bool = False, device: int = 0):
    """step!(mechanism, z, u; opts).  z: [13Nb] or [B, 13Nb]; u: [nu] or [B, nu].  Returns z_next (same shape); for a batch
    also (status, iters)."""
    z = np.asarray(z, dtype=float)
    single = z.ndim == 1
    Z = np.atleast_2d(z)
    U = np.atleast_2d(np.asarray(u, dtype=float))
    s = _stepper(mechanism, Z.shape[0], device)
    Zn, status, iters = s.step(Z, U, opts, flags=DOJO_FLAG_Q1_LITERAL_RETURN if literal_q1 else 0)
    if single:
        _check_single(status)
        return Zn[0]
    return Zn, status, iters


def simulate(mechanism: Mechanism, steps: int, z0=None, control: Optional[Callable] = None, record: bool = False, opts=None, device: int = 0):
    """simulate!(mechanism, steps, storage, control!): `control(k)` returns the input(s) of step k ([nu] or [B, nu]; None = 0).
    Returns the final state(s) and, with record=True, the trajectory [steps, B, 13Nb] (the reference's Storage)."""
    z0 = mechanism.z0 if z0 is None else z0
    z0 = np.asarray(z0, dtype=float)
    single = z0.ndim == 1
    Z = np.atleast_2d(z0)
    B = Z.shape[0]
    s = _stepper(mechanism, B, device)
    U = None
    if control is not None:
        U = np.zeros((steps, B, mechanism.nu))
        for k in range(steps):
            uk = control(k)
            if uk is not None:
                U[k] = np.asarray(uk, dtype=float)
    out = s.rollout(Z, U, steps, opts, record=record)
    Zf, traj = out[0], (out[2] if record else None)
    if single:
        return (Zf[0], traj[:, 0]) if record else Zf[0]
    return (Zf, traj) if record else Zf


def get_maximal_gradients(mechanism: Mechanism, z, u, opts=None, device: int = 0, literal_q2: bool = False):
    """get_maximal_gradients!(mechanism, z, u; opts) -> (jacobian_state [12Nb x 12Nb], jacobian_control [12Nb x nu]);
    batched inputs give [B, 12Nb, 12Nb] and [B, 12Nb, nu].  literal_q2=True reproduces the reference's literal result (data
    Jacobian and chain rule at the state shifted by update_state!, gradients/state.jl:69-76); the default is the consistent
    implicit-function-theorem gradient."""
    z = np.asarray(z, dtype=float)
    single = z.ndim == 1
    Z = np.atleast_2d(z)
    U = np.atleast_2d(np.asarray(u, dtype=float))
    s = _stepper(mechanism, Z.shape[0], device)
    _, Fz, Fu, status, _ = s.step_grad(Z, U, opts, flags=DOJO_FLAG_Q2_LITERAL_GRADIENTS if literal_q2 else 0)
    if single:
        _check_single(status)
        return Fz[0], Fu[0]
    return Fz, Fu


def get_contact_gradients(mechanism: Mechanism, z, u, opts=None, device: int = 0):
    """get_contact_gradients!(mechanism, z, u; opts) (gradients/contact.jl:1-55, the step is taken first as in
    get_maximal_gradients!) -> (jacobian_state [12Nb x 12Nb], jacobian_contact [12Nb x 5Ni]); per contact the data are
    [friction_coefficient, contact_radius, contact_origin(3)].  Batched inputs give [B, ...]."""
    z = np.asarray(z, dtype=float)
    single = z.ndim == 1
    Z = np.atleast_2d(z)
    U = np.atleast_2d(np.asarray(u, dtype=float))
    _, Fz, _, Fc, status, _ = _stepper(mechanism, Z.shape[0], device).step_grad_contact(Z, U, opts)
    if single:
        _check_single(status)
        return Fz[0], Fc[0]
    return Fz, Fc


def minimal_to_maximal(mechanism: Mechanism, x, device: int = 0):
    """minimal_to_maximal(mechanism, x): x [2 nu] or [B, 2 nu] (per joint [c_tra; c_rot; v_tra; v_rot]) -> z [13 Nb] / [B, 13 Nb]."""
    x = np.asarray(x, dtype=float)
    X = np.atleast_2d(x)
    Z = _stepper(mechanism, X.shape[0], device).minimal_to_maximal(X)
    return Z[0] if x.ndim == 1 else Z


def maximal_to_minimal(mechanism: Mechanism, z, device: int = 0):
    """maximal_to_minimal(mechanism, z): z [13 Nb] or [B, 13 Nb] -> x [2 nu] / [B, 2 nu]."""
    z = np.asarray(z, dtype=float)
    Z = np.atleast_2d(z)
    X = _stepper(mechanism, Z.shape[0], device).maximal_to_minimal(Z)
    return X[0] if z.ndim == 1 else X


def step_minimal_coordinates(mechanism: Mechanism, x, u, opts=None, device: int = 0, literal: bool = False):
    """step_minimal_coordinates!(mechanism, x, u; opts) -> x_next (what DojoEnvironments.step! calls); for a batch also
    (status, iters).  The maximal states stay on the device between the three launches.  literal = True returns what the reference
    literally returns (step!'s return value advances the configuration a second time) instead of the state after the step."""
    x = np.asarray(x, dtype=float)
    single = x.ndim == 1
    X = np.atleast_2d(x)
    U = np.atleast_2d(np.asarray(u, dtype=float))
    Xn, status, iters = _stepper(mechanism, X.shape[0], device).step_minimal(X, U, opts, flags=1 if literal else 0)
    if single:
        _check_single(status)
        return Xn[0]
    return Xn, status, iters


def maximal_to_minimal_jacobian(mechanism: Mechanism, z, device: int = 0):
    """maximal_to_minimal_jacobian(mechanism, z) -> [2 nu x 12 Nb] (columns: attitude-reduced [x, v, phi, w] per body);
    batched z gives [B, 2 nu, 12 Nb]."""
    z = np.asarray(z, dtype=float)
    Z = np.atleast_2d(z)
    J = _stepper(mechanism, Z.shape[0], device).maximal_to_minimal_jacobian(Z)
    return J[0] if z.ndim == 1 else J


def minimal_to_maximal_jacobian(mechanism: Mechanism, x, device: int = 0):
    """minimal_to_maximal_jacobian(mechanism, x) -> [12 Nb x 2 nu], the derivative of minimal_to_maximal at x (the
    reference reads the mechanism's stored state, which its callers set to minimal_to_maximal(x) first; it chains the
    partials in mechanism.bodies order, identical to the root -> leaves order used here when parents precede children)."""
    x = np.asarray(x, dtype=float)
    X = np.atleast_2d(x)
    s = _stepper(mechanism, X.shape[0], device)
    J = s.minimal_to_maximal_jacobian(s.minimal_to_maximal(X))
    return J[0] if x.ndim == 1 else J


def get_minimal_gradients(mechanism: Mechanism, x, u, opts=None, device: int = 0):
    """get_minimal_gradients!(mechanism, x, u; opts) -> (minimal_jacobian_state [2nu x 2nu], minimal_jacobian_control
    [2nu x nu]); batched inputs give [B, 2nu, 2nu] and [B, 2nu, nu].  Consistent variant: the map Jacobians are
    taken at z = minimal_to_maximal(x) and at the true next state."""
    x = np.asarray(x, dtype=float)
    single = x.ndim == 1
    X = np.atleast_2d(x)
    U = np.atleast_2d(np.asarray(u, dtype=float))
    _, Gx, Gu, status, _ = _stepper(mechanism, X.shape[0], device).minimal_gradients(X, U, opts)
    if single:
        _check_single(status)
        return Gx[0], Gu[0]
    return Gx, Gu


class Storage:
    """Storage{T,N} (simulation/storage.jl:15-42) for a batch: arrays indexed [step, environment, body, :].
    x, q, v, ω: the state before every solve; px, pq: body momenta (world frame); vl, ωl: velocities derived from them."""

    def __init__(self, traj, sto, diag):
        T, B, _ = traj.shape
        z = traj.reshape(T, B, -1, 13)
        self.x, self.v, self.q, self.ω = z[..., 0:3], z[..., 3:6], z[..., 6:10], z[..., 10:13]
        self.px, self.pq, self.vl, self.ωl = sto[..., 0:3], sto[..., 3:6], sto[..., 6:9], sto[..., 9:12]
        self._diag = diag

    def __len__(self):
        return self.x.shape[0]


def momentum(mechanism: Mechanism, storage: Storage):
    """momentum(mechanism, storage) (mechanics/momentum.jl:1-15): [steps, B, 6] = linear; angular about the centre of mass"""
    return storage._diag[..., 0:6]


def kinetic_energy(mechanism: Mechanism, storage: Storage):
    """kinetic_energy(mechanism, storage) (mechanics/energy.jl:17-41): [steps, B]"""
    return storage._diag[..., 6]


def potential_energy(mechanism: Mechanism, storage: Storage):
    """potential_energy(mechanism, storage) (mechanics/energy.jl:43-93): [steps, B]"""
    return storage._diag[..., 7]


def mechanical_energy(mechanism: Mechanism, storage: Storage):
    """mechanical_energy(mechanism, storage) (mechanics/energy.jl:1-15)"""
    return storage._diag[..., 6] + storage._diag[..., 7]


def simulate_record(mechanism: Mechanism, steps: int, z0=None, control: Optional[Callable] = None, opts=None, device: int = 0) -> Storage:
    """simulate!(mechanism, steps, storage, control!; record=true) -> Storage, everything computed on the device
    (momenta and energies included: save_to_storage!, simulation/storage.jl:50-67)."""
    z0 = mechanism.z0 if z0 is None else z0
    Z = np.atleast_2d(np.asarray(z0, dtype=float))
    B = Z.shape[0]
    s = _stepper(mechanism, B, device)
    U = None
    if control is not None:
        U = np.zeros((steps, B, mechanism.nu))
        for k in range(steps):
            uk = control(k)
            if uk is not None:
                U[k] = np.asarray(uk, dtype=float)
    _, traj, sto, diag, _ = s.simulate_record(Z, U, steps, opts)
    return Storage(traj, sto, diag)


def status_name(code: int) -> str:
    return STATUS.get(int(code), "unknown")
