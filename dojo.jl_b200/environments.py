"""Host-side mirror of DojoEnvironments' environment interface on a batched axis.

    reference (Julia)                                                        here
    -----------------------------------------------------------------------  ----------------------------------------
    get_environment(:ant_ars; ...)          DojoEnvironments/src/environments.jl:36-40     get_environment("ant_ars", batch=B)
    state_map(env, s) / input_map(env, a)   environments/ant_ars.jl:53-61                  env.state_map(s) / env.input_map(a)
    step!(env, state, input; k, record)     environments.jl:77-84, ant_ars.jl:63-70        env.step(state, input)
    get_state(env)                          ant_ars.jl:72-79, quadruped_sampling.jl:68-73  env.get_state()
    initialize!(env, model)                 environments.jl:118-120                        env.initialize()
    reward / failure test of the examples   examples/learning/ant_ars.jl:79-116            returned by env.step: (reward, done)

The environment holds the state of its B mechanisms (the reference's Environment wraps ONE mutable Mechanism).  All compute
is in libdojo_b200.so (dojo_env_step: pre kernel, step kernel, post kernel); this file only keeps names and shapes.
"""
from typing import Optional

import numpy as np

from . import capi
from .mechanism import Mechanism, get_mechanism
from .solver import BatchedStepper


class Environment:
    """Batched Environment{T,N}: `mechanism` + B states.  Subclasses fix the spec (state / input maps, reward, failure test)."""
    mechanism_name = ""
    spec_kwargs = {}

    def __init__(self, batch: int = 1, horizon: int = 100, device: int = 0, mechanism: Optional[Mechanism] = None, **mechanism_kwargs):
        self.mechanism = mechanism if mechanism is not None else get_mechanism(self.mechanism_name, **mechanism_kwargs)
        self.batch, self.horizon = int(batch), int(horizon)
        self.spec = capi.env_spec(**self.spec_kwargs)
        self.stepper = BatchedStepper(self.mechanism, self.batch, device)
        self.ns, self.na = self.stepper.env_sizes(self.spec)
        self.state = np.zeros((self.batch, self.ns))
        self.status = np.zeros(self.batch, dtype=np.int32)
        self.initialize()

    # ---- maps (pure index manipulation, mirrored for callers that want them; the device applies them inside dojo_env_step)
    def state_map(self, state):
        return np.asarray(state)[..., :2 * self.mechanism.nu]

    def input_map(self, action):
        a = np.asarray(action, dtype=float)
        return np.concatenate([np.zeros(a.shape[:-1] + (self.spec.n_unactuated,)), a], axis=-1)

    # ---- initialize!(environment, model): the mechanism's initial pose (initialize_<model>! of the builder)
    def initial_state(self) -> np.ndarray:
        x0 = self.stepper.maximal_to_minimal(self.mechanism.z0[None])[0]
        return np.concatenate([x0, np.zeros(self.ns - x0.size)])

    def initialize(self, mask=None):
        self.stepper.env_reset(self.spec, self.state, self.initial_state(), mask)
        return self.state

    def get_state(self):
        return self.state.copy()

    def step(self, state=None, action=None, opts=None):
        """step!(environment, state, input): advances every environment from `state` (default: the held state) and returns
        (reward [B], done [B]); the new state is get_state()."""
        S = self.state if state is None else np.atleast_2d(np.asarray(state, dtype=float))
        Sn, reward, done, status, _ = self.stepper.env_step(self.spec, S, action, opts)
        self.state, self.status = Sn, status
        return reward, done


    def rollout(self, actions, state=None, opts=None):
        """Open-loop rollout from `state` (default: the held state) with actions [T, B, na]: returns (return [B], failed [B]);
        the held state becomes the final state.  One C call, 3 T launches, nothing crosses the PCIe bus in between."""
        A = np.asarray(actions, dtype=float)
        S = self.state if state is None else np.atleast_2d(np.asarray(state, dtype=float))
        self.state, ret, failed = self.stepper.env_rollout(self.spec, S, A, A.shape[0], opts)
        return ret, failed


    def policy_rollout(self, theta, horizon: int, mean=None, std=None, state=None, opts=None, record_states: bool = False):
        """rollout_policy(θ, env, normalizer, parameters) of examples/learning/ant_ars.jl:79-116 for B policies at once:
        theta [B, na, ns] (one perturbed policy per environment), action = theta_e * normalize(state) with the normaliser
        frozen for the call.  Returns (return [B], failed [B]) (+ the observed states [horizon, B, ns] for observe!)."""
        S = self.state if state is None else np.atleast_2d(np.asarray(state, dtype=float))
        out = self.stepper.env_policy_rollout(self.spec, S, theta, horizon, mean, std, opts, record_states)
        self.state = out[0]
        return out[1:]


class AntARS(Environment):
    """environments/ant_ars.jl: state [minimal state (28); clamped normal contact impulses (9)], 8 actions, the reward of
    examples/learning/ant_ars.jl:98-107 and its failure test :112."""
    mechanism_name = "ant"
    spec_kwargs = dict(n_unactuated=6, contact_obs=True, forward_index=0, healthy_index=2, w_forward=100.0, w_control=0.05 / 10,
                       w_contact=0.5 * 1.0e-3, survive_reward=0.05, healthy_min=0.2, healthy_max=1.0)


class QuadrupedSampling(Environment):
    """environments/quadruped_sampling.jl: state = minimal state (36), 12 actions; failure test of
    examples/learning/quadruped_sampling.jl:69 (x[3] < 0 || !isfinite || |x[1]| > 1000)."""
    mechanism_name = "quadruped"
    spec_kwargs = dict(n_unactuated=6, contact_obs=False, healthy_index=2, healthy_min=0.0, bound_index=0, bound_abs=1000.0)


class QuadrupedWaypoint(QuadrupedSampling):
    """environments/quadruped_waypoint.jl: the quadruped_sampling maps on the builder's other defaults -- timestep 0.001 and
    contact_body = false (the four foot contacts only, quadruped_waypoint.jl:8, :25)."""

    def __init__(self, batch: int = 1, horizon: int = 100, device: int = 0, mechanism: Optional[Mechanism] = None, **mechanism_kwargs):
        if mechanism is None:
            mechanism_kwargs.setdefault("timestep", 0.001)
            mechanism = get_mechanism(self.mechanism_name, **mechanism_kwargs)
            mechanism.contacts = [c for c in mechanism.contacts if c.name.endswith("_calf_contact")]  # contact_feet only
        super().__init__(batch, horizon, device, mechanism)


class Pendulum(Environment):
    """environments/pendulum.jl: identity maps."""
    mechanism_name = "pendulum"
    spec_kwargs = dict()


class CartpoleDQN(Environment):
    """environments/cartpole_dqn.jl: state = minimal state [y, ydot, theta, thetadot], ONE action on the cart (input_map: [a; 0]);
    springs / dampers / joint_limits are the builder's options (cartpole_dqn.jl:5-33).  The unactuated input comes last here, which
    the fused dojo_env_step kernels (leading unactuated inputs) do not express: step! is the composition the reference itself
    performs, state_map -> input_map -> step_minimal_coordinates! (three launches); the DQN example's own reward / termination
    (examples/learning/cartpole_dqn.jl:157-186) stays with the caller."""
    mechanism_name = "cartpole"
    spec_kwargs = dict()

    def __init__(self, batch: int = 1, horizon: int = 100, device: int = 0, **mechanism_kwargs):
        super().__init__(batch, horizon, device, **mechanism_kwargs)
        self.na = 1

    def input_map(self, action):
        a = np.asarray(action, dtype=float).reshape(-1, 1)
        return np.concatenate([a, np.zeros_like(a)], axis=-1)

    def step(self, state=None, action=None, opts=None):
        S = self.state if state is None else np.atleast_2d(np.asarray(state, dtype=float))
        U = self.input_map(np.zeros(S.shape[0]) if action is None else action)
        self.state, self.status, _ = self.stepper.step_minimal(self.state_map(S), U, opts)
        done = (~np.isfinite(self.state).all(axis=1)).astype(np.int32)
        return np.zeros(S.shape[0]), done


_ENVIRONMENTS = {"ant_ars": AntARS, "quadruped_sampling": QuadrupedSampling, "quadruped_waypoint": QuadrupedWaypoint, "pendulum": Pendulum,
                 "cartpole_dqn": CartpoleDQN}


def get_environment(name: str, **kwargs) -> Environment:
    """get_environment(model; kwargs...)  (DojoEnvironments/src/environments.jl:36-40)"""
    try:
        return _ENVIRONMENTS[name](**kwargs)
    except KeyError:
        raise ValueError(f"unknown environment {name!r}; available: {sorted(_ENVIRONMENTS)}") from None
