"""Host-side mirror of the reference's ``Mechanism`` (src/mechanism/constructor.jl:19-35).

The Julia host builds a ``Mechanism`` (bodies, joints, contacts, timestep, gravity,
input_scaling); the device side only ever sees the *flattened* descriptor
(include/dojo_b200.h ``DojoMechanismDesc``).  This module holds that flattened
model in plain Python/numpy, loads/saves it as JSON (dojo.jl_b200/mechanisms/*.json,
generated from the reference's URDFs + builder tables by tools/build_mechanisms.py)
and provides the host helpers that sit either side of the hot path:

* node/solution ordering       -- mechanism/id.jl:5-13, gradients/finite_difference.jl:1-18
* ``zero_coordinates!`` / ``set_minimal_coordinates!`` (forward kinematics)
                               -- mechanism/set.jl:94-127, joints/minimal.jl:21-44,
                                  joints/rotational/minimal.jl:85-98, translational/minimal.jl:70-88
* maximal state packing        -- mechanism/get.jl:107-134, gradients/utilities.jl:36-42
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import quat as Q

MECHANISM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mechanisms")


@dataclass
class Body:
    name: str
    mass: float
    inertia: np.ndarray  # 3x3


@dataclass
class JointElement:
    """One half (translational or rotational) of a JointConstraint.

    nlambda = number of constrained axes (Nλ), nlimits = Nb½ (0, or 3 - nlambda).
    axis_mask rows are V1, V2, V3 of joints/orthogonal.jl.
    """
    nlambda: int
    axis_mask: np.ndarray  # 3x3 rows V1,V2,V3
    spring: float = 0.0
    damper: float = 0.0
    spring_offset: np.ndarray = field(default_factory=lambda: np.zeros(3))
    limit_lo: Optional[np.ndarray] = None
    limit_hi: Optional[np.ndarray] = None

    @property
    def nlimits(self) -> int:
        return 0 if self.limit_lo is None else len(self.limit_lo)

    @property
    def nfree(self) -> int:
        return 3 - self.nlambda

    @property
    def nimpulses(self) -> int:  # N = Nλ + 2 Nb, Nb = 2 Nb½
        return self.nlambda + 4 * self.nlimits

    def constraint_mask(self) -> np.ndarray:  # joints/joint.jl:56-59
        V1, V2, V3 = self.axis_mask
        return {0: np.zeros((0, 3)), 1: V3[None, :], 2: np.stack([V1, V2]), 3: np.eye(3)}[self.nlambda]

    def nullspace_mask(self) -> np.ndarray:  # joints/joint.jl:61-64
        V1, V2, V3 = self.axis_mask
        return {0: np.eye(3), 1: np.stack([V1, V2]), 2: V3[None, :], 3: np.zeros((0, 3))}[self.nlambda]


@dataclass
class Joint:
    name: str
    parent: int  # body index, -1 = origin
    child: int
    tra: JointElement
    rot: JointElement
    vertex_parent: np.ndarray = field(default_factory=lambda: np.zeros(3))
    vertex_child: np.ndarray = field(default_factory=lambda: np.zeros(3))
    orientation_offset: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))

    @property
    def nimpulses(self) -> int:
        return self.tra.nimpulses + self.rot.nimpulses

    @property
    def input_dimension(self) -> int:  # joints/constraints.jl:450-456
        return self.tra.nfree + self.rot.nfree

    @property
    def spring_flag(self) -> bool:
        return self.tra.spring != 0 or self.rot.spring != 0

    @property
    def damper_flag(self) -> bool:
        return self.tra.damper != 0 or self.rot.damper != 0


CONTACT_TYPES = {"impact": 0, "linear": 1, "nonlinear": 2}  # DojoContactDesc.type (contacts/constructor.jl:117-128)
CONTACT_DIMS = {"impact": 2, "linear": 12, "nonlinear": 8}   # N of Contact{T,N}: impact.jl:38, linear.jl:46, nonlinear.jl:47


@dataclass
class Contact:
    """ContactConstraint{model} + SphereHalfSpaceCollision (contacts/constructor.jl:14-43).
    model: "nonlinear" (contacts/nonlinear.jl:12-48), "linear" (contacts/linear.jl:10-47: 4-sided friction pyramid) or
    "impact" (contacts/impact.jl:8-39: no friction) -- the reference's `contact_type` keyword."""
    name: str
    body: int
    friction: float
    normal: np.ndarray  # 3
    tangent: np.ndarray  # 2x3
    origin: np.ndarray  # 3
    radius: float
    offset: np.ndarray = field(default_factory=lambda: np.zeros(3))
    model: str = "nonlinear"

    @property
    def type(self) -> int:
        return CONTACT_TYPES[self.model]

    @property
    def dim(self) -> int:
        return CONTACT_DIMS[self.model]


class Mechanism:
    def __init__(self, name: str, bodies: List[Body], joints: List[Joint], contacts: List[Contact],
                 timestep: float = 0.01, input_scaling: Optional[float] = None,
                 gravity: Sequence[float] = (0.0, 0.0, -9.81)):
        self.name = name
        self.bodies = bodies
        self.joints = joints
        self.contacts = contacts
        self.timestep = float(timestep)
        self.input_scaling = float(timestep if input_scaling is None else input_scaling)
        self.gravity = np.asarray(gravity, dtype=float)
        self.z0: Optional[np.ndarray] = None  # initial maximal state (from initialize_*!)

    # ------------------------------------------------------------------ sizes
    @property
    def Nb(self) -> int:
        return len(self.bodies)

    @property
    def Ne(self) -> int:
        return len(self.joints)

    @property
    def Ni(self) -> int:
        return len(self.contacts)

    @property
    def nz(self) -> int:
        return 13 * self.Nb

    @property
    def nu(self) -> int:
        return sum(j.input_dimension for j in self.joints)

    @property
    def node_dims(self) -> List[int]:
        """joints (by id) | bodies | contacts  (mechanism/id.jl:5-13)."""
        return [j.nimpulses for j in self.joints] + [6] * self.Nb + [c.dim for c in self.contacts]

    @property
    def nres(self) -> int:
        return sum(self.node_dims)

    def node_offsets(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.node_dims)]).astype(int)

    def body_sol_offset(self, b: int) -> int:
        return int(self.node_offsets()[self.Ne + b])

    def joint_sol_offset(self, j: int) -> int:
        return int(self.node_offsets()[j])

    def contact_sol_offset(self, c: int) -> int:
        return int(self.node_offsets()[self.Ne + self.Nb + c])

    def input_offsets(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum([j.input_dimension for j in self.joints])]).astype(int)

    def joint_by_name(self, name: str) -> Joint:
        for j in self.joints:
            if j.name == name:
                return j
        raise KeyError(name)

    def body_index(self, name: str) -> int:
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)

    def parent_joint_of(self, b: int) -> Optional[int]:
        for ji, j in enumerate(self.joints):
            if j.child == b:
                return ji
        return None

    # ----------------------------------------------------------- kinematics
    def root_to_leaves_joints(self) -> List[int]:
        """Tree order of joints, parents before children (mechanism/traversal.jl)."""
        order, placed = [], {-1}
        remaining = list(range(self.Ne))
        while remaining:
            progressed = False
            for ji in list(remaining):
                if self.joints[ji].parent in placed:
                    order.append(ji)
                    placed.add(self.joints[ji].child)
                    remaining.remove(ji)
                    progressed = True
            if not progressed:
                raise ValueError("joint graph is not a tree rooted at the origin")
        return order

    def forward_kinematics(self, coords: Optional[Dict[str, Sequence[float]]] = None) -> np.ndarray:
        """zero_velocities! + zero_coordinates! + set_minimal_coordinates!(joint, xθ) for the
        given joints (others 0).  Returns the maximal state z (13 Nb) with zero velocities.

        xθ layout per joint = [Δx (tra free axes); Δθ (rot free axes)] (joints/minimal.jl:21-44).
        q_b = q_a * q_off * axis_angle_to_quaternion(Aᵀθ)       (rotational/minimal.jl:85-98)
        x_b = x_a + rot(p_a + AᵀΔx, q_a) - rot(p_b, q_b)        (translational/minimal.jl:70-88)
        """
        coords = coords or {}
        x = np.zeros((self.Nb, 3))
        q = np.tile(np.array([1.0, 0, 0, 0]), (self.Nb, 1))
        for ji in self.root_to_leaves_joints():
            j = self.joints[ji]
            c = np.asarray(coords.get(j.name, np.zeros(j.input_dimension)), dtype=float)
            dx, dth = c[: j.tra.nfree], c[j.tra.nfree:]
            xa = np.zeros(3) if j.parent < 0 else x[j.parent]
            qa = np.array([1.0, 0, 0, 0]) if j.parent < 0 else q[j.parent]
            At = j.tra.nullspace_mask().T
            Ar = j.rot.nullspace_mask().T
            qb = Q.qmul(Q.qmul(qa, j.orientation_offset), Q.axis_angle_to_quaternion(Ar @ dth if dth.size else np.zeros(3)))
            xb = xa + Q.qrot(j.vertex_parent + (At @ dx if dx.size else np.zeros(3)), qa) - Q.qrot(j.vertex_child, qb)
            x[j.child], q[j.child] = xb, qb
        return pack_maximal_state(x, np.zeros((self.Nb, 3)), q, np.zeros((self.Nb, 3)))

    def minimal_coordinates(self, z: np.ndarray) -> Dict[str, np.ndarray]:
        """joints/minimal.jl:4-8 → translational/minimal.jl:57-59, rotational/minimal.jl:62-67."""
        x, _, q, _ = unpack_maximal_state(z)
        out = {}
        for j in self.joints:
            xa = np.zeros(3) if j.parent < 0 else x[j.parent]
            qa = np.array([1.0, 0, 0, 0]) if j.parent < 0 else q[j.parent]
            xb, qb = x[j.child], q[j.child]
            d = xb + Q.qrot(j.vertex_child, qb) - (xa + Q.qrot(j.vertex_parent, qa))
            dt = j.tra.nullspace_mask() @ Q.qrot(d, Q.qinv(qa))
            qrel = Q.qmul(Q.qmul(Q.qinv(j.orientation_offset), Q.qinv(qa)), qb)
            dr = j.rot.nullspace_mask() @ Q.rotation_vector(qrel)
            out[j.name] = np.concatenate([dt, dr])
        return out

    # ------------------------------------------------------------------- IO
    def to_dict(self) -> dict:
        def el(e: JointElement):
            return dict(nlambda=e.nlambda, axis_mask=e.axis_mask.tolist(), spring=e.spring, damper=e.damper,
                        spring_offset=np.asarray(e.spring_offset).tolist(),
                        limit_lo=None if e.limit_lo is None else np.asarray(e.limit_lo).tolist(),
                        limit_hi=None if e.limit_hi is None else np.asarray(e.limit_hi).tolist())
        return dict(
            name=self.name, timestep=self.timestep, input_scaling=self.input_scaling, gravity=self.gravity.tolist(),
            bodies=[dict(name=b.name, mass=b.mass, inertia=np.asarray(b.inertia).tolist()) for b in self.bodies],
            joints=[dict(name=j.name, parent=j.parent, child=j.child, tra=el(j.tra), rot=el(j.rot),
                         vertex_parent=j.vertex_parent.tolist(), vertex_child=j.vertex_child.tolist(),
                         orientation_offset=j.orientation_offset.tolist()) for j in self.joints],
            contacts=[dict(name=c.name, body=c.body, friction=c.friction, normal=c.normal.tolist(),
                           tangent=c.tangent.tolist(), origin=c.origin.tolist(), radius=c.radius,
                           offset=c.offset.tolist(), **({} if c.model == "nonlinear" else {"model": c.model})) for c in self.contacts],
            z0=None if self.z0 is None else self.z0.tolist(),
        )

    @staticmethod
    def from_dict(d: dict) -> "Mechanism":
        def el(e):
            return JointElement(nlambda=int(e["nlambda"]), axis_mask=np.array(e["axis_mask"], dtype=float),
                                spring=float(e["spring"]), damper=float(e["damper"]),
                                spring_offset=np.array(e["spring_offset"], dtype=float),
                                limit_lo=None if e["limit_lo"] is None else np.array(e["limit_lo"], dtype=float),
                                limit_hi=None if e["limit_hi"] is None else np.array(e["limit_hi"], dtype=float))
        m = Mechanism(
            d["name"],
            [Body(b["name"], float(b["mass"]), np.array(b["inertia"], dtype=float)) for b in d["bodies"]],
            [Joint(j["name"], int(j["parent"]), int(j["child"]), el(j["tra"]), el(j["rot"]),
                   np.array(j["vertex_parent"], dtype=float), np.array(j["vertex_child"], dtype=float),
                   np.array(j["orientation_offset"], dtype=float)) for j in d["joints"]],
            [Contact(c["name"], int(c["body"]), float(c["friction"]), np.array(c["normal"], dtype=float),
                     np.array(c["tangent"], dtype=float), np.array(c["origin"], dtype=float), float(c["radius"]),
                     np.array(c["offset"], dtype=float), c.get("model", "nonlinear")) for c in d["contacts"]],
            timestep=d["timestep"], input_scaling=d["input_scaling"], gravity=d["gravity"])
        if d.get("z0") is not None:
            m.z0 = np.array(d["z0"], dtype=float)
        return m

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=1)

    @staticmethod
    def load(path: str) -> "Mechanism":
        with open(path) as f:
            return Mechanism.from_dict(json.load(f))


def get_mechanism(name: str, **overrides) -> Mechanism:
    """Mirror of DojoEnvironments.get_mechanism(:name) for the BASELINE models: loads the
    flattened descriptor generated from the reference's builder (tools/build_mechanisms.py).
    Overrides: timestep, input_scaling, gravity (mechanism kwargs, constructor.jl:47); springs / dampers (scalar or per joint) and
    joint_limits = {name: (lo, hi)} as the reference builders apply them (DojoEnvironments/src/utilities.jl:1-59); contact_type = "nonlinear" | "linear" |
    "impact" switches the model of every contact as the builders' `contact_type` keyword does (contacts/constructor.jl:117-128)."""
    m = Mechanism.load(os.path.join(MECHANISM_DIR, f"{name}.json"))
    if "timestep" in overrides:
        h = float(overrides.pop("timestep"))
        if m.input_scaling == m.timestep:
            m.input_scaling = h
        m.timestep = h
    if "input_scaling" in overrides:
        m.input_scaling = float(overrides.pop("input_scaling"))
    if "gravity" in overrides:
        g = overrides.pop("gravity")
        m.gravity = np.array([0.0, 0.0, g], dtype=float) if np.isscalar(g) else np.asarray(g, dtype=float)
    for key in ("springs", "dampers"):  # set_springs! / set_dampers! (DojoEnvironments/src/utilities.jl:1-39): every joint but a floating base
        if key in overrides:
            val = overrides.pop(key)
            vals = [float(val)] * m.Ne if np.isscalar(val) else [float(v) for v in val]
            for j, v in zip(m.joints, vals):
                if v == 0 or j.nimpulses == 0:
                    continue
                setattr(j.tra, key[:-1], v)
                setattr(j.rot, key[:-1], v)
    if "joint_limits" in overrides:  # set_limits (utilities.jl:41-59): one-dimensional joints only
        for jname, (lo, hi) in dict(overrides.pop("joint_limits")).items():
            j = m.joint_by_name(str(jname).lstrip(":"))
            if j.tra.nfree == 0 and j.rot.nfree == 1:
                j.rot.limit_lo, j.rot.limit_hi = np.array([lo], float), np.array([hi], float)
            elif j.tra.nfree == 1 and j.rot.nfree == 0:
                j.tra.limit_lo, j.tra.limit_hi = np.array([lo], float), np.array([hi], float)
            else:
                raise ValueError("joint limits can only be set for one-dimensional joints")
    if "contact_type" in overrides:
        ct = str(overrides.pop("contact_type")).lstrip(":")
        if ct not in CONTACT_TYPES:
            raise ValueError(f"unknown contact_type {ct!r}")
        for c in m.contacts:
            c.model = ct
    if overrides:
        raise TypeError(f"unknown mechanism options: {sorted(overrides)}")
    return m


# ---------------------------------------------------------------------- state packing
def pack_maximal_state(x, v, q, w) -> np.ndarray:
    """z = [x2(3) v15(3) q2(s,v1,v2,v3) ω15(3)] per body (mechanism/get.jl:107-120)."""
    x, v, q, w = (np.asarray(a, dtype=float) for a in (x, v, q, w))
    return np.concatenate([x, v, q, w], axis=1).reshape(-1)


def unpack_maximal_state(z):
    z = np.asarray(z, dtype=float).reshape(-1, 13)
    return z[:, 0:3], z[:, 3:6], z[:, 6:10], z[:, 10:13]
