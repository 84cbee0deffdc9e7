#!/usr/bin/env python
"""Generate the flattened mechanism descriptors dojo.jl_b200/mechanisms/*.json.

The reference builds its models in Julia (no Julia in this image), so this script
restates the model-building path for the BASELINE mechanisms and writes the result
as data:

  URDF parsing ............ /root/reference/src/mechanism/urdf.jl:32-58 (inertia, pose),
                            :171-199 (parse_link), :252-261 (parse_joint), :283-331 (parse_joints)
  frame fix-up ............ urdf.jl:420-495 (set_parsed_values!)
  joint prototypes ........ src/joints/prototypes.jl:6-14 (Fixed), :94-117 (Revolute), :429-447 (Floating)
  joint limits ............ DojoEnvironments/src/utilities.jl:41-57 (set_limits), src/joints/limits.jl:31-61
  contact tables/poses .... DojoEnvironments/src/mechanisms/{pendulum,ant,quadruped,atlas}/mechanism.jl

It READS the reference's URDF data files and therefore only runs in the build
container (``/root/reference`` does not exist on the GPU box); the JSON it writes is
committed.  Body order = URDF file order (the reference's order is Julia Dict hash
order, urdf.jl:285 -- SURVEY.md Q7); joint order = floating base first, then URDF order.

    python tools/build_mechanisms.py [--reference /root/reference]
"""
import argparse
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dojo_jl_b200  # noqa: E402
from dojo_jl_b200 import quat as Q  # noqa: E402
from dojo_jl_b200.mechanism import Body, Contact, Joint, JointElement, Mechanism, MECHANISM_DIR, pack_maximal_state  # noqa: E402

Z_AXIS = np.array([0.0, 0.0, 1.0])
X_AXIS = np.array([1.0, 0.0, 0.0])
Y_AXIS = np.array([0.0, 1.0, 0.0])


def fvec(s, default):
    return np.array([float(t) for t in (s if s is not None else default).split()])


def parse_pose(el):
    if el is None:
        return np.zeros(3), np.array([1.0, 0, 0, 0])
    return fvec(el.get("xyz"), "0 0 0"), Q.rpy_to_quat(fvec(el.get("rpy"), "0 0 0"))


def element(nlambda, axis=None, damper=0.0, spring=0.0, limits=None):
    """Translational{T,Nλ} / Rotational{T,Nλ} constructor: masks from orthogonal_rows(axis)."""
    V1, V2, V3 = Q.orthogonal_rows(np.zeros(3) if axis is None else axis)
    e = JointElement(nlambda=nlambda, axis_mask=np.stack([V1, V2, V3]), spring=spring, damper=damper,
                     spring_offset=np.zeros(3 - nlambda))
    if limits is not None:
        e.limit_lo, e.limit_hi = np.atleast_1d(limits[0]).astype(float), np.atleast_1d(limits[1]).astype(float)
    return e


def make_joint(jtype, name, parent, child, axis, parent_vertex, orientation_offset, damper):
    if jtype in ("revolute", "continuous"):  # prototypes.jl:94-117
        tra, rot = element(3, damper=damper), element(2, axis=axis, damper=damper)
    elif jtype == "fixed":  # prototypes.jl:6-14 (takes no damper, urdf.jl:229-230)
        tra, rot = element(3), element(3)
    elif jtype == "floating":  # prototypes.jl:429-447
        tra, rot = element(0, damper=damper), element(0, damper=damper)
    elif jtype == "prismatic":
        tra, rot = element(2, axis=axis, damper=damper), element(3, damper=damper)
    else:
        raise NotImplementedError(jtype)
    return Joint(name, parent, child, tra, rot, vertex_parent=np.asarray(parent_vertex, float),
                 vertex_child=np.zeros(3), orientation_offset=np.asarray(orientation_offset, float))


def mechanism_from_urdf(path, name, floating, parse_dampers=True, **kw):
    root = ET.parse(path).getroot()
    assert root.tag == "robot"
    links, local_pose = [], {}
    for xl in root.findall("link"):
        inert = xl.find("inertial")
        if inert is None:
            x, q, m, J = np.zeros(3), np.array([1.0, 0, 0, 0]), 0.0, np.zeros((3, 3))
        else:
            x, q = parse_pose(inert.find("origin"))
            xi = inert.find("inertia")
            g = lambda k, d=None: float(xi.get(k, d))  # noqa: E731
            J = np.array([[g("ixx"), g("ixy", "0"), g("ixz", "0")],
                          [g("ixy", "0"), g("iyy"), g("iyz", "0")],
                          [g("ixz", "0"), g("iyz", "0"), g("izz")]])
            m = float(inert.find("mass").get("value", "0"))
        links.append(Body(xl.get("name"), m, J))
        local_pose[xl.get("name")] = (x, q)
    index = {b.name: i for i, b in enumerate(links)}
    xjoints = root.findall("joint")
    children = {xj.find("child").get("link") for xj in xjoints}
    roots = [b.name for b in links if b.name not in children]
    assert len(roots) == 1 and floating, "only floating-base URDFs are needed for the BASELINE models"
    joints = [make_joint("floating", "floating_base", -1, index[roots[0]], None, np.zeros(3), [1.0, 0, 0, 0], 0.0)]
    for xj in xjoints:
        x, q = parse_pose(xj.find("origin"))
        ax = xj.find("axis")
        axis = fvec(ax.get("xyz") if ax is not None else None, "1 0 0")
        dyn = xj.find("dynamics")
        damper = float(dyn.get("damping", "0")) if (parse_dampers and dyn is not None) else 0.0
        joints.append(make_joint(xj.get("type"), xj.get("name"), index[xj.find("parent").get("link")],
                                 index[xj.find("child").get("link")], axis, x, q, damper))
    mech = Mechanism(name, links, joints, [], **kw)

    # set_parsed_values! (urdf.jl:420-495), root → leaves
    xw = np.zeros((len(links), 3))
    qw = np.tile(np.array([1.0, 0, 0, 0]), (len(links), 1))
    xjw, qjw = {}, {}
    for ji in mech.root_to_leaves_joints():
        j = joints[ji]
        x_cj, q_cj = j.vertex_parent, j.orientation_offset
        xbl, qbl = local_pose[links[j.child].name]
        if j.parent < 0:
            xpb, qpb = np.zeros(3), np.array([1.0, 0, 0, 0])
            xpj, qpj = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            xpb, qpb = xw[j.parent], qw[j.parent]
            pj = mech.parent_joint_of(j.parent)
            xpj, qpj = xjw[pj], qjw[pj]
        xjl = Q.qrot(xpj + Q.qrot(x_cj, qpj) - xpb, Q.qinv(qpb))
        qjl = Q.qmul(Q.qmul(Q.qinv(qpb), qpj), q_cj)
        xjw[ji] = xpb + Q.qrot(xjl, qpb)
        qjw[ji] = Q.qmul(qpb, qjl)
        j.orientation_offset = Q.qmul(qjl, qbl)
        j.vertex_parent = xjl
        j.vertex_child = Q.qrot(-xbl, Q.qinv(qbl))
        # place the child (bodies/set.jl:59-70)
        qc = Q.qmul(qpb, j.orientation_offset)
        xw[j.child] = xpb + Q.qrot(j.vertex_parent, qpb) - Q.qrot(j.vertex_child, qc)
        qw[j.child] = qc
    return mech


def add_limits(mech, limits):
    for jname, (lo, hi) in limits.items():
        j = mech.joint_by_name(jname)
        if j.tra.nfree == 0 and j.rot.nfree == 1:
            j.rot.limit_lo, j.rot.limit_hi = np.array([lo], float), np.array([hi], float)
        elif j.tra.nfree == 1 and j.rot.nfree == 0:
            j.tra.limit_lo, j.tra.limit_hi = np.array([lo], float), np.array([hi], float)
        else:
            raise ValueError("joint limits can only be set for one-dimensional joints")


def nonlinear_contact(mech, name, body, normal, friction, origin, radius):
    """contacts/nonlinear.jl:26-48: tangent/normal from inv([V1 V2 V3])."""
    V1, V2, V3 = Q.orthogonal_rows(normal)
    Ainv = np.linalg.inv(np.stack([V1, V2, V3], axis=1))
    return Contact(name, mech.body_index(body), float(friction), Ainv[2].copy(), Ainv[0:2].copy(),
                   np.asarray(origin, float), float(radius))


def build_pendulum():
    m, L = 1.0, 1.0
    x, y, z = 0.1, 0.1, L  # Box(0.1, 0.1, L, m): bodies/shapes.jl:83-92
    J = m / 12.0 * np.diag([y ** 2 + z ** 2, x ** 2 + z ** 2, x ** 2 + y ** 2])
    body = Body("pendulum", m, J)
    j = make_joint("revolute", "joint", -1, 0, X_AXIS, (L + 0.1) * Z_AXIS, [1.0, 0, 0, 0], 0.0)
    j.vertex_child = 0.5 * L * Z_AXIS
    mech = Mechanism("pendulum", [body], [j], [], timestep=0.01)
    mech.z0 = mech.forward_kinematics({"joint": [np.pi / 4]})
    return mech


def capsule_inertia(r, h, m):
    """Capsule(r, h, m): bodies/shapes.jl:157-179"""
    vc, vh = np.pi * h * r ** 2, np.pi * 4.0 / 3.0 * r ** 3 / 2.0
    mc, mh = m * vc / (vc + 2 * vh), m * vh / (vc + 2 * vh)
    ixx = mc * (h ** 2 / 12.0 + r ** 2 / 4.0) + 2.0 * (83.0 / 320 * mh * r ** 2 + mh * (3.0 / 8.0 * r + 0.5 * h) ** 2)
    izz = mc * 0.5 * r ** 2 + 2.0 * (mh * 2.0 / 5.0 * r ** 2 / 2.0)
    return np.diag([ixx, ixx, izz])


def build_cartpole():
    """DojoEnvironments/src/mechanisms/cartpole/mechanism.jl:1-62 (defaults): cart on a Prismatic joint along y, pole on a Revolute
    joint about x with child_vertex = -L/2 z; initialize_cartpole!: position 0, orientation pi/4.  Springs / dampers / joint
    limits are the builder's options (set on the loaded Mechanism: the prismatic joint exercises the translational ones)."""
    r, L = 0.075, 1.0
    bodies = [Body("cart", 1.0, capsule_inertia(1.5 * r, 1.0, 1.0)), Body("pole", 1.0, capsule_inertia(r, L, 1.0))]
    j0 = make_joint("prismatic", "cart_joint", -1, 0, Y_AXIS, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    j1 = make_joint("revolute", "pole_joint", 0, 1, X_AXIS, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    j1.vertex_child = -0.5 * L * Z_AXIS
    mech = Mechanism("cartpole", bodies, [j0, j1], [], timestep=0.01)
    mech.z0 = mech.forward_kinematics({"cart_joint": [0.0], "pole_joint": [np.pi / 4]})
    return mech


def build_slider():
    """DojoEnvironments/src/mechanisms/slider/mechanism.jl:1-53 (defaults): Box(0.1, 0.1, 1, 1) on a Prismatic joint along z with
    child_vertex = z/2; initialize_slider!: position 0 (body centre at z = -1/2).  Springs / dampers are the builder's options."""
    x, y, z, m = 0.1, 0.1, 1.0, 1.0
    body = Body("pbody", m, m / 12.0 * np.diag([y ** 2 + z ** 2, x ** 2 + z ** 2, x ** 2 + y ** 2]))
    j = make_joint("prismatic", "joint", -1, 0, Z_AXIS, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    j.vertex_child = 0.5 * Z_AXIS
    mech = Mechanism("slider", [body], [j], [], timestep=0.01)
    mech.z0 = mech.forward_kinematics({"joint": [0.0]})
    return mech


def build_raiberthopper():
    """DojoEnvironments/src/mechanisms/raiberthopper/mechanism.jl:1-86 (defaults): Sphere(0.1, 4.18) body on a Floating joint, Sphere(0.05,
    0.52) foot on a Prismatic joint along z with damper 0.1 (set_dampers!(joints, [0; 0.1]): a TRANSLATIONAL damper), contacts on the
    foot and on the body (friction 0.5); initialize_raiberthopper!: leg_length 0.5, body at z = 0.5 + 0.05."""
    rb, rf = 0.1, 0.05
    bodies = [Body("body", 4.18, 0.4 * 4.18 * rb * rb * np.eye(3)), Body("foot", 0.52, 0.4 * 0.52 * rf * rf * np.eye(3))]
    j0 = make_joint("floating", "floating_joint", -1, 0, None, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    j1 = make_joint("prismatic", "prismatic_joint", 0, 1, Z_AXIS, np.zeros(3), [1.0, 0, 0, 0], 0.1)
    mech = Mechanism("raiberthopper", bodies, [j0, j1], [], timestep=0.05)
    mech.contacts = [nonlinear_contact(mech, "foot_contact", "foot", Z_AXIS, 0.5, np.zeros(3), rf),
                     nonlinear_contact(mech, "body_contact", "body", Z_AXIS, 0.5, np.zeros(3), rb)]
    mech.z0 = mech.forward_kinematics({"floating_joint": [0, 0, 0.5 + rf, 0, 0, 0], "prismatic_joint": [-0.5]})
    return mech


def build_sphere():
    """DojoEnvironments/src/mechanisms/sphere/mechanism.jl:1-67 (defaults): Sphere(0.5, 1) on a Floating joint, one contact of
    radius 0.5 at the centre; initialize_sphere!: centre at z = 0.5 + r, velocity [1, 0, 0]."""
    m, r = 1.0, 0.5
    body = Body("sphere", m, 0.4 * m * r * r * np.eye(3))  # bodies/shapes.jl:247-256
    j = make_joint("floating", "floating_joint", -1, 0, None, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    mech = Mechanism("sphere", [body], [j], [], timestep=0.01)
    mech.contacts = [nonlinear_contact(mech, "contact", "sphere", Z_AXIS, 0.8, np.zeros(3), r)]
    mech.z0 = pack_maximal_state([[0.0, 0.0, 0.5 + r]], [[1.0, 0.0, 0.0]], [[1.0, 0, 0, 0]], [[0.0, 0.0, 0.0]])
    return mech


def build_block():
    """DojoEnvironments/src/mechanisms/block/mechanism.jl:1-95 (defaults): Box(0.5, 0.5, 0.5, 1) on a Floating joint, 8 corner
    contacts of radius 0; initialize_block!: centre at z = 1 + 0.25 (angular velocity: the reference draws randn(3); here 0)."""
    m, a = 1.0, 0.5
    body = Body("block", m, m / 12.0 * np.diag([2 * a * a, 2 * a * a, 2 * a * a]))  # bodies/shapes.jl:83-92
    j = make_joint("floating", "joint", -1, 0, None, np.zeros(3), [1.0, 0, 0, 0], 0.0)
    mech = Mechanism("block", [body], [j], [], timestep=0.01)
    k = 0
    for sz in (-1, 1):
        for sx in (1, -1):
            for sy in (1, -1):
                k += 1
                mech.contacts.append(nonlinear_contact(mech, f"contact{k}", "block", Z_AXIS, 0.8, 0.5 * a * np.array([sx, sy, sz], float), 0.0))
    mech.z0 = pack_maximal_state([[0.0, 0.0, 1.0 + 0.5 * a]], [[0.0, 0.0, 0.0]], [[1.0, 0, 0, 0]], [[0.0, 0.0, 0.0]])
    return mech


def build_ant(ref):
    path = os.path.join(ref, "DojoEnvironments/src/mechanisms/ant/dependencies/ant.urdf")
    mech = mechanism_from_urdf(path, "ant", floating=True, timestep=0.05)
    d = np.pi / 180
    add_limits(mech, {"hip_1": (-30 * d, 30 * d), "ankle_1": (30 * d, 70 * d),
                      "hip_2": (-30 * d, 30 * d), "ankle_2": (-70 * d, -30 * d),
                      "hip_3": (-30 * d, 30 * d), "ankle_3": (-70 * d, -30 * d),
                      "hip_4": (-30 * d, 30 * d), "ankle_4": (30 * d, 70 * d)})
    mu = 0.5
    feet = ["front_left_foot", "front_right_foot", "left_back_foot", "right_back_foot"]
    forig = [[0.2, 0.2, 0], [-0.2, 0.2, 0], [-0.2, -0.2, 0], [0.2, -0.2, 0]]
    cs = [nonlinear_contact(mech, f"{b}_contact", b, Z_AXIS, mu, o, 0.08) for b, o in zip(feet, forig)]
    cs.append(nonlinear_contact(mech, "torso_contact", "torso", Z_AXIS, mu, np.zeros(3), 0.25))
    aux = ["aux_1", "aux_2", "aux_3", "aux_4"]
    aorig = [[-0.1, -0.1, 0], [0.1, -0.1, 0], [0.1, 0.1, 0], [-0.1, 0.1, 0]]
    cs += [nonlinear_contact(mech, f"{b}_contact", b, Z_AXIS, mu, o, 0.08) for b, o in zip(aux, aorig)]
    mech.contacts = cs
    a = 0.25 * np.pi  # initialize_ant! (ant/mechanism.jl:93-111)
    mech.z0 = mech.forward_kinematics({"floating_base": [0, 0, 0.5, 0, 0, 0],
                                       "ankle_1": [a], "ankle_4": [a], "ankle_2": [-a], "ankle_3": [-a]})
    return mech


def build_quadruped(ref):
    path = os.path.join(ref, "DojoEnvironments/src/mechanisms/quadruped/dependencies/gazebo_a1.urdf")
    mech = mechanism_from_urdf(path, "quadruped", floating=True, timestep=0.01)
    groups = ["FR", "FL", "RR", "RL"]
    for g in groups:  # spring_offset=true (quadruped/mechanism.jl:31-49); springs themselves are 0
        mech.joint_by_name(f"{g}_hip_joint").rot.spring_offset = np.array([0.0])
        mech.joint_by_name(f"{g}_thigh_joint").rot.spring_offset = np.array([0.9])
        mech.joint_by_name(f"{g}_calf_joint").rot.spring_offset = np.array([-1.425])
    lim = {}
    for g in groups:
        lim[f"{g}_hip_joint"] = (-0.5, 0.5)
        lim[f"{g}_thigh_joint"] = (-0.5, 1.5)
        lim[f"{g}_calf_joint"] = (-2.5, -1.0)
    add_limits(mech, lim)
    mu = 0.8
    cs = [nonlinear_contact(mech, f"{g}_calf_contact", f"{g}_calf", Z_AXIS, mu, [-0.006, 0, -0.092], 0.021) for g in groups]
    torig = [[-0.005, -0.023, -0.16], [-0.005, 0.023, -0.16], [-0.005, -0.023, -0.16], [-0.005, 0.023, -0.16]]
    cs += [nonlinear_contact(mech, f"{g}_thigh_contact", f"{g}_thigh", Z_AXIS, mu, o, 0.023) for g, o in zip(groups, torig)]
    cs += [nonlinear_contact(mech, f"{g}_hip_contact", f"{g}_hip", Z_AXIS, mu, [0, 0.05, 0], 0.05) for g in groups]
    mech.contacts = cs
    coords = {"floating_base": [0, 0, 0.43, 0, 0, 0]}  # initialize_quadruped! (:111-127)
    for g in groups:
        coords[f"{g}_hip_joint"] = [0.0]
        coords[f"{g}_thigh_joint"] = [np.pi / 4]
        coords[f"{g}_calf_joint"] = [-np.pi / 2]
    mech.z0 = mech.forward_kinematics(coords)
    return mech


def build_atlas(ref):
    path = os.path.join(ref, "DojoEnvironments/src/mechanisms/atlas/dependencies/atlas_simple.urdf")
    mech = mechanism_from_urdf(path, "atlas", floating=True, timestep=0.01)
    mu = 0.8
    forig = [[-0.08, -0.04, 0.015], [0.12, -0.02, 0.015], [-0.08, 0.04, 0.015], [0.12, 0.02, 0.015]]
    cs = []
    for side in ("l", "r"):
        for nm, o in zip(["RR", "FR", "RL", "RR2"], forig):
            cs.append(nonlinear_contact(mech, f"{side}_{nm}", f"{side}_foot", Z_AXIS, mu, o, 0.025))
    bnames = ["l_hand", "r_hand", "l_lleg", "r_lleg", "l_clav", "r_clav", "pelvis", "l_uarm", "r_uarm", "head", "utorso", "utorso"]
    names = ["l_hand", "r_hand", "l_knee", "r_knee", "l_clavis", "r_clavis", "pelvis", "l_elbow", "r_elbow", "head", "backpack_bottom", "backpack_top"]
    borig = [[0, 0, 0], [0, 0, 0], [0.025, 0, 0.175], [0.025, 0, 0.175], [0, -0.05, -0.075], [0, -0.05, -0.075], [0, 0, 0.05],
             [0, -0.185, 0], [0, -0.185, 0], [0, 0, 0], [-0.095, 0, 0.25], [-0.095, 0, -0.2]]
    radii = [0.06, 0.06, 0.075, 0.075, 0.11, 0.11, 0.19, 0.085, 0.085, 0.175, 0.15, 0.15]
    cs += [nonlinear_contact(mech, n, b, Z_AXIS, mu, o, r) for n, b, o, r in zip(names, bnames, borig, radii)]
    mech.contacts = cs
    mech.z0 = mech.forward_kinematics({"floating_base": [0, 0, 0.9385, 0, 0, 0]})  # initialize_atlas!
    return mech


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    os.makedirs(MECHANISM_DIR, exist_ok=True)
    for mech in (build_pendulum(), build_cartpole(), build_slider(), build_raiberthopper(), build_sphere(), build_block(), build_ant(args.reference), build_quadruped(args.reference), build_atlas(args.reference)):
        mech.save(os.path.join(MECHANISM_DIR, f"{mech.name}.json"))
        print(f"{mech.name}: Nb={mech.Nb} Ne={mech.Ne} Ni={mech.Ni} nres={mech.nres} nu={mech.nu}")


if __name__ == "__main__":
    main()
