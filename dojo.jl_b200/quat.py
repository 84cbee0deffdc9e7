"""Host-side quaternion helpers (scalar-first Hamilton quaternions, numpy).

Only used for model building / state initialisation on the host; the per-step
math lives in csrc/ (CUDA).  Conventions follow the reference:
src/orientation/quaternion.jl:13-32 (vector, Lmat, Rmat), rotate.jl:2-5
(vector_rotate), axis_angle.jl:1-11, mrp.jl:1-64 (rotation_vector).
"""
import numpy as np


def qmul(a, b):
    a0, a1, a2, a3 = a
    b0, b1, b2, b3 = b
    return np.array([
        a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0,
    ])


def qconj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def qinv(q):
    q = np.asarray(q, dtype=float)
    return qconj(q) / np.dot(q, q)


def qrot(v, q):
    """vector_rotate(v, q) = Vmat(q * v / q)  (rotate.jl:2-5)."""
    p = np.array([0.0, v[0], v[1], v[2]])
    return qmul(qmul(q, p), qinv(q))[1:]


def rot_x(t):
    return np.array([np.cos(t / 2), np.sin(t / 2), 0.0, 0.0])


def rot_y(t):
    return np.array([np.cos(t / 2), 0.0, np.sin(t / 2), 0.0])


def rot_z(t):
    return np.array([np.cos(t / 2), 0.0, 0.0, np.sin(t / 2)])


def rpy_to_quat(rpy):
    """urdf.jl:48-58: q = RotZ(y) * RotY(p) * RotX(r)."""
    return qmul(qmul(rot_z(rpy[2]), rot_y(rpy[1])), rot_x(rpy[0]))


def axis_angle_to_quaternion(x):
    x = np.asarray(x, dtype=float)
    th = np.linalg.norm(x)
    if th > 0.0:
        r = x / th
        return np.concatenate([[np.cos(0.5 * th)], np.sin(0.5 * th) * r])
    return np.array([1.0, 0.0, 0.0, 0.0])


def rotation_vector(q):
    """mrp.jl:1-64: axis * 4 atan(|mrp|), mrp = v / (1 + s)."""
    m = np.asarray(q[1:], dtype=float) / (q[0] + 1.0)
    mag = np.linalg.norm(m)
    if mag > 0:
        return 4.0 * np.arctan(mag) * m / mag
    return np.zeros(3)


def skew(p):
    return np.array([[0.0, -p[2], p[1]], [p[2], 0.0, -p[0]], [-p[1], p[0], 0.0]])


def orthogonal_rows(axis):
    """joints/orthogonal.jl:1-12.  V1, V2 from the SVD of skew(axis), V3 = axis."""
    axis = np.asarray(axis, dtype=float)
    n = np.linalg.norm(axis)
    if n > 0:
        axis = axis / n
    vt = np.linalg.svd(skew(axis))[2]
    return vt[0].copy(), vt[1].copy(), axis.copy()


def quaternion_map(w, h):
    return np.array([np.sqrt(4.0 / h ** 2 - np.dot(w, w)), w[0], w[1], w[2]])


def next_orientation(q2, w, h):
    """integrators/integrator.jl:15."""
    return qmul(q2, quaternion_map(w, h)) * h / 2.0
