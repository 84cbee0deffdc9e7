"""ctypes wrapper of tests/hostcheck/hostcheck.cpp -- TEST INFRASTRUCTURE (see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dojo.jl_b200", "csrc")
LIB = os.path.join(HERE, "_build", "libdojo_hostcheck.so")
_SRCS = [os.path.join(HERE, "hostcheck.cpp")] + [os.path.join(CSRC, f) for f in ("dojo_kinjac.cuh", "dojo_kin.cuh", "dojo_envs.cuh", "dojo_storage.cuh", "dojo_math.cuh", "dojo_plan.h")]


def build() -> str:
    srcs = [s for s in _SRCS if os.path.exists(s)]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", LIB + ".tmp", _SRCS[0]])
        os.replace(LIB + ".tmp", LIB)
    return LIB


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


class HostCheck:
    def __init__(self, mech):
        L = C.CDLL(build())
        L.hostcheck_create.restype = C.c_void_p
        L.hostcheck_create.argtypes = [C.c_int, C.c_int, C.c_double, _ip, _dp, _ip]
        L.hostcheck_destroy.argtypes = [C.c_void_p]
        L.hostcheck_num_input.argtypes = [C.c_void_p]
        for n in ("hostcheck_minimal_to_maximal", "hostcheck_maximal_to_minimal", "hostcheck_max_to_min_jacobian", "hostcheck_min_to_max_jacobian"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.hostcheck_minimal_gradients.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.hostcheck_env_pre.argtypes = [C.c_void_p, _ip, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.hostcheck_env_post.argtypes = [C.c_void_p, _ip, _dp, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]
        L.hostcheck_env_policy.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.hostcheck_storage.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        self.L, self.mech = L, mech
        jint = np.zeros((mech.Ne, 4), dtype=np.int32)
        jdbl = np.zeros((mech.Ne, 28))
        for i, j in enumerate(mech.joints):
            jint[i] = (j.parent, j.child, j.tra.nlambda, j.rot.nlambda)
            jdbl[i] = np.concatenate([j.vertex_parent, j.vertex_child, j.orientation_offset, np.asarray(j.tra.axis_mask).ravel(),
                                      np.asarray(j.rot.axis_mask).ravel()])
        order = np.asarray(mech.root_to_leaves_joints(), dtype=np.int32)
        self.h = C.c_void_p(L.hostcheck_create(mech.Ne, mech.Nb, float(mech.timestep), jint.ctypes.data_as(_ip), _d(jdbl), order.ctypes.data_as(_ip)))
        assert L.hostcheck_num_input(self.h) == mech.nu
        self.nm, self.ns = 2 * mech.nu, 12 * mech.Nb

    def __del__(self):
        try:
            self.L.hostcheck_destroy(self.h)
        except Exception:
            pass

    def minimal_to_maximal(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=float)
        Z = np.empty((X.shape[0], 13 * self.mech.Nb))
        self.L.hostcheck_minimal_to_maximal(self.h, X.shape[0], _d(X), _d(Z))
        return Z

    def maximal_to_minimal(self, Z):
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        X = np.empty((Z.shape[0], self.nm))
        self.L.hostcheck_maximal_to_minimal(self.h, Z.shape[0], _d(Z), _d(X))
        return X

    def maximal_to_minimal_jacobian(self, Z):
        """[B, 2nu, 12Nb]"""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        J = np.empty((Z.shape[0], self.ns, self.nm))  # column-major per environment
        self.L.hostcheck_max_to_min_jacobian(self.h, Z.shape[0], _d(Z), _d(J))
        return J.transpose(0, 2, 1)

    def minimal_to_maximal_jacobian(self, Z):
        """[B, 12Nb, 2nu]"""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        J = np.empty((Z.shape[0], self.nm, self.ns))
        self.L.hostcheck_min_to_max_jacobian(self.h, Z.shape[0], _d(Z), _d(J))
        return J.transpose(0, 2, 1)

    def minimal_gradients(self, Z, Zn, Fz, Fu):
        """Fz [B, 12Nb, 12Nb], Fu [B, 12Nb, nu] (math layout) -> Gx [B, 2nu, 2nu], Gu [B, 2nu, nu]"""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        Zn = np.ascontiguousarray(np.atleast_2d(Zn), dtype=float)
        B = Z.shape[0]
        Fzc = np.ascontiguousarray(np.asarray(Fz, dtype=float).reshape(B, self.ns, self.ns).transpose(0, 2, 1))
        Fuc = np.ascontiguousarray(np.asarray(Fu, dtype=float).reshape(B, self.ns, self.mech.nu).transpose(0, 2, 1))
        Gx = np.empty((B, self.nm, self.nm))
        Gu = np.empty((B, self.mech.nu, self.nm))
        self.L.hostcheck_minimal_gradients(self.h, B, _d(Z), _d(Zn), _d(Fzc), _d(Fuc), _d(Gx), _d(Gu))
        return Gx.transpose(0, 2, 1), Gu.transpose(0, 2, 1)

    # ---- environment layer (dojo_envs.cuh)
    @staticmethod
    def _spec_arrays(spec):
        si = np.array([spec.n_unactuated, spec.contact_obs, spec.forward_index, spec.healthy_index, spec.bound_index], dtype=np.int32)
        sd = np.array([spec.w_forward, spec.w_control, spec.w_contact, spec.survive_reward, spec.healthy_min, spec.healthy_max, spec.bound_abs])
        return si, sd

    def env_pre(self, spec, S, A):
        m = self.mech
        si, sd = self._spec_arrays(spec)
        S = np.ascontiguousarray(np.atleast_2d(S), dtype=float)
        A = None if A is None else np.ascontiguousarray(np.atleast_2d(A), dtype=float)
        B = S.shape[0]
        Z, U = np.empty((B, 13 * m.Nb)), np.empty((B, m.nu))
        self.L.hostcheck_env_pre(self.h, si.ctypes.data_as(_ip), _d(sd), m.Ni, B, _d(S), None if A is None else _d(A), _d(Z), _d(U))
        return Z, U

    def env_post(self, spec, S, A, Zn, sol, ret=None, dead=None):
        m = self.mech
        si, sd = self._spec_arrays(spec)
        S = np.ascontiguousarray(np.atleast_2d(S), dtype=float)
        A = None if A is None else np.ascontiguousarray(np.atleast_2d(A), dtype=float)
        Zn = np.ascontiguousarray(np.atleast_2d(Zn), dtype=float)
        sol = np.ascontiguousarray(np.atleast_2d(sol), dtype=float)
        B = S.shape[0]
        offs = np.array([m.contact_sol_offset(c) for c in range(m.Ni)] or [0], dtype=np.int32)
        Sn, reward, done = np.empty_like(S), np.empty(B), np.zeros(B, dtype=np.int32)
        self.L.hostcheck_env_post(self.h, si.ctypes.data_as(_ip), _d(sd), m.Ni, m.nres, offs.ctypes.data_as(_ip), B, _d(S),
                                  None if A is None else _d(A), _d(Zn), _d(sol), _d(Sn), _d(reward), done.ctypes.data_as(_ip),
                                  None if ret is None else _d(ret), None if dead is None else dead.ctypes.data_as(_ip))
        return Sn, reward, done

    # ---- storage / diagnostics (dojo_storage.cuh)
    def storage(self, Z, Zn, U, sol):
        """-> (per body [B, Nb, 12] = px pq vl wl, diag [B, 8] = p_linear p_angular kinetic potential)"""
        m = self.mech
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        Zn = np.ascontiguousarray(np.atleast_2d(Zn), dtype=float)
        U = np.ascontiguousarray(np.atleast_2d(U), dtype=float)
        sol = np.ascontiguousarray(np.atleast_2d(sol), dtype=float)
        B = Z.shape[0]
        def half(e):
            off = list(np.asarray(e.spring_offset, dtype=float).ravel()[:3])
            return [e.nlimits, e.spring if e.nlambda < 3 else 0.0, e.damper if e.nlambda < 3 else 0.0] + off + [0.0] * (3 - len(off))
        jext = np.array([half(j.rot) + half(j.tra) for j in m.joints], dtype=float)
        bdbl = np.array([[b.mass] + list(np.asarray(b.inertia, dtype=float).ravel()) for b in m.bodies], dtype=float)
        g = np.ascontiguousarray(m.gravity, dtype=float)
        body, diag = np.empty((B, m.Nb, 12)), np.empty((B, 8))
        self.L.hostcheck_storage(self.h, _d(jext), _d(bdbl), m.nres, float(m.input_scaling), _d(g), B, _d(Z), _d(Zn), _d(U), _d(sol), _d(body), _d(diag))
        return body, diag

    def env_policy(self, S, Theta, mean=None, std=None):
        """Theta [B, na, ns] -> actions [B, na]"""
        S = np.ascontiguousarray(np.atleast_2d(S), dtype=float)
        B, ns = S.shape
        na = Theta.shape[1]
        Tc = np.ascontiguousarray(np.asarray(Theta, dtype=float).transpose(0, 2, 1))
        A = np.empty((B, na))
        self.L.hostcheck_env_policy(ns, na, B, _d(S), _d(Tc), None if mean is None else _d(np.ascontiguousarray(mean, dtype=float)),
                                    None if std is None else _d(np.ascontiguousarray(std, dtype=float)), _d(A))
        return A
