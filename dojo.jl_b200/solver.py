"""ctypes binding of the product library libdojo_b200.so (include/dojo_b200.h).

This is the ONLY compute path of the package: if the CUDA library is missing or no CUDA device is
present, construction fails loudly (RuntimeError) -- there is no CPU fallback.
"""
import ctypes as C
import os
from typing import Optional

import numpy as np

from . import capi
from .mechanism import Mechanism

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdojo_b200.so")

DOJO_FLAG_Q1_LITERAL_RETURN = 1
DOJO_FLAG_Q2_LITERAL_GRADIENTS = 2  # get_maximal_gradients! literally: data Jacobian after update_state! (include/dojo_b200.h)
STATUS = {0: "success", 1: "failed", 2: "excessive_angular_velocity", 3: "nonfinite"}

EXPORTS = ["dojo_default_options", "dojo_create", "dojo_destroy", "dojo_last_error", "dojo_num_state", "dojo_num_input",
           "dojo_num_residual", "dojo_num_grad_state", "dojo_shared_bytes_per_env", "dojo_step", "dojo_step_async",
           "dojo_step_grad", "dojo_step_grad_async", "dojo_rollout", "dojo_rollout_async", "dojo_launch_count",
           "dojo_num_minimal", "dojo_minimal_to_maximal", "dojo_maximal_to_minimal", "dojo_minimal_to_maximal_async",
           "dojo_maximal_to_minimal_async", "dojo_step_minimal", "dojo_step_minimal_flags", "dojo_maximal_to_minimal_jacobian", "dojo_minimal_to_maximal_jacobian",
           "dojo_maximal_to_minimal_jacobian_async", "dojo_minimal_to_maximal_jacobian_async", "dojo_minimal_gradients", "dojo_env_num_state", "dojo_env_num_action", "dojo_env_step",
           "dojo_env_step_async", "dojo_env_reset", "dojo_env_rollout", "dojo_env_policy_rollout", "dojo_update_params", "dojo_num_contact_data", "dojo_step_grad_contact",
           "dojo_step_grad_contact_async", "dojo_step_record", "dojo_step_record_async", "dojo_simulate_record",
           "dojo_gather_create", "dojo_gather_export", "dojo_gather_connect", "dojo_gather_buffer", "dojo_gather_destroy", "dojo_step_gather_async",
           "dojo_step_grad_gather_async"]

_lib = None


def load_library():
    """Load libdojo_b200.so (built in-tree by build.py).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  dojo.jl_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = capi.c_double_p, capi.c_int32_p, C.c_void_p
    op = C.POINTER(capi.DojoSolverOptions)
    L.dojo_default_options.argtypes = [op]
    L.dojo_create.argtypes = [C.POINTER(capi.DojoMechanismDesc), C.c_int, C.c_int, C.POINTER(vp)]
    L.dojo_create.restype = C.c_int
    L.dojo_destroy.argtypes = [vp]
    L.dojo_last_error.argtypes = [vp]
    L.dojo_last_error.restype = C.c_char_p
    for n in ("dojo_num_state", "dojo_num_input", "dojo_num_residual", "dojo_num_grad_state", "dojo_shared_bytes_per_env"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = C.c_int
    L.dojo_launch_count.argtypes = [vp]
    L.dojo_launch_count.restype = C.c_int64
    L.dojo_step.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_uint32]
    L.dojo_step.restype = C.c_int
    L.dojo_step_async.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.dojo_step_async.restype = C.c_int
    L.dojo_step_grad.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32]
    L.dojo_step_grad.restype = C.c_int
    L.dojo_step_grad_async.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.dojo_step_grad_async.restype = C.c_int
    L.dojo_rollout.argtypes = [vp, op, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.dojo_rollout.restype = C.c_int
    L.dojo_rollout_async.argtypes = [vp, op, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    L.dojo_rollout_async.restype = C.c_int
    L.dojo_num_minimal.argtypes = [vp]
    L.dojo_num_minimal.restype = C.c_int
    for name in ("dojo_minimal_to_maximal", "dojo_maximal_to_minimal"):
        getattr(L, name).argtypes = [vp, C.c_int, vp, vp]
        getattr(L, name).restype = C.c_int
        getattr(L, name + "_async").argtypes = [vp, C.c_int, vp, vp, vp]
        getattr(L, name + "_async").restype = C.c_int
    L.dojo_step_minimal.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp]
    L.dojo_step_minimal.restype = C.c_int
    L.dojo_step_minimal_flags.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, C.c_uint32]
    L.dojo_step_minimal_flags.restype = C.c_int
    for name in ("dojo_maximal_to_minimal_jacobian", "dojo_minimal_to_maximal_jacobian"):
        getattr(L, name).argtypes = [vp, C.c_int, vp, vp]
        getattr(L, name).restype = C.c_int
        getattr(L, name + "_async").argtypes = [vp, C.c_int, vp, vp, vp]
        getattr(L, name + "_async").restype = C.c_int
    L.dojo_minimal_gradients.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_minimal_gradients.restype = C.c_int
    ep = C.POINTER(capi.DojoEnvSpec)
    for name in ("dojo_env_num_state", "dojo_env_num_action"):
        getattr(L, name).argtypes = [vp, ep]
        getattr(L, name).restype = C.c_int
    L.dojo_env_step.argtypes = [vp, op, ep, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_env_step.restype = C.c_int
    L.dojo_env_step_async.argtypes = [vp, op, ep, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_env_step_async.restype = C.c_int
    L.dojo_env_reset.argtypes = [vp, ep, C.c_int, vp, vp, vp]
    L.dojo_env_reset.restype = C.c_int
    L.dojo_env_rollout.argtypes = [vp, op, ep, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.dojo_env_rollout.restype = C.c_int
    L.dojo_env_policy_rollout.argtypes = [vp, op, ep, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_env_policy_rollout.restype = C.c_int
    L.dojo_num_contact_data.argtypes = [vp]
    L.dojo_num_contact_data.restype = C.c_int
    L.dojo_step_grad_contact.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_step_grad_contact.restype = C.c_int
    L.dojo_step_grad_contact_async.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.dojo_step_grad_contact_async.restype = C.c_int
    L.dojo_update_params.argtypes = [vp, C.POINTER(capi.DojoMechanismDesc)]
    L.dojo_update_params.restype = C.c_int
    L.dojo_step_record.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_step_record.restype = C.c_int
    L.dojo_step_record_async.argtypes = [vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_step_record_async.restype = C.c_int
    L.dojo_simulate_record.argtypes = [vp, op, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.dojo_simulate_record.restype = C.c_int
    L.dojo_gather_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.dojo_gather_create.restype = C.c_int
    L.dojo_gather_export.argtypes = [vp, vp]
    L.dojo_gather_export.restype = C.c_int
    L.dojo_gather_connect.argtypes = [vp, vp]
    L.dojo_gather_connect.restype = C.c_int
    L.dojo_gather_buffer.argtypes = [vp]
    L.dojo_gather_buffer.restype = vp
    L.dojo_gather_destroy.argtypes = [vp]
    L.dojo_gather_destroy.restype = C.c_int
    L.dojo_step_gather_async.argtypes = [vp, vp, op, C.c_int, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.dojo_step_gather_async.restype = C.c_int
    L.dojo_step_grad_gather_async.argtypes = [vp, vp, op, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.dojo_step_grad_gather_async.restype = C.c_int
    _lib = L
    return L


def _p(a):
    """numpy array / int (device pointer) / None -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


class BatchedStepper:
    """A mechanism bound to one GPU: the handle of include/dojo_b200.h.

    Host arrays are [B, feature] C-contiguous numpy arrays, i.e. exactly the column-major
    [feature x B] Julia matrices of the ABI.  Device buffers are passed as integer pointers
    (``tensor.data_ptr()``) to the *_device methods together with a CUDA stream handle.
    """

    def __init__(self, mech: Mechanism, max_batch: int, device: int = 0):
        self.mech = mech
        self.L = load_library()
        desc, keep = capi.flatten(mech)
        h = C.c_void_p()
        rc = self.L.dojo_create(C.byref(desc), int(device), int(max_batch), C.byref(h))
        if rc != 0:
            msg = self.L.dojo_last_error(None).decode()
            raise RuntimeError(f"dojo_create failed ({rc}): {msg}")
        self.h = h
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.nz = self.L.dojo_num_state(h)
        self.nu = self.L.dojo_num_input(h)
        self.nres = self.L.dojo_num_residual(h)
        self.ngrad = self.L.dojo_num_grad_state(h)
        assert (self.nz, self.nu, self.nres) == (mech.nz, mech.nu, mech.nres)

    def close(self):
        if getattr(self, "h", None):
            self.L.dojo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.L.dojo_last_error(self.h).decode()}")

    def update_params(self, mech: Mechanism):
        """Swap in the parameters of `mech` (same topology: bodies, joints, joint types, limits, contacts) without re-creating
        the handle -- the system-identification loop of examples/system_identification/utilities.jl:41-87."""
        desc, keep = capi.flatten(mech)
        self._check(self.L.dojo_update_params(self.h, C.byref(desc)), "dojo_update_params")
        self.mech = mech

    @property
    def shared_bytes_per_env(self) -> int:
        return self.L.dojo_shared_bytes_per_env(self.h)

    @property
    def launch_config(self) -> dict:
        """Launch configuration chosen by dojo_create (diagnostics; dojo_debug_config is not part of the public header)."""
        out = (C.c_int * 16)()
        self.L.dojo_debug_config.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        self.L.dojo_debug_config(self.h, out)
        keys = ("slots", "slots_grad", "grad_chunk", "arena_bytes", "grad_arena_bytes", "smem_fwd", "smem_grad", "plan_smem_mask", "plan_smem_mask_grad",
                "ls_pair", "phases", "steps", "plan_blob_bytes", "warps_per_env", "ctas_per_sm", "ctas_per_sm_grad")
        return dict(zip(keys, list(out)))

    @property
    def launch_count(self) -> int:
        return int(self.L.dojo_launch_count(self.h))

    # ------------------------------------------------------------------ host buffers
    def step(self, Z, U=None, opts: Optional[capi.DojoSolverOptions] = None, fext=None, flags: int = 0, return_sol: bool = False, out=None):
        """One step! of every environment.  Host arrays in, host arrays out; page-locked arrays (e.g. views of torch pinned
        tensors) are copied from / to directly, pageable ones go through the library's pinned staging buffers.
        out = (Z_next, status, iters) reuses caller-provided result arrays."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        B = Z.shape[0]
        assert Z.shape[1] == self.nz
        U = np.zeros((B, self.nu)) if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        assert U.shape == (B, self.nu)
        if fext is not None:
            fext = np.ascontiguousarray(fext, dtype=np.float64).reshape(B, 6 * self.mech.Nb)
        if out is not None:
            Zn, status, iters = out
            assert Zn.shape == Z.shape and Zn.dtype == np.float64 and Zn.flags.c_contiguous
            assert status.shape == (B,) and status.dtype == np.int32 and iters.shape == (B,) and iters.dtype == np.int32
        else:
            Zn = np.empty_like(Z)
            status = np.zeros(B, dtype=np.int32)
            iters = np.zeros(B, dtype=np.int32)
        sol = np.empty((B, self.nres)) if return_sol else None
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step(self.h, C.byref(o), B, _p(Z), _p(U), _p(fext), _p(Zn), _p(sol), _p(status), _p(iters), flags)
        self._check(rc, "dojo_step")
        return (Zn, status, iters, sol) if return_sol else (Zn, status, iters)

    def step_grad(self, Z, U=None, opts=None, flags: int = 0, out=None):
        """step! + IFT gradients.  Returns (Z_next, Fz [B, 12Nb, 12Nb], Fu [B, 12Nb, nu], status, iters) with Fz[e] = dz'/dz.
        out = (Z_next, Fz_raw [B, 12Nb, 12Nb], Fu_raw [B, nu, 12Nb], status, iters) reuses caller buffers (e.g. page-locked
        ones); the raw arrays hold each environment's Jacobian column-major, the returned Fz / Fu are transposed views."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        B = Z.shape[0]
        U = np.zeros((B, self.nu)) if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        ng = self.ngrad
        if out is not None:
            Zn, Fz, Fu, status, iters = out
            assert Zn.shape == Z.shape and Fz.shape == (B, ng, ng) and Fu.shape == (B, self.nu, ng)
            assert all(a.flags.c_contiguous for a in (Zn, Fz, Fu, status, iters)) and status.dtype == np.int32 and iters.dtype == np.int32
        else:
            Zn = np.empty_like(Z)
            Fz = np.empty((B, ng, ng))   # per env column-major [ng x ng]  ==  Fz[e].T is the Jacobian
            Fu = np.empty((B, self.nu, ng))
            status = np.zeros(B, dtype=np.int32)
            iters = np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_grad(self.h, C.byref(o), B, _p(Z), _p(U), None, _p(Zn), _p(Fz), _p(Fu), _p(status), _p(iters), flags)
        self._check(rc, "dojo_step_grad")
        return Zn, np.transpose(Fz, (0, 2, 1)), np.transpose(Fu, (0, 2, 1)), status, iters

    def step_grad_contact(self, Z, U=None, opts=None):
        """step! + get_contact_gradients (gradients/contact.jl:1-55).  Returns (Z_next, Fz [B, 12Nb, 12Nb], Fu [B, 12Nb, nu],
        Fc [B, 12Nb, 5Ni] = dz'/d[friction_coefficient, contact_radius, contact_origin(3)] per contact, status, iters)."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        B = Z.shape[0]
        U = np.zeros((B, self.nu)) if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        ng, nc = self.ngrad, self.L.dojo_num_contact_data(self.h)
        Zn = np.empty_like(Z)
        Fz, Fu, Fc = np.empty((B, ng, ng)), np.empty((B, self.nu, ng)), np.empty((B, nc, ng))
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_grad_contact(self.h, C.byref(o), B, _p(Z), _p(U), _p(Zn), _p(Fz), _p(Fu), _p(Fc), _p(status), _p(iters))
        self._check(rc, "dojo_step_grad_contact")
        return Zn, np.transpose(Fz, (0, 2, 1)), np.transpose(Fu, (0, 2, 1)), np.transpose(Fc, (0, 2, 1)), status, iters

    def rollout(self, Z0, U=None, T: int = 1, opts=None, record: bool = False):
        Z0 = np.ascontiguousarray(np.atleast_2d(Z0), dtype=np.float64)
        B = Z0.shape[0]
        if U is not None:
            U = np.ascontiguousarray(U, dtype=np.float64)
            assert U.shape == (T, B, self.nu)
        Zf = np.empty_like(Z0)
        traj = np.empty((T, B, self.nz)) if record else None
        st = np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_rollout(self.h, C.byref(o), B, int(T), _p(Z0), _p(U), _p(Zf), _p(traj), _p(st))
        self._check(rc, "dojo_rollout")
        return (Zf, st, traj) if record else (Zf, st)

    # ------------------------------------------------------------------ minimal coordinates
    @property
    def nmin(self) -> int:
        return self.L.dojo_num_minimal(self.h)

    def minimal_to_maximal(self, X):
        """minimal_to_maximal (mechanism/state.jl:9-22), batched: X [B, 2 nu] -> Z [B, 13 Nb]."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        assert X.shape[1] == self.nmin
        Z = np.empty((X.shape[0], self.nz))
        self._check(self.L.dojo_minimal_to_maximal(self.h, X.shape[0], _p(X), _p(Z)), "dojo_minimal_to_maximal")
        return Z

    def maximal_to_minimal(self, Z):
        """maximal_to_minimal (mechanism/state.jl:44-66), batched: Z [B, 13 Nb] -> X [B, 2 nu]."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        assert Z.shape[1] == self.nz
        X = np.empty((Z.shape[0], self.nmin))
        self._check(self.L.dojo_maximal_to_minimal(self.h, Z.shape[0], _p(Z), _p(X)), "dojo_maximal_to_minimal")
        return X

    def step_minimal(self, X, U=None, opts=None, flags: int = 0):
        """step_minimal_coordinates! (simulation/step.jl:42-61), batched.  Returns (X_next, status, iters).  flags =
        DOJO_FLAG_Q1_LITERAL_RETURN: the reference's literal return value in minimal coordinates."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        B = X.shape[0]
        U = None if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        Xn = np.empty_like(X)
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        self._check(self.L.dojo_step_minimal_flags(self.h, C.byref(o), B, _p(X), _p(U), _p(Xn), _p(status), _p(iters), C.c_uint32(flags)), "dojo_step_minimal")
        return Xn, status, iters

    def maximal_to_minimal_jacobian(self, Z):
        """maximal_to_minimal_jacobian (gradients/state.jl:9-56), batched: Z [B, 13 Nb] -> M [B, 2 nu, 12 Nb]."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        assert Z.shape[1] == self.nz
        J = np.empty((Z.shape[0], self.ngrad, self.nmin))  # column-major [2 nu x 12 Nb] per environment
        self._check(self.L.dojo_maximal_to_minimal_jacobian(self.h, Z.shape[0], _p(Z), _p(J)), "dojo_maximal_to_minimal_jacobian")
        return np.transpose(J, (0, 2, 1))

    def minimal_to_maximal_jacobian(self, Z):
        """minimal_to_maximal_jacobian (gradients/state.jl:136-179) evaluated at the maximal states Z [B, 13 Nb]
        (= minimal_to_maximal(X)): N [B, 12 Nb, 2 nu]."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        assert Z.shape[1] == self.nz
        J = np.empty((Z.shape[0], self.nmin, self.ngrad))  # column-major [12 Nb x 2 nu] per environment
        self._check(self.L.dojo_minimal_to_maximal_jacobian(self.h, Z.shape[0], _p(Z), _p(J)), "dojo_minimal_to_maximal_jacobian")
        return np.transpose(J, (0, 2, 1))

    def minimal_gradients(self, X, U=None, opts=None):
        """get_minimal_gradients! (gradients/state.jl:182-217), batched.
        Returns (X_next [B, 2nu], dx'/dx [B, 2nu, 2nu], dx'/du [B, 2nu, nu], status, iters)."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        B = X.shape[0]
        assert X.shape[1] == self.nmin
        U = None if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        Xn = np.empty_like(X)
        Gx = np.empty((B, self.nmin, self.nmin))
        Gu = np.empty((B, self.nu, self.nmin))
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_minimal_gradients(self.h, C.byref(o), B, _p(X), _p(U), _p(Xn), _p(Gx), _p(Gu), _p(status), _p(iters))
        self._check(rc, "dojo_minimal_gradients")
        return Xn, np.transpose(Gx, (0, 2, 1)), np.transpose(Gu, (0, 2, 1)), status, iters

    # ------------------------------------------------------------------ trajectory recording / diagnostics
    def step_record(self, Z, U=None, opts=None):
        """step! + save_to_storage! (simulation/storage.jl:50-67).  Returns (Z_next, storage [B, Nb, 12] = px pq vl wl per body,
        diag [B, 8] = linear momentum, angular momentum about the centre of mass, kinetic, potential energy, status, iters)."""
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=np.float64)
        B = Z.shape[0]
        U = None if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=np.float64)
        Zn = np.empty_like(Z)
        sto, diag = np.empty((B, self.mech.Nb, 12)), np.empty((B, 8))
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_record(self.h, C.byref(o), B, _p(Z), _p(U), _p(Zn), _p(sto), _p(diag), _p(status), _p(iters))
        self._check(rc, "dojo_step_record")
        return Zn, sto, diag, status, iters

    def simulate_record(self, Z0, U=None, T: int = 1, opts=None):
        """simulate!(...; record=true) with open-loop inputs U [T, B, nu].  Returns (Z_final, Z_traj [T, B, 13Nb] = the state
        before every solve (Storage.x, q, v, w), storage [T, B, Nb, 12], diag [T, B, 8], status_any)."""
        Z0 = np.ascontiguousarray(np.atleast_2d(Z0), dtype=np.float64)
        B = Z0.shape[0]
        if U is not None:
            U = np.ascontiguousarray(U, dtype=np.float64)
            assert U.shape == (T, B, self.nu)
        Zf = np.empty_like(Z0)
        traj, sto, diag = np.empty((T, B, self.nz)), np.empty((T, B, self.mech.Nb, 12)), np.empty((T, B, 8))
        st = np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_simulate_record(self.h, C.byref(o), B, int(T), _p(Z0), _p(U), _p(Zf), _p(traj), _p(sto), _p(diag), _p(st))
        self._check(rc, "dojo_simulate_record")
        return Zf, traj, sto, diag, st

    def step_record_device(self, dZ: int, dU: Optional[int], dZn: int, dstorage: int, ddiag: int, B: int, opts=None, dstatus=None, diters=None, stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_record_async(self.h, C.byref(o), int(B), _p(dZ), _p(dU), _p(dZn), _p(dstorage), _p(ddiag), _p(dstatus), _p(diters),
                                           C.c_void_p(int(stream)))
        self._check(rc, "dojo_step_record_async")

    # ------------------------------------------------------------------ environment layer
    def env_sizes(self, spec):
        return self.L.dojo_env_num_state(self.h, C.byref(spec)), self.L.dojo_env_num_action(self.h, C.byref(spec))

    def env_step(self, spec, S, A=None, opts=None):
        """step!(environment, s, a) + get_state + reward + failure test for a batch (DojoEnvironments/src/environments.jl:77-109,
        examples/learning/ant_ars.jl:79-116).  S [B, ns], A [B, na] -> (S_next, reward, done, status, iters)."""
        ns, na = self.env_sizes(spec)
        S = np.ascontiguousarray(np.atleast_2d(S), dtype=np.float64)
        B = S.shape[0]
        assert S.shape[1] == ns
        if A is not None:
            A = np.ascontiguousarray(np.atleast_2d(A), dtype=np.float64)
            assert A.shape == (B, na)
        Sn = np.empty_like(S)
        reward = np.empty(B)
        done, status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_env_step(self.h, C.byref(o), C.byref(spec), B, _p(S), _p(A), _p(Sn), _p(reward), _p(done), _p(status), _p(iters))
        self._check(rc, "dojo_env_step")
        return Sn, reward, done, status, iters

    def env_step_device(self, spec, dS: int, dA: Optional[int], dSn: int, B: int, opts=None, dreward=None, ddone=None, dstatus=None, diters=None,
                        stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_env_step_async(self.h, C.byref(o), C.byref(spec), int(B), _p(dS), _p(dA), _p(dSn), _p(dreward), _p(ddone), _p(dstatus),
                                        _p(diters), C.c_void_p(int(stream)))
        self._check(rc, "dojo_env_step_async")

    def env_rollout(self, spec, S0, A=None, T: int = 1, opts=None):
        """Open-loop rollout of T environment steps on the device (examples/learning/ant_ars.jl:79-116 without the policy):
        S0 [B, ns], A [T, B, na] -> (S_final, return [B], failed [B])."""
        ns, na = self.env_sizes(spec)
        S0 = np.ascontiguousarray(np.atleast_2d(S0), dtype=np.float64)
        B = S0.shape[0]
        if A is not None:
            A = np.ascontiguousarray(A, dtype=np.float64)
            assert A.shape == (T, B, na)
        Sf, ret, failed = np.empty_like(S0), np.empty(B), np.zeros(B, dtype=np.int32)
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_env_rollout(self.h, C.byref(o), C.byref(spec), B, int(T), _p(S0), _p(A), _p(Sf), _p(ret), _p(failed))
        self._check(rc, "dojo_env_rollout")
        return Sf, ret, failed

    def env_policy_rollout(self, spec, S0, Theta, T: int, mean=None, std=None, opts=None, record_states: bool = False):
        """Closed-loop rollout with one linear policy per environment (ARS evaluation): Theta [B, na, ns], a = Theta_e ((s - mean) / std).
        Returns (S_final, return [B], failed [B]) and, with record_states, the states observed before every step [T, B, ns]."""
        ns, na = self.env_sizes(spec)
        S0 = np.ascontiguousarray(np.atleast_2d(S0), dtype=np.float64)
        B = S0.shape[0]
        Theta = np.asarray(Theta, dtype=np.float64)
        assert Theta.shape == (B, na, ns)
        ThetaC = np.ascontiguousarray(Theta.transpose(0, 2, 1))  # column-major [na x ns] per environment
        mean = None if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
        std = None if std is None else np.ascontiguousarray(std, dtype=np.float64)
        Sf, ret, failed = np.empty_like(S0), np.empty(B), np.zeros(B, dtype=np.int32)
        traj = np.empty((T, B, ns)) if record_states else None
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_env_policy_rollout(self.h, C.byref(o), C.byref(spec), B, int(T), _p(S0), _p(ThetaC), _p(mean), _p(std), _p(Sf), _p(ret),
                                            _p(failed), _p(traj))
        self._check(rc, "dojo_env_policy_rollout")
        return (Sf, ret, failed, traj) if record_states else (Sf, ret, failed)

    def env_reset(self, spec, S, s0, mask=None):
        """S[e] = s0 where mask[e] != 0 (all if mask is None).  S / mask: numpy arrays (in place) or device pointers + B."""
        s0 = np.ascontiguousarray(s0, dtype=np.float64)
        if isinstance(S, tuple):  # (device pointer, B)
            dS, B = S
            self._check(self.L.dojo_env_reset(self.h, C.byref(spec), int(B), _p(s0), _p(mask), _p(dS)), "dojo_env_reset")
            return None
        assert S.flags.c_contiguous and S.dtype == np.float64
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.int32)
        self._check(self.L.dojo_env_reset(self.h, C.byref(spec), S.shape[0], _p(s0), _p(mask), _p(S)), "dojo_env_reset")
        return S

    # ------------------------------------------------------------------ device buffers (resident data)
    def step_device(self, dZ: int, dU: Optional[int], dZn: int, B: int, opts=None, dstatus: Optional[int] = None, diters: Optional[int] = None,
                    dsol: Optional[int] = None, dfext: Optional[int] = None, flags: int = 0, stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_async(self.h, C.byref(o), int(B), _p(dZ), _p(dU), _p(dfext), _p(dZn), _p(dsol), _p(dstatus), _p(diters), flags,
                                    C.c_void_p(int(stream)))
        self._check(rc, "dojo_step_async")

    def minimal_to_maximal_device(self, dX: int, dZ: int, B: int, stream: int = 0):
        self._check(self.L.dojo_minimal_to_maximal_async(self.h, int(B), _p(dX), _p(dZ), C.c_void_p(int(stream))), "dojo_minimal_to_maximal_async")

    def maximal_to_minimal_device(self, dZ: int, dX: int, B: int, stream: int = 0):
        self._check(self.L.dojo_maximal_to_minimal_async(self.h, int(B), _p(dZ), _p(dX), C.c_void_p(int(stream))), "dojo_maximal_to_minimal_async")

    def maximal_to_minimal_jacobian_device(self, dZ: int, dJ: int, B: int, stream: int = 0):
        self._check(self.L.dojo_maximal_to_minimal_jacobian_async(self.h, int(B), _p(dZ), _p(dJ), C.c_void_p(int(stream))), "dojo_maximal_to_minimal_jacobian_async")

    def minimal_to_maximal_jacobian_device(self, dZ: int, dJ: int, B: int, stream: int = 0):
        self._check(self.L.dojo_minimal_to_maximal_jacobian_async(self.h, int(B), _p(dZ), _p(dJ), C.c_void_p(int(stream))), "dojo_minimal_to_maximal_jacobian_async")

    def minimal_gradients_device(self, dX: int, dU: Optional[int], dXn: int, dGx: int, dGu: int, B: int, opts=None, dstatus=None, diters=None):
        """device-resident get_minimal_gradients! (synchronises the handle's stream before returning)"""
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_minimal_gradients(self.h, C.byref(o), int(B), _p(dX), _p(dU), _p(dXn), _p(dGx), _p(dGu), _p(dstatus), _p(diters))
        self._check(rc, "dojo_minimal_gradients")

    def rollout_device(self, dZ0: int, dU: Optional[int], dZf: int, B: int, T: int, opts=None, dtraj: Optional[int] = None, dstatus: Optional[int] = None,
                       stream: int = 0):
        """T steps fused in one launch on resident data (U is [T, B, nu] on the device)."""
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_rollout_async(self.h, C.byref(o), int(B), int(T), _p(dZ0), _p(dU), _p(dZf), _p(dtraj), _p(dstatus), C.c_void_p(int(stream)))
        self._check(rc, "dojo_rollout_async")

    def step_grad_device(self, dZ: int, dU: Optional[int], dZn: int, dFz: int, dFu: int, B: int, opts=None, dstatus=None, diters=None, flags: int = 0,
                         stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_grad_async(self.h, C.byref(o), int(B), _p(dZ), _p(dU), None, _p(dZn), _p(dFz), _p(dFu), _p(dstatus), _p(diters), flags,
                                         C.c_void_p(int(stream)))
        self._check(rc, "dojo_step_grad_async")

    # ---- multi-GPU: the exchange of the next states fused into the step (include/dojo_b200.h)
    def gather_create(self, world: int, rank: int, B_local: int):
        g = C.c_void_p()
        self._check(self.L.dojo_gather_create(self.h, int(world), int(rank), int(B_local), C.byref(g)), "dojo_gather_create")
        return g

    def gather_export(self, g) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.L.dojo_gather_export(g, buf), "dojo_gather_export")
        return buf.raw

    def gather_connect(self, g, all_handles: bytes):
        buf = C.create_string_buffer(all_handles, len(all_handles))
        self._check(self.L.dojo_gather_connect(g, buf), "dojo_gather_connect")

    def gather_buffer(self, g) -> int:
        return int(self.L.dojo_gather_buffer(g) or 0)

    def gather_destroy(self, g):
        self.L.dojo_gather_destroy(g)

    def step_gather_device(self, g, dZ: int, dU: Optional[int], dZn: int, B: int, opts=None, dstatus=None, diters=None, flags: int = 0, stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_gather_async(self.h, g, C.byref(o), int(B), _p(dZ), _p(dU), None, _p(dZn), _p(dstatus), _p(diters), flags, C.c_void_p(int(stream)))
        self._check(rc, "dojo_step_gather_async")

    def step_grad_gather_device(self, g, dZ: int, dU: Optional[int], dZn: int, dFz: int, dFu: int, B: int, opts=None, dstatus=None, diters=None, flags: int = 0,
                                stream: int = 0):
        o = opts if opts is not None else capi.solver_options()
        rc = self.L.dojo_step_grad_gather_async(self.h, g, C.byref(o), int(B), _p(dZ), _p(dU), None, _p(dZn), _p(dFz), _p(dFu), _p(dstatus), _p(diters), flags,
                                                C.c_void_p(int(stream)))
        self._check(rc, "dojo_step_grad_gather_async")
