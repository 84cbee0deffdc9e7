"""BatchedStepper look-alike on top of the kernel emulation -- TEST INFRASTRUCTURE.

tests/test_gpu_logic_on_emulation.py runs the bodies of `-m gpu` tests on the CPU with this class in place of
dojo_jl_b200.solver.BatchedStepper: the step / gradient kernels run on CPU fibers (tests/hostemu), the thread-per-environment kernels
around them (coordinate maps, environment pre / post, storage) are the device headers compiled for the host (tests/hostcheck),
composed in the order the C-ABI launches them.  It checks the LOGIC and the THRESHOLDS of those tests without a GPU (no FMA
contraction here: rounding differs from the device in the last bits).  It is not a fallback: the product never imports it."""
import numpy as np

from dojo_jl_b200 import capi
from hostcheck.harness import HostCheck
from .harness import HostEmu


class EmuStepper:
    def __init__(self, mech, max_batch, device=0):
        self.mech, self.max_batch = mech, int(max_batch)
        self.em, self.hc = HostEmu(mech), HostCheck(mech)
        self.nz, self.nu, self.nres, self.ngrad = mech.nz, mech.nu, mech.nres, 12 * mech.Nb
        self.slots = 4

    # ---- step / gradients / rollout
    def _u(self, U, B):
        return np.zeros((B, self.nu)) if U is None else np.ascontiguousarray(np.atleast_2d(U), dtype=float)

    def step(self, Z, U=None, opts=None, fext=None, flags=0, return_sol=False, out=None):
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        Zn, st, it, sol = self.em.step(Z, self._u(U, Z.shape[0]), opts, fext=fext, flags=flags, slots=self.slots)
        return (Zn, st, it, sol) if return_sol else (Zn, st, it)

    def step_grad(self, Z, U=None, opts=None, flags=0, out=None):
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        return self.em.step_grad(Z, self._u(U, Z.shape[0]), opts, slots=2, slots_grad=1 if self.mech.Nb > 13 else 2, flags=flags)

    def rollout(self, Z0, U=None, T=1, opts=None, record=False):
        Z0 = np.ascontiguousarray(np.atleast_2d(Z0), dtype=float)
        out = self.em.step(Z0, U, opts, T=T, slots=self.slots, record=record)
        return (out[0], out[1], out[4]) if record else (out[0], out[1])

    # ---- coordinate maps
    def minimal_to_maximal(self, X):
        return self.hc.minimal_to_maximal(np.atleast_2d(X))

    def maximal_to_minimal(self, Z):
        return self.hc.maximal_to_minimal(np.atleast_2d(Z))

    def step_minimal(self, X, U=None, opts=None):
        Zn, st, it = self.step(self.minimal_to_maximal(X), U, opts)
        return self.maximal_to_minimal(Zn), st, it

    def minimal_gradients(self, X, U=None, opts=None):
        Z = self.minimal_to_maximal(X)
        Zn, Fz, Fu, st, it = self.step_grad(Z, U, opts)
        Gx, Gu = self.hc.minimal_gradients(Z, Zn, Fz, Fu)
        return self.maximal_to_minimal(Zn), Gx, Gu, st, it

    # ---- recording
    def step_record(self, Z, U=None, opts=None):
        Z = np.ascontiguousarray(np.atleast_2d(Z), dtype=float)
        U = self._u(U, Z.shape[0])
        Zn, st, it, sol = self.em.step(Z, U, opts, slots=self.slots)
        sto, diag = self.hc.storage(Z, Zn, U, sol)
        return Zn, sto, diag, st, it

    # ---- environment layer
    def env_sizes(self, spec):
        return 2 * self.nu + (self.mech.Ni if spec.contact_obs else 0), self.nu - spec.n_unactuated

    def env_step(self, spec, S, A=None, opts=None):
        Z, U = self.hc.env_pre(spec, S, A)
        Zn, st, it, sol = self.em.step(Z, U, opts, slots=self.slots)
        Sn, reward, done = self.hc.env_post(spec, S, A, Zn, sol)
        return Sn, reward, done, st, it

    def env_reset(self, spec, S, s0, mask=None):
        for e in range(S.shape[0]):
            if mask is None or mask[e]:
                S[e] = s0
        return S
