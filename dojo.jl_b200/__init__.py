"""dojo.jl_b200 -- B200-native batched differentiable-physics step for Dojo mechanisms.

Only the hot path lives here: csrc/ (CUDA kernels + C-ABI) and the
host-side mirror of the reference interface.
"""
from .mechanism import Mechanism, get_mechanism, pack_maximal_state, unpack_maximal_state  # noqa: F401
