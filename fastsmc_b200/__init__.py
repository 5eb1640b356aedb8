"""fastsmc_b200 — B200-native (sm_100a) implementation of FastSMC's IBD-detection hot path.

The compute path is hand-written CUDA behind the C ABI of include/fastsmc_b200.h
(fastsmc_b200/lib/libfastsmc_b200.so); there is no CPU fallback.
"""
from . import _native  # noqa: F401
from ._native import Context, FastSMCError, pack_haplotypes  # noqa: F401

__version__ = "0.1.0"
