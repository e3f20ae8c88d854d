"""cvsteer_b200 -- B200 (sm_100a) implementation of cvsteer's steerable-filter hot path.

Layout:
    csrc/            CUDA kernels + the C ABI (include/cvsteer_c.h) -> libcvsteer_b200.so
    capi.py          ctypes binding of the C ABI
    filters.py       SteerableFiltersG2 / SteerableFiltersG4: the reference's class surface over host arrays
    batch.py         device-resident batch + pyramid path over torch CUDA tensors (plumbing only)
    multi.py         frame / row-band sharding over torch.distributed
"""
from . import capi  # noqa: F401
from .filters import SteerableFiltersG2, SteerableFiltersG4, make_taps_g2, make_taps_g4  # noqa: F401

__all__ = ["capi", "SteerableFiltersG2", "SteerableFiltersG4", "make_taps_g2", "make_taps_g4"]
