"""inquistr_b200 -- B200 (sm_100a) implementation of the `inquiSTR call` hot path.

The product is the C-ABI library inquistr_b200/lib/libinqcall.so (include/inqcall.h) and the
C++ host `inquistr-b200 call`. This package is the thin ctypes binding the tests and the
benchmark use; it has no CPU fallback and raises if the CUDA library is missing.
"""
from .api import Context, GenotypeResult, InqError, Stats, free_pinned, load_library, pinned_empty  # noqa: F401

__version__ = "0.1.0"
