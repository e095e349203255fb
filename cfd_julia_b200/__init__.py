"""cfd_julia_b200 -- B200 (sm_100a) implementation of CFD_Julia's 2-D vortex-merger solver step.

The product is cfd_julia_b200/libvmk.so (hand-written CUDA kernels behind the C ABI of include/vmk.h);
`common` mirrors the reference's Julia functions on numpy arrays and `julia/CommonB200.jl` does the same
for Julia callers.  Importing the package does not need a GPU; calling into it does.
"""
from . import _build  # noqa: F401
from .common import (Common, Plan, VmkError, compute_l2norm_bnds, exact_tgv, fps, julia_float_str,  # noqa: F401
                     numerical, numerical_hybrid, numerical_ldc, numerical_ps23, numerical_ps32, numerical_tgv, plan, ps_fft,
                     read_field, vm_ic, vm_rhs, write_field)
from ._lib import SYMBOLS, VmkLibrary, default_library  # noqa: F401

__version__ = "0.1.0"


def build(force: bool = False) -> str:
    """Compile libvmk.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    return _build.build(force=force)
