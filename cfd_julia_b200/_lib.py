"""ctypes binding of the C ABI declared in include/vmk.h.

The same binding class serves the product library (libvmk.so, symbols vmk_*) and -- in the CPU test
suite only -- the host emulation build (tests/emul/libvmk_emul.so, symbols vmke_*).  The package
itself only ever loads libvmk.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

SNAPSHOT_FN = C.CFUNCTYPE(None, C.c_int64, C.POINTER(C.c_double), C.c_void_p)
BARRIER_FN = C.CFUNCTYPE(None, C.c_void_p)

VMK_OK, VMK_ESIZE, VMK_ECUDA, VMK_EARG, VMK_ESTATE = 0, 1, 2, 3, 4

# name -> (restype, argtypes); every symbol of include/vmk.h
_P = C.c_void_p
_D = C.c_double
_PROTOS = {
    "version": (C.c_int, []),
    "last_error": (C.c_char_p, []),
    "plan_create": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(_P)]),
    "plan_create_slab": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "plan_create_on": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "plan_destroy": (C.c_int, [_P]),
    "peer_blob_bytes": (C.c_size_t, []),
    "peer_export": (C.c_int, [_P, _P]),
    "peer_import": (C.c_int, [_P, _P]),
    "peer_attach_local": (C.c_int, [_P, C.POINTER(_P)]),
    "barrier_hook": (C.c_int, [_P, BARRIER_FN, _P]),
    "fps": (C.c_int, [_P, _D, _D, _P, _P, _D]),
    "ps_fft": (C.c_int, [_P, _D, _D, _P, _P, _D]),
    "rhs": (C.c_int, [_P, _D, _D, _D, _P, _P, _P, _P]),
    "numerical": (C.c_int, [_P, C.c_int64, _D, _D, _D, _D, _P, _P, C.c_int64, SNAPSHOT_FN, _P]),
    "hybrid_numerical": (C.c_int, [_P, C.c_int64, _D, _D, _D, _D, _P, _P, C.c_int64, SNAPSHOT_FN, _P]),
    "ps23_numerical": (C.c_int, [_P, C.c_int64, _D, _D, _D, _D, _P, _P, C.c_int64, SNAPSHOT_FN, _P]),
    "ps32_numerical": (C.c_int, [_P, C.c_int64, _D, _D, _D, _D, _P, _P, C.c_int64, SNAPSHOT_FN, _P]),
    "print_float64": (C.c_int, [_D, C.c_char_p]),
    "write_field": (C.c_int, [C.c_char_p, _P, _P, _P, C.c_int64, C.c_int64]),
    "read_field": (C.c_int, [C.c_char_p, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "ldc_numerical": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _D, _D, _D, _D, _P, _P, _P]),
    "upload": (C.c_int, [_P, _P]),
    "step": (C.c_int, [_P, _D, _D, _D, _D, C.c_int64]),
    "download": (C.c_int, [_P, _P, _P]),
    "sync": (C.c_int, [_P]),
    "stream": (_P, [_P]),
    "step_elapsed_ms": (C.c_int, [_P, C.POINTER(_D)]),
    "profile_steps": (C.c_int, [_P, _D, _D, _D, _D, C.c_int64, C.POINTER(_D), C.POINTER(C.c_int64)]),
    "profile_read": (C.c_int, [_P, C.POINTER(_D), C.POINTER(C.c_int64)]),
    "profile_tri": (C.c_int, [_P, C.POINTER(_D), C.POINTER(C.c_int64)]),
    "launch_count": (C.c_int64, [_P]),
    "set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "device_bytes": (C.c_int64, [_P]),
}
SYMBOLS = tuple("vmk_" + n for n in _PROTOS)


class VmkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vmk error {code}: {msg}")
        self.code = code


class VmkLibrary:
    """A loaded C-ABI library.  `prefix` selects the symbol family (vmk_ = product)."""

    def __init__(self, path: str, prefix: str = "vmk_"):
        self.path = path
        self.prefix = prefix
        self.cdll = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(self.cdll, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int):
        if rc:
            raise VmkError(rc, self.last_error().decode(errors="replace"))


_default = None


def default_library() -> VmkLibrary:
    """The product library.  Built on first use if the sources are newer; never replaced by anything else."""
    global _default
    if _default is None:
        so = os.environ.get("VMK_LIB")  # an explicitly chosen build of the same CUDA library (tuning experiments)
        if not so:
            so = _build.SO
            if _build.stale():
                so = _build.build()
        if not os.path.exists(so):
            raise RuntimeError("cfd_julia_b200: libvmk.so is missing and could not be built")
        _default = VmkLibrary(so, "vmk_")
    return _default
