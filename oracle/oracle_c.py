"""ctypes binding of oracle/liboracle.so (the C restatement in vm_oracle.c).

TEST INFRASTRUCTURE, NOT THE PRODUCT -- see the header of vm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="F_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_fps.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_double]
        L.orc_ps_fft.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_double]
        L.orc_vm_rhs.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.orc_numerical.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                    _dp, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_vm_ic.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_vm_ic.restype = None
        L.orc_exact_tgv.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_exact_tgv.restype = None
        L.orc_l2norm_bnds.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_l2norm_bnds.restype = C.c_double
        L.orc_ghost_fill.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_ghost_fill.restype = None
        L.orc_divisor_tables.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_double), _dp, _dp]
        L.orc_fft2.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def _chk(rc, what):
    if rc:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


def fps(nx, ny, dx, dy, f, s, eps=1.e-6):
    _chk(lib().orc_fps(nx, ny, dx, dy, f, s, eps), "fps")


def ps_fft(nx, ny, dx, dy, f, eps=1.e-6):
    u = np.zeros((nx, ny), order="F")
    _chk(lib().orc_ps_fft(nx, ny, dx, dy, f, u, eps), "ps_fft")
    return u


def vm_rhs(nx, ny, dx, dy, re, w, r, s, f):
    _chk(lib().orc_vm_rhs(nx, ny, dx, dy, re, w, r, s, f), "vm_rhs")


def numerical(nx, ny, nt, dx, dy, dt, re, wn):
    """Returns (wn[2:nx+2,2:ny+2], psi ghosted of the last rhs call); mutates wn."""
    out = np.zeros((nx + 1, ny + 1), order="F")
    s = np.zeros((nx + 2, ny + 2), order="F")
    _chk(lib().orc_numerical(nx, ny, nt, dx, dy, dt, re, wn, out.ctypes.data, s.ctypes.data, 0, None, None),
         "numerical")
    return out, s


def vm_ic(nx, ny, x, y, w):
    lib().orc_vm_ic(nx, ny, np.asfortranarray(x), np.asfortranarray(y), w)


def exact_tgv(nx, ny, x, y, time, re):
    ue = np.zeros((nx + 1, ny + 1), order="F")
    lib().orc_exact_tgv(nx, ny, np.asfortranarray(x), np.asfortranarray(y), time, re, ue)
    return ue


def compute_l2norm_bnds(nx, ny, r):
    return float(lib().orc_l2norm_bnds(nx, ny, np.asfortranarray(r)))


def divisor_tables(nx, ny, dx, dy, eps=1.e-6):
    aa = C.c_double()
    bbcos = np.zeros(nx)
    cccos = np.zeros(ny)
    _chk(lib().orc_divisor_tables(nx, ny, dx, dy, eps, C.byref(aa), bbcos, cccos), "divisor_tables")
    return aa.value, bbcos, cccos


def fft2(a, sign=-1):
    """In-place 2-D DFT of a Fortran-ordered complex128 array."""
    assert a.dtype == np.complex128 and a.flags.f_contiguous
    _chk(lib().orc_fft2(a.shape[0], a.shape[1], a.ctypes.data, sign), "fft2")
    return a


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))
