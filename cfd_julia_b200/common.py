"""Host-side mirror of the reference's operator interface for the vortex-merger path.

Same function names, positional arguments, in-place mutation and return values as the Julia
originals (file:line relative to the CFD_Julia checkout):

    fps(nx, ny, dx, dy, u, e, data, data1, f, s, eps=1e-6)          Common.jl:97-125
    vm_rhs(nx, ny, dx, dy, re, w, u, e, data, data1, r, s, f)       Common.jl:132-182
    numerical(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns)             19_NS2D_Vortex_Merger/vm.jl:12-90
    numerical_tgv(nx, ny, nt, dx, dy, dt, re, wn)                   19_NS2D_Vortex_Merger/tgv.jl:13-79
    ps_fft(nx, ny, dx, dy, f, eps=1e-6)                             12_Poisson_Solver_FFT/fft_p.jl:8-42
    vm_ic / exact_tgv / compute_l2norm_bnds                         Common.jl:208-219,234-237, tgv.jl:82-90

Arrays are numpy float64, Fortran-ordered (column-major like Julia); "ghosted" arrays have shape
(nx+2, ny+2).  Everything on the path runs in libvmk.so on the GPU through the C ABI of
include/vmk.h; this module only marshals pointers (the Julia wrapper julia/CommonB200.jl does the same
with ccall).  There is no CPU implementation here: without the CUDA library or a device the calls raise.

Reference behaviours kept: fps/vm_rhs leave the dead scratch arguments u, e, data, data1 untouched
(`e` is rebound and `u` never used in the reference either); vm_rhs writes r's interior only, all of s,
and f = -w interior; numerical mutates wn (all ghosts valid) and returns wn[2:nx+2, 2:ny+2].
Reference behaviour NOT kept: vm.jl:78-86 writes every snapshot to "vm1.txt" because its record index
m is never incremented (hybrid.jl:81 does increment it); numerical() here numbers the files vm1..vm{ns}.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from ._lib import SNAPSHOT_FN, VmkError, VmkLibrary, default_library

__all__ = ["Common", "fps", "vm_rhs", "numerical", "numerical_tgv", "numerical_hybrid", "numerical_ps23", "numerical_ps32", "numerical_ldc", "ps_fft", "vm_ic", "exact_tgv",
           "compute_l2norm_bnds", "write_field", "read_field", "julia_float_str", "Plan", "VmkError"]


def _ptr(a: np.ndarray, shape, what: str):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError(f"{what}: expected a float64 numpy array")
    if tuple(a.shape) != tuple(shape):
        raise IndexError(f"{what}: expected shape {tuple(shape)}, got {tuple(a.shape)}")  # Julia: BoundsError
    if not a.flags.f_contiguous:
        raise TypeError(f"{what}: expected a Fortran-ordered (column-major) array")
    return a.ctypes.data


class Plan:
    """Owns a vmk_plan (device buffers, tables, stream).  The reference has no plan object: Common keeps a
    cache keyed by grid size, as the Julia wrapper does."""

    def __init__(self, lib: VmkLibrary, nx: int, ny: int, rank: int = 0, nranks: int = 1):
        self.lib = lib
        self.nx, self.ny, self.rank, self.nranks = nx, ny, rank, nranks
        h = C.c_void_p()
        lib.check(lib.plan_create_slab(nx, ny, rank, nranks, C.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, lib.plan_destroy, h)

    def close(self):
        self._fin()

    # slab decomposition: exchange the buffer handles of all ranks (one process per GPU, CUDA IPC)
    def attach_peers(self, all_gather_bytes):
        """`all_gather_bytes(b: bytes) -> list[bytes]` returns every rank's blob in rank order
        (e.g. built on torch.distributed.all_gather_object)."""
        nb = self.lib.peer_blob_bytes()
        mine = C.create_string_buffer(nb)
        self.lib.check(self.lib.peer_export(self.handle, C.cast(mine, C.c_void_p)))
        blobs = all_gather_bytes(mine.raw)
        assert len(blobs) == self.nranks and all(len(b) == nb for b in blobs)
        joined = C.create_string_buffer(b"".join(blobs), nb * self.nranks)
        self.lib.check(self.lib.peer_import(self.handle, C.cast(joined, C.c_void_p)))

    # device-resident path
    def upload(self, wn):
        self.lib.check(self.lib.upload(self.handle, _ptr(wn, (self.nx + 2, self.ny + 2), "wn")))

    def step(self, dx, dy, dt, re, nsteps=1):
        self.lib.check(self.lib.step(self.handle, dx, dy, dt, re, nsteps))

    def download(self, wn=None, psi=None):
        sh = (self.nx + 2, self.ny + 2)
        self.lib.check(self.lib.download(self.handle, None if wn is None else _ptr(wn, sh, "wn"),
                                         None if psi is None else _ptr(psi, sh, "psi")))

    def sync(self):
        self.lib.check(self.lib.sync(self.handle))

    def step_elapsed_ms(self) -> float:
        ms = C.c_double()
        self.lib.check(self.lib.step_elapsed_ms(self.handle, C.byref(ms)))
        return ms.value

    def profile_steps(self, dx, dy, dt, re, nsteps):
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        self.lib.check(self.lib.profile_steps(self.handle, dx, dy, dt, re, nsteps, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(("k1", "k2", "k3", "k4"))}

    def profile_tri(self):
        """The recurrence kernels' part of the last profile_steps' "k2" class (csrc/vmk_tri.cuh); see vmk_profile_tri."""
        ms = (C.c_double * 3)()
        n = (C.c_int64 * 3)()
        self.lib.check(self.lib.profile_tri(self.handle, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(("kt_totals", "kt_scan", "kt_solve"))}

    def profile_read(self):
        """Per-class kernel time since the last read (after set_option("profile", 1)); see vmk_profile_read."""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        self.lib.check(self.lib.profile_read(self.handle, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(("k1", "k2", "k3", "k4"))}

    def set_option(self, key: str, value: int):
        self.lib.check(self.lib.set_option(self.handle, key.encode(), value))

    @property
    def launch_count(self) -> int:
        return int(self.lib.launch_count(self.handle))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.device_bytes(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.stream(self.handle) or 0)


class Common:
    """The reference's `Common` module surface for this path, bound to one C-ABI library."""

    def __init__(self, lib: VmkLibrary | None = None):
        self._lib = lib
        self._plans: dict[tuple[int, int], Plan] = {}

    @property
    def lib(self) -> VmkLibrary:
        if self._lib is None:
            self._lib = default_library()
        return self._lib

    def plan(self, nx: int, ny: int) -> Plan:
        key = (int(nx), int(ny))
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = Plan(self.lib, *key)
        return p

    def clear_plans(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()

    # ---- Common.jl:97-125 ---------------------------------------------------------------------
    def fps(self, nx, ny, dx, dy, u, e, data, data1, f, s, eps=1.e-6):
        p = self.plan(nx, ny)
        self.lib.check(self.lib.fps(p.handle, dx, dy, _ptr(f, (nx, ny), "f"), _ptr(s, (nx + 2, ny + 2), "s"), eps))

    # ---- Common.jl:132-182 --------------------------------------------------------------------
    def vm_rhs(self, nx, ny, dx, dy, re, w, u, e, data, data1, r, s, f):
        p = self.plan(nx, ny)
        g = (nx + 2, ny + 2)
        self.lib.check(self.lib.rhs(p.handle, dx, dy, re, _ptr(w, g, "w"), _ptr(r, g, "r"), _ptr(s, g, "s"),
                                    None if f is None else _ptr(f, (nx, ny), "f")))

    # ---- fft_p.jl:8-42 ------------------------------------------------------------------------
    def ps_fft(self, nx, ny, dx, dy, f, eps=1.e-6):
        p = self.plan(nx, ny)
        u = np.zeros((nx, ny), order="F")
        self.lib.check(self.lib.ps_fft(p.handle, dx, dy, _ptr(f, (nx + 1, ny + 1), "f"), u.ctypes.data, eps))
        return u

    # ---- vm.jl:12-90 --------------------------------------------------------------------------
    def numerical(self, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
        """snapshot(k, ut) is called every nt // ns steps with ut = wn[2:nx+2, 2:ny+2] (vm.jl:78-80); if
        `outdir` is given the reference's text files are written there (numbered, see module docstring)."""
        if ns <= 0 or nt // ns == 0:
            raise ZeroDivisionError("mod(k, nt ÷ ns) with nt ÷ ns == 0")  # Julia: DivideError at vm.jl:78
        freq = nt // ns
        return self._numerical(nx, ny, nt, dx, dy, dt, re, wn, freq, x, y, snapshot, outdir)

    # ---- tgv.jl:13-79 -------------------------------------------------------------------------
    def numerical_tgv(self, nx, ny, nt, dx, dy, dt, re, wn):
        return self._numerical(nx, ny, nt, dx, dy, dt, re, wn, 0, None, None, None, None)

    # ---- 20_NS2D_Hybrid_Solver/hybrid.jl:14-90 ---------------------------------------------------
    def numerical_hybrid(self, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
        """The hybrid RK3 / Crank-Nicolson solver's `numerical`: same arguments as vm.jl's, wn is only read, returns
        ut = real(ifft(wf)) as an (nx+1) x (ny+1) array.  snapshot(k, ut) / the text files vm{m}.txt every nt // ns
        steps (hybrid.jl:71-86; this script does increment its record index)."""
        return self._numerical_spectral(self.lib.hybrid_numerical, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot,
                                        outdir)

    # ---- 22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl:13-89 ---------------------------
    def numerical_ps23(self, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
        """The pseudo-spectral solver's `numerical` (2/3 truncation rule): same arguments as vm.jl's, wn is only read,
        returns ut = real(ifft(wf)) as an (nx+1) x (ny+1) array; snapshot(k, ut) / vm{m}.txt every nt // ns steps
        (pseudospectral_23_rule.jl:69-86)."""
        return self._numerical_spectral(self.lib.ps23_numerical, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot,
                                        outdir)

    # ---- 21_NS2D_PseudoSpectral_32_Rule/pseudospectral_32_rule.jl:13-89 ---------------------------
    def numerical_ps32(self, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
        """The pseudo-spectral solver's `numerical` (3/2 padding rule): same arguments and return value as
        numerical_ps23."""
        return self._numerical_spectral(self.lib.ps32_numerical, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot,
                                        outdir)

    def _numerical_spectral(self, entry, nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot, outdir):
        if ns <= 0 or nt // ns == 0:
            raise ZeroDivisionError("mod(k, nt ÷ ns) with nt ÷ ns == 0")  # Julia: DivideError at hybrid.jl:71
        freq = nt // ns
        p = self.plan(nx, ny)
        ut = np.zeros((nx + 1, ny + 1), order="F")
        rec = [0]

        failed = []  # an exception cannot cross the C frames of the trampoline: keep the first, re-raise on return

        def _snap(k, _ptr, _user):
            if failed:
                return
            try:
                rec[0] += 1
                if snapshot is not None:
                    snapshot(int(k), ut)
                if outdir is not None:
                    write_field(f"{outdir}/vm{rec[0]}.txt", x, y, ut, lib=self.lib)
            except BaseException as ex:  # noqa: BLE001
                failed.append(ex)

        want = snapshot is not None or outdir is not None
        cb = SNAPSHOT_FN(_snap) if want else SNAPSHOT_FN()
        self.lib.check(entry(p.handle, int(nt), dx, dy, dt, re, _ptr(wn, (nx + 2, ny + 2), "wn"), ut.ctypes.data,
                             freq if want else 0, cb, None))
        if failed:
            raise failed[0]
        return ut

    # ---- 18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl:59-117 -----------------------------------
    def numerical_ldc(self, nx, ny, nt, dx, dy, dt, re, wn, sn, rms):
        """The lid-driven cavity script's `numerical`: same arguments, wn and sn ((nx+1) x (ny+1) node arrays) and
        rms[0:nt] are mutated in place, nothing is returned."""
        p = self.plan(2 * nx, 2 * ny)  # the sine transform runs as a periodic transform of the odd extension
        if not isinstance(rms, np.ndarray) or rms.dtype != np.float64 or rms.size < nt or not rms.flags.c_contiguous:
            raise IndexError("rms: expected a contiguous float64 array of at least nt elements")  # Julia: BoundsError
        self.lib.check(self.lib.ldc_numerical(p.handle, nx, ny, int(nt), dx, dy, dt, re,
                                              _ptr(wn, (nx + 1, ny + 1), "wn"), _ptr(sn, (nx + 1, ny + 1), "sn"),
                                              rms.ctypes.data))

    def _numerical(self, nx, ny, nt, dx, dy, dt, re, wn, freq, x, y, snapshot, outdir):
        p = self.plan(nx, ny)
        g = (nx + 2, ny + 2)
        out = np.zeros((nx + 1, ny + 1), order="F")
        rec = [0]

        failed = []  # see above: the reference would abort at the failing open(); so does this, after the C call

        def _snap(k, ptr, _user):
            if failed:
                return
            try:
                rec[0] += 1
                ut = wn[1:nx + 2, 1:ny + 2]
                if snapshot is not None:
                    snapshot(int(k), ut)
                if outdir is not None:
                    write_field(f"{outdir}/vm{rec[0]}.txt", x, y, ut, lib=self.lib)
            except BaseException as ex:  # noqa: BLE001
                failed.append(ex)

        want = freq > 0 and (snapshot is not None or outdir is not None)
        cb = SNAPSHOT_FN(_snap) if want else SNAPSHOT_FN()
        self.lib.check(self.lib.numerical(p.handle, int(nt), dx, dy, dt, re, _ptr(wn, g, "wn"), out.ctypes.data,
                                          freq if want else 0, cb, None))
        if failed:
            raise failed[0]
        return out


def write_field(path, x, y, ut, lib=None):
    """The reference's text dump "x y w" with j outer, i inner (vm.jl:81-85, 132-136, 142-146), written by the
    library's native writer (csrc/vmk_io.hpp) with Julia's own Float64 formatting, so the files are byte-identical to
    what the scripts write for the same values."""
    lib = lib or default_library()
    ut = np.asarray(ut)
    nx1, ny1 = ut.shape
    xa = np.ascontiguousarray(np.asarray(x, dtype=np.float64)[:nx1])
    ya = np.ascontiguousarray(np.asarray(y, dtype=np.float64)[:ny1])
    ua = np.asfortranarray(ut, dtype=np.float64)
    if xa.size != nx1 or ya.size != ny1:
        raise IndexError("write_field: x / y shorter than the field")  # Julia: BoundsError
    lib.check(lib.write_field(os.fsencode(path), xa.ctypes.data, ya.ctypes.data, ua.ctypes.data, nx1, ny1))


def read_field(path, nx, ny, lib=None):
    """plotting.jl:14-28: readdlm + reshape.  Returns (x[0:nx+1], y[0:ny+1], w (nx+1) x (ny+1))."""
    lib = lib or default_library()
    n = (nx + 1) * (ny + 1)
    cols = [np.zeros(n) for _ in range(3)]
    rows = C.c_int64(0)
    lib.check(lib.read_field(os.fsencode(path), cols[0].ctypes.data, cols[1].ctypes.data, cols[2].ctypes.data, n,
                             C.byref(rows)))
    if rows.value != n:
        raise ValueError(f"{path}: {rows.value} rows, expected {(nx + 1)}*{(ny + 1)}")  # Julia: DimensionMismatch in reshape
    xx = cols[0][:nx + 1].copy()
    yy = cols[1].reshape((nx + 1, ny + 1), order="F")[0, :].copy()
    return xx, yy, cols[2].reshape((nx + 1, ny + 1), order="F")


def julia_float_str(v, lib=None) -> str:
    """Julia's print(::Float64) (what "$(x)" interpolates), from the library's formatter."""
    lib = lib or default_library()
    buf = C.create_string_buffer(40)
    n = lib.print_float64(float(v), buf)
    return buf.raw[:n].decode()


# ---- setup helpers of the callers (host-side, not on the hot path) ----------------------------------
def vm_ic(nx, ny, x, y, w):
    """Common.jl:208-219: two Gaussian vortices on w[2:nx+2, 2:ny+2] (ghost fill is the caller's, vm.jl:121-128)."""
    sigma = np.pi
    xc1, yc1 = np.pi - np.pi / 4., np.pi
    xc2, yc2 = np.pi + np.pi / 4., np.pi
    X = np.asarray(x)[:nx + 1, None]
    Y = np.asarray(y)[None, :ny + 1]
    w[1:nx + 2, 1:ny + 2] = (np.exp(-sigma * ((X - xc1)**2 + (Y - yc1)**2)) +
                            np.exp(-sigma * ((X - xc2)**2 + (Y - yc2)**2)))


def exact_tgv(nx, ny, x, y, time, re):
    """tgv.jl:82-90."""
    nq = 4.
    X = np.asarray(x)[:nx + 1, None]
    Y = np.asarray(y)[None, :ny + 1]
    return np.asfortranarray(2 * nq * np.cos(nq * X) * np.cos(nq * Y) * np.exp(-2 * nq**2 * time / re))


def compute_l2norm_bnds(nx, ny, r):
    """Common.jl:234-237."""
    return float(np.sqrt(np.sum(np.asarray(r)[:nx + 1, :ny + 1]**2) / ((nx + 1) * (ny + 1))))


_common = Common()


def fps(nx, ny, dx, dy, u, e, data, data1, f, s, eps=1.e-6):
    return _common.fps(nx, ny, dx, dy, u, e, data, data1, f, s, eps)


def vm_rhs(nx, ny, dx, dy, re, w, u, e, data, data1, r, s, f):
    return _common.vm_rhs(nx, ny, dx, dy, re, w, u, e, data, data1, r, s, f)


def ps_fft(nx, ny, dx, dy, f, eps=1.e-6):
    return _common.ps_fft(nx, ny, dx, dy, f, eps)


def numerical(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
    return _common.numerical(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot, outdir)


def numerical_tgv(nx, ny, nt, dx, dy, dt, re, wn):
    return _common.numerical_tgv(nx, ny, nt, dx, dy, dt, re, wn)


def numerical_hybrid(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
    return _common.numerical_hybrid(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot, outdir)


def numerical_ps23(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
    return _common.numerical_ps23(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot, outdir)


def numerical_ps32(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot=None, outdir=None):
    return _common.numerical_ps32(nx, ny, nt, dx, dy, dt, re, x, y, wn, ns, snapshot, outdir)


def numerical_ldc(nx, ny, nt, dx, dy, dt, re, wn, sn, rms):
    return _common.numerical_ldc(nx, ny, nt, dx, dy, dt, re, wn, sn, rms)


def plan(nx, ny) -> Plan:
    return _common.plan(nx, ny)
