"""Golden vectors computed BY THE REFERENCE'S OWN CODE (not by this repo's oracles).

The reference ships two Python twins of script 19 (the vortex-merger step):

  19_NS2D_Vortex_Merger/Python_Vectorized/fdm_vortex_merge_vectorized.py
        :31-78 fps (same eps quirk :46-47), :84-91 bc, :98-129 rhs (Arakawa + Laplacian), :135-150 vm_ic,
        :221-256 the RK3 loop (nt steps)
  19_NS2D_Vortex_Merger/Python/fdm_vortex_merger.py
        the same with explicit loops (:28-78 fps, :95-131 rhs, :208-252 RK3 loop, nt-1 steps)

Both are executed here UNMODIFIED (read from /root/reference, compiled and exec'd as they are) with three shims
for what this image lacks:
  * `pyfftw`      -> numpy.fft behind pyfftw.FFTW / pyfftw.empty_aligned (FFTW itself is absent; any FP64 FFT
                     agrees with it to ~4e-16, and pyfftw's inverse is normalised by default like numpy's)
  * `matplotlib`  -> no-op stub (the scripts plot after the run)
  * `input.txt`   -> generated (the file is git-ignored upstream; format :152-165 of either script)

Outputs: tests/golden/ref_py_*.npz -- initial field, final vorticity and streamfunction, and rhs/fps samples on a
white-noise field, all converted to the Julia scripts' ghosted (nx+2) x (ny+2) column-major layout:

    Python twin: (nx+3) x (ny+3) arrays, index p <-> grid point p-1 (0 and nx+2 are ghosts, nx+1 is the periodic
                 duplicate);  Julia: index i (1-based) <-> grid point i-2.  Hence J = P[0:nx+2, 0:ny+2].

Known differences between the twins and vm.jl/Common.jl (all at rounding level; the tests document the resulting
tolerance): no wavenumber wrap (kx = hx*i for all i, :42-44 -- cos is periodic); `gg` multiplies each of j1, j2, j3;
`lap/re` instead of pre-divided aa, bb (vectorized twin); stage 3 uses (1/3)*w instead of w/3; x from linspace; the
rhs is also evaluated on the duplicate point.

This script needs /root/reference and therefore runs in the build container only; the .npz files travel.
Re-run:  python tests/golden/make_ref_fixtures.py
"""
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CFD_JULIA_REFERENCE", "/root/reference")
VEC = os.path.join(REF, "19_NS2D_Vortex_Merger", "Python_Vectorized", "fdm_vortex_merge_vectorized.py")
LOOP = os.path.join(REF, "19_NS2D_Vortex_Merger", "Python", "fdm_vortex_merger.py")


def _pyfftw_shim():
    m = types.ModuleType("pyfftw")

    def empty_aligned(shape, dtype="float64", **kw):
        return np.empty(shape, dtype=dtype)

    class FFTW:
        def __init__(self, a, b, axes=(-1,), direction="FFTW_FORWARD", **kw):
            self.a, self.b, self.axes, self.direction = a, b, tuple(axes), direction

        def __call__(self, data=None, out=None, normalise_idft=True):
            if data is not None:
                self.a[...] = data
            if self.direction == "FFTW_FORWARD":
                self.b[...] = np.fft.fftn(self.a, axes=self.axes)
            else:
                self.b[...] = np.fft.ifftn(self.a, axes=self.axes)
                if not normalise_idft:
                    self.b[...] *= np.prod([self.a.shape[ax] for ax in self.axes])
            return self.b

    m.empty_aligned = empty_aligned
    m.FFTW = FFTW
    return m


def _matplotlib_shim():
    mpl = types.ModuleType("matplotlib")
    plt = mock.MagicMock(name="matplotlib.pyplot")
    plt.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    mpl.pyplot = plt
    mpl.ticker = mock.MagicMock(name="matplotlib.ticker")
    return {"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.ticker": mpl.ticker}


def run_reference_script(path, nd, nt, re, dt, ns=1):
    """exec the unmodified reference script; returns its module globals (w0, w, s, fps, rhs, bc, ...)."""
    with open(path) as fh:
        code = compile(fh.read(), path, "exec")
    mods = {"pyfftw": _pyfftw_shim(), **_matplotlib_shim()}
    cwd = os.getcwd()
    g = {"__name__": "__ref_twin__", "__file__": path}
    with tempfile.TemporaryDirectory() as tmp, mock.patch.dict(sys.modules, mods):
        # :152-165: nd, nt, re, dt, ns, isolver, isc, ich (must be 19), ipr, ndc -- one value per line
        with open(os.path.join(tmp, "input.txt"), "w") as fh:
            fh.write("\n".join(str(v) for v in (nd, nt, re, dt, ns, 3, 0, 19, 1, nd)) + "\n")
        os.chdir(tmp)
        try:
            exec(code, g)
        finally:
            os.chdir(cwd)
    return g


def to_julia(p, n):
    """(n+3) x (n+3) twin array -> ghosted (n+2) x (n+2) column-major Julia layout."""
    return np.asfortranarray(p[0:n + 2, 0:n + 2])


def from_julia(j, n, bc):
    p = np.zeros((n + 3, n + 3))
    p[0:n + 2, 0:n + 2] = j
    return bc(n, n, p)


def make(path, tag, nd, nt_script, nsteps, re, dt, with_samples):
    g = run_reference_script(path, nd, nt_script, re, dt)
    n = nd
    out = dict(n=n, nsteps=nsteps, re=re, dt=dt, dx=float(g["dx"]), dy=float(g["dy"]),
               w0=to_julia(g["w0"], n), w=to_julia(g["w"], n), s=to_julia(g["s"], n))
    if with_samples:
        # the twin's own fps / rhs on a white-noise periodic field (every Fourier mode, incl. the eps row / column)
        rng = np.random.default_rng(2024)
        wj = np.zeros((n + 2, n + 2), order="F")
        wj[1:n + 1, 1:n + 1] = rng.uniform(-1, 1, (n, n))
        wj[n + 1, :] = wj[1, :]
        wj[:, n + 1] = wj[:, 1]
        wj[0, :] = wj[n, :]
        wj[:, 0] = wj[:, n]
        wp = from_julia(wj, n, g["bc"])
        sp = g["bc"](n, n, g["fps"](n, n, g["dx"], g["dy"], -wp))
        rp = g["rhs"](n, n, g["dx"], g["dy"], re, wp, sp)
        out.update(noise_w=wj, noise_s=to_julia(sp, n), noise_r=np.asfortranarray(rp[1:n + 1, 1:n + 1]))
    fn = os.path.join(HERE, f"ref_py_{tag}.npz")
    np.savez_compressed(fn, **out)
    print(f"{fn}: n={n} steps={nsteps} max|w|={np.abs(out['w']).max():.6f} max|s|={np.abs(out['s']).max():.6f}")


def main():
    make(VEC, "vec_64_50", 64, 50, 50, 1000., .01, True)
    make(VEC, "vec_128_20", 128, 20, 20, 1000., .01, True)
    # the loop-based twin runs range(1, nt): nt - 1 steps
    make(LOOP, "loop_32_10", 32, 11, 10, 1000., .01, False)


if __name__ == "__main__":
    main()
