"""Generates the committed golden fixtures in tests/golden/*.npz.

The reference is Julia (not installed here, and its FFTW.jl/Unroll/Utils deps are un-vendored),
so it cannot generate vectors itself.  These fixtures are outputs of the numpy restatement
oracle/oracle_np.py (which reproduces the reference's five recorded numbers, order.jl:13) on
deterministic inputs.  Re-run:  python tests/golden/make_golden.py   (--f-rows: only the fixtures of SURVEY 8f's rows)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_np as onp  # noqa: E402


def main():
    # (1) fps on uniform(-1,1) noise, 32x32, L = 2 pi (exercises every mode incl. the eps quirk row/col)
    rng = np.random.default_rng(0)
    n = 32
    dx, dy, x, y = onp.grid(n, n)
    f = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
    s = np.zeros((n + 2, n + 2), order="F")
    onp.fps(n, n, dx, dy, f, s)
    np.savez_compressed(os.path.join(HERE, "fps_noise_32.npz"), f=f, s=s, dx=dx, dy=dy)

    # (2) vm_rhs on the vortex-merger IC, 64x64, Re=1000
    n = 64
    dx, dy, x, y = onp.grid(n, n)
    w = np.zeros((n + 2, n + 2), order="F")
    onp.vm_ic(n, n, x, y, w)
    r = np.zeros_like(w)
    s = np.zeros_like(w)
    f = np.zeros((n, n), order="F")
    onp.vm_rhs(n, n, dx, dy, 1000., w, r, s, f)
    np.savez_compressed(os.path.join(HERE, "vm_rhs_64.npz"), w=w, r=r, s=s, f=f, dx=dx, dy=dy, re=1000.)

    # (3) numerical: vm.jl defaults but 64x64, 25 steps, dt=.01
    wn = w.copy(order="F")
    out, s = onp.numerical(n, n, 25, dx, dy, .01, 1000., wn)
    np.savez_compressed(os.path.join(HERE, "vm_numerical_64_25.npz"), w0=w, out=out, wn=wn, s=s,
                        dx=dx, dy=dy, dt=.01, re=1000., nt=25)

    # (4) tgv.jl defaults: 64x64, Re=10, dt=.01, 100 steps
    wn = np.zeros((n + 2, n + 2), order="F")
    wn[1:n + 2, 1:n + 2] = onp.exact_tgv(n, n, x, y, 0., 10.)
    wn[0, :] = wn[n, :]
    wn[:, 0] = wn[:, n]
    w0 = wn.copy(order="F")
    out, s = onp.numerical(n, n, 100, dx, dy, .01, 10., wn)
    ue = onp.exact_tgv(n, n, x, y, 1., 10.)
    np.savez_compressed(os.path.join(HERE, "tgv_64_100.npz"), w0=w0, out=out, dx=dx, dy=dy, dt=.01, re=10., nt=100,
                        l2=onp.compute_l2norm_bnds(n, n, out - ue), mx=np.max(np.abs(out - ue)))


def main_f_rows():
    """the hybrid solver, both pseudo-spectral rules and the lid-driven cavity (numpy restatements of
    hybrid.jl, pseudospectral_23_rule.jl, pseudospectral_32_rule.jl, lid_driven_cavity.jl)."""
    # (5) spectral-space solvers: vm_ic + 5 % noise, 64x64, Re=1000, dt=.01, 20 steps
    n, nt, dt, re = 64, 20, .01, 1000.
    dx, dy, x, y = onp.grid(n, n)
    w = np.zeros((n + 2, n + 2), order="F")
    onp.vm_ic(n, n, x, y, w)
    rng = np.random.default_rng(11)
    w[1:n + 1, 1:n + 1] += .05 * rng.uniform(-1, 1, (n, n))
    w[n + 1, :] = w[1, :]
    w[:, n + 1] = w[:, 1]
    w[0, :] = w[n, :]
    w[:, 0] = w[:, n]
    np.savez_compressed(os.path.join(HERE, "spectral_64_20.npz"), w0=w, dx=dx, dy=dy, dt=dt, re=re, nt=nt,
                        hybrid=onp.hybrid_numerical(n, n, nt, dx, dy, dt, re, w),
                        ps23=onp.ps_numerical(23, n, n, nt, dx, dy, dt, re, w),
                        ps32=onp.ps_numerical(32, n, n, nt, dx, dy, dt, re, w))

    # (6) lid-driven cavity: 32x32 cells, Re=100, noisy start, 20 steps
    n, nt, re = 32, 20, 100.
    dx = 1. / n
    dt = min(.001, 0.2 * dx * dx * re)
    rng = np.random.default_rng(12)
    wn = np.asfortranarray(rng.uniform(-1, 1, (n + 1, n + 1)))
    sn = np.zeros((n + 1, n + 1), order="F")
    sn[1:n, 1:n] = 1e-2 * rng.uniform(-1, 1, (n - 1, n - 1))
    w0, s0 = wn.copy(order="F"), sn.copy(order="F")
    rms = np.zeros(nt)
    onp.ldc_numerical(n, n, nt, dx, dx, dt, re, wn, sn, rms)
    np.savez_compressed(os.path.join(HERE, "ldc_32_20.npz"), w0=w0, s0=s0, wn=wn, sn=sn, rms=rms, dx=dx, dt=dt, re=re, nt=nt)


if __name__ == "__main__":
    if "--f-rows" not in sys.argv:
        main()
    main_f_rows()
