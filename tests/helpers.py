"""Shared input builders and comparison helpers for the parity tests."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def grid(n, L=2 * np.pi):
    """vm.jl:94-117 / tgv.jl:93-115: dx, dy and the nx+1 / ny+1 node coordinates."""
    dx = L / n
    x = dx * np.arange(n + 1)
    return dx, dx, x, x.copy()


def ghost_fill(n, a):
    """Common.jl:138-146 order."""
    a[n + 1, :] = a[1, :]
    a[:, n + 1] = a[:, 1]
    a[0, :] = a[n, :]
    a[:, 0] = a[:, n]
    return a


def noise_field(n, seed=0):
    """uniform(-1,1) interior with periodic ghosts: exercises every Fourier mode."""
    rng = np.random.default_rng(seed)
    w = np.zeros((n + 2, n + 2), order="F")
    w[1:n + 1, 1:n + 1] = rng.uniform(-1, 1, (n, n))
    return ghost_fill(n, w)


def vm_field(n):
    """vm_ic + main()'s ghost fill (Common.jl:208-219, vm.jl:121-128)."""
    from cfd_julia_b200.common import vm_ic
    dx, dy, x, y = grid(n)
    w = np.zeros((n + 2, n + 2), order="F")
    vm_ic(n, n, x, y, w)
    w[0, :] = w[n, :]
    w[:, 0] = w[:, n]
    w[n + 1, :] = w[1, :]
    w[:, n + 1] = w[:, 1]
    return w


def tgv_field(n, re=10.):
    """tgv.jl:117-123."""
    from cfd_julia_b200.common import exact_tgv
    dx, dy, x, y = grid(n)
    w = np.zeros((n + 2, n + 2), order="F")
    w[1:n + 2, 1:n + 2] = exact_tgv(n, n, x, y, 0., re)
    w[0, :] = w[n, :]
    w[:, 0] = w[:, n]
    return w


def stable_dt(n, re):
    """Diffusive RK3 limit dt <= 2.5127 re dx^2 / 8, with a 2x margin, capped at the scripts' .01."""
    dx = 2 * np.pi / n
    return min(.01, 0.5 * 2.5127 * re * dx * dx / 8.)


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def mms(n):
    """fft_p.jl:44-82 manufactured source/solution on the unit square, (n+1) x (n+1)."""
    dx = 1. / n
    x = dx * np.arange(n + 1)
    X, Y = x[:, None], x[None, :]
    km, c2 = 16, -8 * np.pi**2
    c1 = (1. / km)**2
    ue = np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) + c1 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y)
    f = c2 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) + c2 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y)
    return dx, np.asfortranarray(f), np.asfortranarray(ue)


# 13_Poisson_Solver_FFT_Spectral/specrtral_vs_FDM/order.jl:13
ORDER_JL_FFT_FDM = {32: .0015607100315532957, 64: .0005987381110678801, 128: .00014313734718665358,
                    256: 3.549617203207291e-5, 512: 8.865373334924762e-6}
