"""Pins the two CPU oracles: against the reference's only recorded outputs (order.jl:13),
against the analytic Taylor-Green solution (tgv.jl:87), against each other, and against the
committed golden fixtures."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mms(n):
    """fft_p.jl:44-82 manufactured source on the unit square."""
    dx = 1. / n
    x = dx * np.arange(n + 1)
    X, Y = x[:, None], x[None, :]
    km, c2 = 16, -8 * np.pi**2
    c1 = (1. / km)**2
    ue = np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) + c1 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y)
    f = c2 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) + c2 * np.sin(km * 2 * np.pi * X) * np.sin(km * 2 * np.pi * Y)
    return dx, np.asfortranarray(f), np.asfortranarray(ue)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512])
def test_order_jl_golden_numpy(oracle_np, n):
    l2, _ = oracle_np.fft_p_case(n)
    assert abs(l2 - oracle_np.ORDER_JL_FFT_FDM[n]) / oracle_np.ORDER_JL_FFT_FDM[n] < 1e-11


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512])
def test_order_jl_golden_c(oracle_c, oracle_np, n):
    dx, f, ue = _mms(n)
    un = np.zeros_like(f)
    un[:n, :n] = oracle_c.ps_fft(n, n, dx, dx, f)
    un[n, :] = un[0, :]
    un[:, n] = un[:, 0]
    l2 = oracle_c.compute_l2norm_bnds(n, n, np.asfortranarray(un - ue))
    g = oracle_np.ORDER_JL_FFT_FDM[n]
    assert abs(l2 - g) / g < 1e-11


def test_c_vs_numpy_rhs_noise(oracle_c, oracle_np):
    n = 64
    rng = np.random.default_rng(1)
    dx, dy, x, y = oracle_np.grid(n, n)
    w = np.zeros((n + 2, n + 2), order="F")
    w[1:n + 1, 1:n + 1] = rng.uniform(-1, 1, (n, n))
    oracle_np.ghost_fill(n, n, w)
    outs = []
    for o in (oracle_c, oracle_np):
        r = np.zeros_like(w)
        s = np.zeros_like(w)
        f = np.zeros((n, n), order="F")
        o.vm_rhs(n, n, dx, dy, 1000., w, r, s, f)
        outs.append((r, s, f))
    for a, b in zip(*outs):
        assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(b)


def test_c_vs_golden(oracle_c):
    g = np.load(os.path.join(GOLD, "vm_numerical_64_25.npz"))
    n = 64
    wn = np.asfortranarray(g["w0"].copy())
    out, s = oracle_c.numerical(n, n, int(g["nt"]), float(g["dx"]), float(g["dy"]), float(g["dt"]), float(g["re"]), wn)
    assert np.linalg.norm(out - g["out"]) <= 1e-13 * np.linalg.norm(g["out"])
    assert np.linalg.norm(s - g["s"]) <= 1e-13 * np.linalg.norm(g["s"])
    g = np.load(os.path.join(GOLD, "fps_noise_32.npz"))
    s = np.zeros((34, 34), order="F")
    oracle_c.fps(32, 32, float(g["dx"]), float(g["dy"]), np.asfortranarray(g["f"]), s)
    assert np.linalg.norm(s - g["s"]) <= 1e-13 * np.linalg.norm(g["s"])


def test_tgv_defaults_analytic(oracle_c, oracle_np):
    """tgv.jl defaults (64^2, Re=10, dt=.01, 100 steps): survey-probe values L2=6.9131011113e-3, max=1.3616714310e-2."""
    n = 64
    dx, dy, x, y = oracle_np.grid(n, n)
    wn = np.zeros((n + 2, n + 2), order="F")
    wn[1:n + 2, 1:n + 2] = oracle_c.exact_tgv(n, n, x, y, 0., 10.)
    wn[0, :] = wn[n, :]
    wn[:, 0] = wn[:, n]
    out, _ = oracle_c.numerical(n, n, 100, dx, dy, .01, 10., wn)
    ue = oracle_c.exact_tgv(n, n, x, y, 1., 10.)
    l2 = oracle_c.compute_l2norm_bnds(n, n, np.asfortranarray(out - ue))
    assert abs(l2 - 6.9131011113e-3) < 1e-12
    assert abs(np.max(np.abs(out - ue)) - 1.3616714310e-2) < 1e-11


def test_eps_quirk_matters(oracle_np):
    """A 'clean' solver (kx[1]=0) differs from the reference arithmetic by > 1e-10 (SURVEY finding 2)."""
    n = 128
    dx, dy, x, y = oracle_np.grid(n, n)
    w = np.zeros((n + 2, n + 2), order="F")
    oracle_np.vm_ic(n, n, x, y, w)
    f = -w[1:n + 1, 1:n + 1]
    a = oracle_np.poisson(n, n, dx, dy, f, 1e-6)
    d = oracle_np.divisor(n, n, dx, dy, 0.0)
    d[0, 0] = 1.0
    e = np.fft.fft2(f)
    e[0, 0] = 0
    b = np.real(np.fft.ifft2(e / d))
    assert np.linalg.norm(a - b) / np.linalg.norm(a) > 1e-10


def test_divisor_tables_bitwise(oracle_c, oracle_np):
    n = 128
    dx, dy, _, _ = oracle_np.grid(n, n)
    aa, bbcos, cccos = oracle_c.divisor_tables(n, n, dx, dy)
    d = (aa + bbcos[:, None]) + cccos[None, :]
    # hermitian symmetry that makes the real-input FFT legitimate
    idx = (-np.arange(n)) % n
    assert np.array_equal(d, d[idx][:, idx])
