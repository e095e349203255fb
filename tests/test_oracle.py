"""Pins the two CPU oracles: against the reference's only recorded outputs (order.jl:13),
against the analytic Taylor-Green solution (tgv.jl:87), against each other, and against the
committed golden fixtures."""
import os

import numpy as np
import pytest

import parity_cases as pc

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


def tgv_spectral_known_answer(n, nt, dt, re, nq=4.):
    """Known answer for the three spectral-space scripts (hybrid.jl, pseudospectral_23_rule.jl, pseudospectral_32_rule.jl):
    the Taylor-Green field w0 = 2 nq cos(nq x) cos(nq y) is an eigenfunction of the Laplacian with psi = w / (2 nq^2), so
    its Jacobian vanishes identically (also the discrete Arakawa one: J(a, c a) = 0) and every RK3/Crank-Nicolson stage
    (hybrid.jl:40-66) just multiplies the four modes (+-nq, +-nq) by (1 - d_s)/(1 + d_s), d_s = alpha_s (dt/2) k2 / re,
    k2 = 2 nq^2 (hx = 1 on the 2 pi box)."""
    x = 2 * np.pi / n * np.arange(n + 1)
    w0 = 2 * nq * np.cos(nq * x)[:, None] * np.cos(nq * x)[None, :]
    k2 = 2 * nq**2
    fac = 1.
    for alpha in (8. / 15., 2. / 15., 1. / 3.):
        d = alpha * (.5 * dt * k2 / re)
        fac *= (1. - d) / (1. + d)
    return w0, w0 * fac**nt


@pytest.mark.parametrize("which", ["hybrid", "ps23", "ps32"])
def test_spectral_scripts_tgv_known_answer(oracle_np, which):
    """pins the numpy restatements of the three spectral-space scripts on a closed-form answer (no recorded reference
    output exists for them)"""
    n, nt, dt, re = 64, 25, .01, 10.
    w0, exact = tgv_spectral_known_answer(n, nt, dt, re)
    wn = np.zeros((n + 2, n + 2), order="F")
    wn[1:n + 2, 1:n + 2] = w0
    dx = 2 * np.pi / n
    if which == "hybrid":
        ut = oracle_np.hybrid_numerical(n, n, nt, dx, dx, dt, re, wn)
    else:
        ut = oracle_np.ps_numerical(int(which[2:]), n, n, nt, dx, dx, dt, re, wn)
    assert np.linalg.norm(ut - exact) / np.linalg.norm(exact) < 1e-12
    assert 0.2 < np.abs(ut).max() / np.abs(w0).max() < 0.5  # the decay over 25 steps is substantial: exp(-2 nq^2 t / re)


def test_pseudospectral_rules_dealias(oracle_np):
    """the two de-aliasing rules must agree (to rounding) on a field whose products stay inside both retained bands, and
    differ from the aliased product otherwise: checks the band edges of both restatements against each other"""
    n = 128  # 2/3 rule: retained band -42 .. 41
    dx = 2 * np.pi / n
    k2 = oracle_np.wavespace(n, n, dx, dx)
    rng = np.random.default_rng(3)
    kk = np.abs(np.fft.fftfreq(n, 1. / n))
    band = (kk[:, None] <= 20) & (kk[None, :] <= 20)  # products reach |k| <= 40 < 42: inside both retained bands
    wf = np.fft.fft2(rng.uniform(-1, 1, (n, n))) * band
    wf[0, 0] = 0
    j23 = oracle_np.ps23_jacobian(n, n, dx, dx, wf, k2)
    j32 = oracle_np.ps32_jacobian(n, n, dx, dx, wf, k2)
    assert np.linalg.norm(j23 - j32) / np.linalg.norm(j32) < 1e-13


# ---- pinned on the reference's own code: its Python twins of script 19 run unmodified (make_ref_fixtures.py) ----
@pytest.mark.parametrize("tag", pc.REF_PY)
def test_ref_py_oracle_c(oracle_c, tag):
    pc.check_ref_py_oracle(oracle_c, tag)


@pytest.mark.parametrize("tag", pc.REF_PY)
def test_ref_py_oracle_np(oracle_np, tag):
    pc.check_ref_py_oracle(oracle_np, tag)


# ---- design models of the recurrence forms of the solve along j (tests/models/tri_model.py) against the oracle ----
@pytest.mark.parametrize("n,k0,units", [(256, 16, 7), (512, 32, 37), (1024, 64, 9), (1024, 16, 148)])
def test_fused_form_model(oracle_np, n, k0, units):
    """numpy model of the fused form with the kernels' blocking and operations (Horner in 1/r for the block totals,
    q recursion for the carry from the left): blocks of up to 114 rows, ragged block lengths, more units than pairs
    would allow at the small end -- all within 1e-14 of the FFT x FFT oracle (the B200 measured 5e-16 at 8192^2)"""
    import importlib.util
    from helpers import ROOT
    spec = importlib.util.spec_from_file_location("tri_model", os.path.join(ROOT, "tests", "models", "tri_model.py"))
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)
    dx = dy = 2 * np.pi / n
    f = np.random.default_rng(n + units).standard_normal((n, n))
    ref = oracle_np.poisson(n, n, dx, dy, f)
    got = tm.poisson_fused(n, dx, dy, f, k0, units)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-14
    got2 = tm.poisson_tri(n, dx, dy, f, k0, 32)
    assert np.linalg.norm(got2 - ref) / np.linalg.norm(ref) < 1e-14
