"""Parity checks written once and run against two libraries: the product (libvmk.so on the GPU, `-m gpu`)
and the host emulation of the same kernel bodies (tests/emul, CPU suite).  `cm` is a
cfd_julia_b200.common.Common bound to the library under test, `oc` the C oracle.

Tolerance: north_star asks for rel-L2 <= 1e-10 on vorticity and streamfunction; single calls are held
to 1e-12 here (observed ~5e-16), multi-step runs to 1e-10.
"""
import numpy as np

from helpers import (GOLD, ORDER_JL_FFT_FDM, ghost_fill, grid, mms, noise_field, rel_l2, stable_dt, tgv_field,
                     vm_field)

TOL_CALL = 1e-12
TOL_RUN = 1e-10


def check_fps_noise(cm, oc, n, seed=0):
    dx, dy, _, _ = grid(n)
    f = np.asfortranarray(np.random.default_rng(seed).uniform(-1, 1, (n, n)))
    s = np.full((n + 2, n + 2), 7.0, order="F")
    ref = np.full((n + 2, n + 2), 7.0, order="F")
    cm.fps(n, n, dx, dy, None, None, None, None, f, s)
    oc.fps(n, n, dx, dy, f, ref)
    assert rel_l2(s[1:n + 1, 1:n + 1], ref[1:n + 1, 1:n + 1]) < TOL_CALL
    # the reference writes the interior only (Common.jl:123): ghosts keep the caller's values
    for edge in (s[0, :], s[n + 1, :], s[:, 0], s[:, n + 1]):
        assert np.all(edge == 7.0)


def check_rhs(cm, oc, w, re=1000.):
    n = w.shape[0] - 2
    dx, dy, _, _ = grid(n)
    r = np.full_like(w, 3.0)
    s = np.zeros_like(w)
    f = np.zeros((n, n), order="F")
    r2 = np.full_like(w, 3.0)
    s2 = np.zeros_like(w)
    f2 = np.zeros((n, n), order="F")
    w_before = w.copy(order="F")
    cm.vm_rhs(n, n, dx, dy, re, w, None, None, None, None, r, s, f)
    oc.vm_rhs(n, n, dx, dy, re, w_before, r2, s2, f2)
    assert np.array_equal(w, w_before)
    assert np.array_equal(f, f2)  # f = -w is exact
    assert rel_l2(s, s2) < TOL_CALL  # all cells incl. ghosts (Common.jl:138-146)
    assert rel_l2(r[1:n + 1, 1:n + 1], r2[1:n + 1, 1:n + 1]) < TOL_CALL
    for edge in (r[0, :], r[n + 1, :], r[:, 0], r[:, n + 1]):  # r's ghosts are never written
        assert np.all(edge == 3.0)


def check_numerical(cm, oc, w0, nt, dt, re, tol=TOL_RUN):
    n = w0.shape[0] - 2
    dx, dy, _, _ = grid(n)
    wa = w0.copy(order="F")
    wb = w0.copy(order="F")
    out = cm.numerical_tgv(n, n, nt, dx, dy, dt, re, wa)
    ref, psi_ref = oc.numerical(n, n, nt, dx, dy, dt, re, wb)
    assert out.shape == (n + 1, n + 1)
    assert rel_l2(out, ref) < tol
    assert rel_l2(wa, wb) < tol  # wn mutated in place, ghosts valid
    assert np.array_equal(out, wa[1:n + 2, 1:n + 2])
    # north_star: vorticity AND streamfunction.  The plan still holds the run's state: psi of the last rhs
    # evaluation (stage 3's input), which is what the reference's `s` holds on return (vm.jl:60, Common.jl:123,138-146)
    psi = np.zeros_like(wa)
    cm.plan(n, n).download(None, psi)
    assert rel_l2(psi, psi_ref) < tol
    return out


def check_golden(cm):
    g = np.load(f"{GOLD}/fps_noise_32.npz")
    s = np.zeros((34, 34), order="F")
    cm.fps(32, 32, float(g["dx"]), float(g["dy"]), None, None, None, None, np.asfortranarray(g["f"]), s)
    assert rel_l2(s, g["s"]) < TOL_CALL
    g = np.load(f"{GOLD}/vm_rhs_64.npz")
    w = np.asfortranarray(g["w"])
    r = np.zeros_like(w)
    s = np.zeros_like(w)
    f = np.zeros((64, 64), order="F")
    cm.vm_rhs(64, 64, float(g["dx"]), float(g["dy"]), float(g["re"]), w, None, None, None, None, r, s, f)
    assert rel_l2(r, g["r"]) < TOL_CALL and rel_l2(s, g["s"]) < TOL_CALL and np.array_equal(f, g["f"])
    g = np.load(f"{GOLD}/vm_numerical_64_25.npz")
    wn = np.asfortranarray(g["w0"].copy())
    out = cm.numerical_tgv(64, 64, int(g["nt"]), float(g["dx"]), float(g["dy"]), float(g["dt"]), float(g["re"]), wn)
    assert rel_l2(out, g["out"]) < TOL_RUN and rel_l2(wn, g["wn"]) < TOL_RUN
    g = np.load(f"{GOLD}/tgv_64_100.npz")
    wn = np.asfortranarray(g["w0"].copy())
    out = cm.numerical_tgv(64, 64, int(g["nt"]), float(g["dx"]), float(g["dy"]), float(g["dt"]), float(g["re"]), wn)
    assert rel_l2(out, g["out"]) < TOL_RUN


# ---- vectors computed by the REFERENCE'S OWN code (its Python twins of script 19, tests/golden/make_ref_fixtures.py)
REF_PY = ("vec_64_50", "vec_128_20", "loop_32_10")
# The twins differ from vm.jl / Common.jl at rounding level only (no wavenumber wrap, `gg` inside j1..j3, lap/re,
# (1/3)*w instead of w/3, rhs also on the duplicate point): observed 4e-16 .. 4e-15 for both oracles and the kernels.
TOL_REF_PY = 1e-12


def check_ref_py(numerical, vm_rhs, tag):
    """`numerical(n, nt, dx, dy, dt, re, wn)` mutates wn; `vm_rhs(n, dx, dy, re, w, r, s, f)` fills r, s, f -- bound to
    an oracle or to a library under test.  Same initial field as the reference run (identical initial conditions),
    then vorticity and streamfunction after the fixture's number of steps, and one rhs / Poisson solve on white noise."""
    g = np.load(f"{GOLD}/ref_py_{tag}.npz")
    n, nt = int(g["n"]), int(g["nsteps"])
    dx, dy, dt, re = float(g["dx"]), float(g["dy"]), float(g["dt"]), float(g["re"])
    assert np.array_equal(vm_field(n), g["w0"])  # this repo's vm_ic + ghost fill == the twin's, bit for bit
    wn = np.asfortranarray(g["w0"].copy())
    numerical(n, nt, dx, dy, dt, re, wn)
    assert rel_l2(wn, g["w"]) < TOL_REF_PY
    r, s, f = np.zeros_like(wn), np.zeros_like(wn), np.zeros((n, n), order="F")
    vm_rhs(n, dx, dy, re, wn, r, s, f)  # the twin's `s` on exit is fps(-w_final) + periodic fill (:252-253)
    assert rel_l2(s, g["s"]) < TOL_REF_PY
    if "noise_w" in g:
        w = np.asfortranarray(g["noise_w"].copy())
        vm_rhs(n, dx, dy, re, w, r, s, f)
        assert rel_l2(s, g["noise_s"]) < TOL_REF_PY
        assert rel_l2(r[1:n + 1, 1:n + 1], g["noise_r"]) < TOL_REF_PY


def check_ref_py_lib(cm, tag):
    check_ref_py(lambda n, nt, dx, dy, dt, re, wn: cm.numerical_tgv(n, n, nt, dx, dy, dt, re, wn),
                 lambda n, dx, dy, re, w, r, s, f: cm.vm_rhs(n, n, dx, dy, re, w, None, None, None, None, r, s, f), tag)


def check_ref_py_oracle(o, tag):
    check_ref_py(lambda n, nt, dx, dy, dt, re, wn: o.numerical(n, n, nt, dx, dy, dt, re, wn),
                 lambda n, dx, dy, re, w, r, s, f: o.vm_rhs(n, n, dx, dy, re, w, r, s, f), tag)


def check_golden_f_rows(cm):
    """committed fixtures of SURVEY 8f's rows (tests/golden/make_golden.py --f-rows)"""
    g = np.load(f"{GOLD}/spectral_64_20.npz")
    n, nt = 64, int(g["nt"])
    dx, dy, x, y = grid(n)
    w0 = np.asfortranarray(g["w0"].copy())
    args = (n, n, nt, float(g["dx"]), float(g["dy"]), float(g["dt"]), float(g["re"]), x, y, w0, 1)
    assert rel_l2(cm.numerical_hybrid(*args), g["hybrid"]) < TOL_RUN
    assert rel_l2(cm.numerical_ps23(*args), g["ps23"]) < TOL_RUN
    assert rel_l2(cm.numerical_ps32(*args), g["ps32"]) < TOL_RUN
    g = np.load(f"{GOLD}/ldc_32_20.npz")
    n, nt = 32, int(g["nt"])
    wn, sn = np.asfortranarray(g["w0"].copy()), np.asfortranarray(g["s0"].copy())
    rms = np.zeros(nt)
    cm.numerical_ldc(n, n, nt, float(g["dx"]), float(g["dx"]), float(g["dt"]), float(g["re"]), wn, sn, rms)
    assert rel_l2(wn, g["wn"]) < TOL_RUN and rel_l2(sn, g["sn"]) < TOL_RUN
    assert np.max(np.abs(rms / g["rms"] - 1.)) < 1e-8


def check_order_jl(cm, n):
    """The reference's only recorded outputs: fft_p.jl L2 errors hard-coded at order.jl:13."""
    from cfd_julia_b200.common import compute_l2norm_bnds
    dx, f, ue = mms(n)
    un = np.zeros_like(f)
    un[:n, :n] = cm.ps_fft(n, n, dx, dx, f)
    un[n, :] = un[0, :]  # fft_p.jl:95-98
    un[:, n] = un[:, 0]
    l2 = compute_l2norm_bnds(n, n, un - ue)
    assert abs(l2 - ORDER_JL_FFT_FDM[n]) / ORDER_JL_FFT_FDM[n] < 1e-10


def check_tgv_defaults(cm):
    """tgv.jl defaults: 64^2, Re=10, dt=.01, 100 steps; survey-probe values of the printed errors."""
    from cfd_julia_b200.common import compute_l2norm_bnds, exact_tgv
    n = 64
    dx, dy, x, y = grid(n)
    wn = tgv_field(n)
    out = cm.numerical_tgv(n, n, 100, dx, dy, .01, 10., wn)
    ue = exact_tgv(n, n, x, y, 1., 10.)
    assert abs(compute_l2norm_bnds(n, n, out - ue) - 6.9131011113e-3) < 1e-11
    assert abs(np.max(np.abs(out - ue)) - 1.3616714310e-2) < 1e-10


def check_snapshots(cm, oc, tmp_path):
    """vm.jl:78-86 cadence: every nt // ns steps; text format "x y w", j outer."""
    n, nt, ns = 32, 10, 5
    dx, dy, x, y = grid(n)
    wn = vm_field(n)
    ref_w = wn.copy(order="F")
    seen = []
    out = cm.numerical(n, n, nt, dx, dy, .01, 1000., x, y, wn, ns, snapshot=lambda k, ut: seen.append((k, ut.copy())),
                       outdir=str(tmp_path))
    assert [k for k, _ in seen] == [2, 4, 6, 8, 10]
    ref2, _ = oc.numerical(n, n, 2, dx, dy, .01, 1000., ref_w)
    assert rel_l2(seen[0][1], ref2) < TOL_RUN
    assert np.array_equal(seen[-1][1], out)
    rows = np.loadtxt(tmp_path / "vm5.txt")
    assert rows.shape == ((n + 1) * (n + 1), 3)
    assert np.array_equal(rows[:, 2].reshape(n + 1, n + 1).T, out)  # j outer, i inner
    assert np.array_equal(rows[:n + 1, 0], x)


def check_errors(cm):
    import pytest
    from cfd_julia_b200.common import VmkError
    with pytest.raises(VmkError) as e:
        cm.plan(48, 48)  # not a power of two
    assert e.value.code == 1
    with pytest.raises(VmkError):
        cm.plan(64, 128)  # ky = kx aliasing of the reference needs nx == ny
    with pytest.raises(VmkError):
        cm.plan(16, 16)
    with pytest.raises(IndexError):
        cm.fps(32, 32, .1, .1, None, None, None, None, np.zeros((32, 32), order="F"), np.zeros((32, 32), order="F"))
    from cfd_julia_b200.common import Plan
    p = Plan(cm.lib, 32, 32)
    with pytest.raises(VmkError) as e:
        p.step(.1, .1, .01, 100., 1)  # step before upload
    assert e.value.code == 4
    p.close()


def check_hybrid(cm, onp, n, nt, dt=None, re=1000., ns=1, tol=TOL_RUN):
    """20_NS2D_Hybrid_Solver/hybrid.jl `numerical` against the numpy restatement (oracle_np.hybrid_numerical)."""
    dx, dy, x, y = grid(n)
    w = vm_field(n) + 0.05 * noise_field(n, 3)
    dt = stable_dt(n, re) if dt is None else dt
    wb = w.copy(order="F")
    snaps_ref, snaps = [], []
    ref = onp.hybrid_numerical(n, n, nt, dx, dy, dt, re, w, nt // ns, lambda k, ut: snaps_ref.append((k, ut.copy())))
    ut = cm.numerical_hybrid(n, n, nt, dx, dy, dt, re, x, y, w, ns, snapshot=lambda k, u: snaps.append((k, u.copy())))
    assert ut.shape == (n + 1, n + 1) and np.array_equal(w, wb)  # wn is only read (hybrid.jl:24)
    assert rel_l2(ut, ref) < tol
    assert np.array_equal(ut[n, :], ut[0, :]) and np.array_equal(ut[:, n], ut[:, 0])  # hybrid.jl:75-78
    assert abs(ut[:n, :n].mean()) < 1e-12  # the mean mode is dropped (hybrid.jl:27)
    assert [k for k, _ in snaps] == [k for k, _ in snaps_ref]
    for (_, a), (_, b) in zip(snaps, snaps_ref):
        assert rel_l2(a, b) < tol
    return ut


def check_ps23(cm, onp, n, nt, dt=None, re=1000., ns=1, tol=TOL_RUN, noise=0.05, rule=23):
    """22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl (rule=23) or 21_NS2D_PseudoSpectral_32_Rule/
    pseudospectral_32_rule.jl (rule=32) `numerical` against the literal numpy restatement (oracle_np.ps_numerical, full
    complex spectra, numpy's C2C transforms, 1.5n-point transforms for the 3/2 rule).  The noisy start puts energy into
    every mode, including the half-weighted mode at the edge of the (asymmetric) retained band and the Nyquist lines."""
    dx, dy, x, y = grid(n)
    w = vm_field(n) + noise * noise_field(n, 5)
    dt = stable_dt(n, re) if dt is None else dt
    wb = w.copy(order="F")
    snaps_ref, snaps = [], []
    ref = onp.ps_numerical(rule, n, n, nt, dx, dy, dt, re, w, nt // ns, lambda k, ut: snaps_ref.append((k, ut.copy())))
    numerical = cm.numerical_ps23 if rule == 23 else cm.numerical_ps32
    ut = numerical(n, n, nt, dx, dy, dt, re, x, y, w, ns, snapshot=lambda k, u: snaps.append((k, u.copy())))
    assert ut.shape == (n + 1, n + 1) and np.array_equal(w, wb)  # wn is only read (:22)
    assert rel_l2(ut, ref) < tol
    assert np.array_equal(ut[n, :], ut[0, :]) and np.array_equal(ut[:, n], ut[:, 0])  # :73-76
    assert abs(ut[:n, :n].mean()) < 1e-12  # the mean mode is dropped (:27) and only re-enters at rounding level
    assert [k for k, _ in snaps] == [k for k, _ in snaps_ref]
    for (_, a), (_, b) in zip(snaps, snaps_ref):
        assert rel_l2(a, b) < tol
    return ut


def check_ps32(cm, onp, n, nt, **kw):
    return check_ps23(cm, onp, n, nt, rule=32, **kw)


def check_spectral_tgv(cm, which, n, nt, dt=.01, re=10., tol=1e-12):
    """Closed-form answer for the spectral-space solvers (hybrid / ps23 / ps32), independent of any oracle: the
    Taylor-Green field is an eigenfunction whose Jacobian vanishes, so each RK3/CN stage multiplies it by
    (1 - d_s)/(1 + d_s), d_s = alpha_s (dt/2) 2 nq^2 / re (hybrid.jl:29-66; tests/test_oracle.py pins the numpy
    restatements on the same answer)."""
    nq = 4.
    dx, dy, x, y = grid(n)
    w0 = 2 * nq * np.cos(nq * x)[:, None] * np.cos(nq * y)[None, :]
    fac = 1.
    for alpha in (8. / 15., 2. / 15., 1. / 3.):
        d = alpha * (.5 * dt * 2 * nq**2 / re)
        fac *= (1. - d) / (1. + d)
    wn = np.zeros((n + 2, n + 2), order="F")
    wn[1:n + 2, 1:n + 2] = w0
    numerical = {"hybrid": cm.numerical_hybrid, "ps23": cm.numerical_ps23, "ps32": cm.numerical_ps32}[which]
    ut = numerical(n, n, nt, dx, dy, dt, re, x, y, wn, 1)
    assert rel_l2(ut, w0 * fac**nt) < tol
    return ut


def check_ps32_fused(cm, onp, n, nt, noise=.5, mode=1):
    """the opt-in fused kernels of the 3/2 rule (set_option "ps32_fuse") against the default path and the oracle.
    mode 1: spectra computed in the load stage of the inverse row transform -- same arithmetic, so the fields must be
    bit-identical; mode 2: also folded along i there (fold before the transform instead of after: equal to rounding)"""
    dx, dy, x, y = grid(n)
    w = vm_field(n) + noise * noise_field(n, 5)
    p = cm.plan(n, n)
    p.set_option("ps32_fuse", 0)
    a = cm.numerical_ps32(n, n, nt, dx, dy, 1e-3, 1000., x, y, w, 1)
    l0 = p.launch_count
    p.set_option("ps32_fuse", mode)
    try:
        b = cm.numerical_ps32(n, n, nt, dx, dy, 1e-3, 1000., x, y, w, 1)
    finally:
        p.set_option("ps32_fuse", 0)
    per_stage = {1: 11, 2: 10}[mode]  # default: 12
    assert p.launch_count - l0 == 7 + 3 * per_stage * nt  # 4 (upload, K1, KX, E0) + stages + 3 (final field)
    if mode == 1:
        assert np.array_equal(a, b)
    else:
        assert rel_l2(b, a) < 1e-13
    assert rel_l2(b, onp.ps_numerical(32, n, n, nt, dx, dy, 1e-3, 1000., w)) < TOL_RUN


def check_ldc(cm, onp, n, nt, dt=None, re=100., from_rest=False, tol=TOL_RUN):
    """18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl `numerical` against the numpy/scipy restatement
    (oracle_np.ldc_numerical).  from_rest: the script's own initial condition (wn = sn = 0, the lid drives the flow);
    otherwise a noisy start that exercises every sine mode and all wall formulas."""
    dx = 1. / n
    dt = min(.001, 0.2 * dx * dx * re) if dt is None else dt
    wn = np.zeros((n + 1, n + 1), order="F")
    sn = np.zeros((n + 1, n + 1), order="F")
    if not from_rest:
        rng = np.random.default_rng(n)
        wn[...] = rng.uniform(-1, 1, (n + 1, n + 1))
        sn[1:n, 1:n] = 1e-2 * rng.uniform(-1, 1, (n - 1, n - 1))
    w2, s2 = wn.copy(order="F"), sn.copy(order="F")
    rms, rms2 = np.zeros(nt), np.zeros(nt)
    onp.ldc_numerical(n, n, nt, dx, dx, dt, re, w2, s2, rms2)
    assert cm.numerical_ldc(n, n, nt, dx, dx, dt, re, wn, sn, rms) is None  # mutates wn, sn, rms like the reference
    assert rel_l2(wn, w2) < tol and rel_l2(sn, s2) < tol
    assert np.all(sn[0, :] == 0) and np.all(sn[n, :] == 0) and np.all(sn[:, 0] == 0) and np.all(sn[:, n] == 0)
    assert np.max(np.abs(rms / rms2 - 1.)) < 1e-8
    return wn, sn, rms
