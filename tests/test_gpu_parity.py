"""GPU suite (-m gpu): the product library libvmk.so, called through the C ABI (ctypes mirror in
cfd_julia_b200.common), against the CPU oracle on the same inputs; golden fixtures; the reference's
recorded numbers; and size-independent properties at BASELINE.json's full 8192^2 size.

Tolerance (north_star): relative L2 <= 1e-10 on vorticity and streamfunction; single operator calls
are held to 1e-12."""
import os

import numpy as np
import pytest

import parity_cases as pc
from helpers import ghost_fill, grid, noise_field, rel_l2, stable_dt, tgv_field, vm_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU suite needs a CUDA device"
    import cfd_julia_b200
    from cfd_julia_b200.common import Common
    lib = cfd_julia_b200.default_library()  # raises if libvmk.so is missing: no fallback
    # (VMK_LIB: an explicitly chosen build of the same CUDA library, e.g. tools/devbuild.sh during kernel tuning)
    assert lib.prefix == "vmk_" and (lib.path.endswith("libvmk.so") or os.environ.get("VMK_LIB") == lib.path)
    cm = Common(lib)
    yield cm
    cm.clear_plans()


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_fps_noise(gpu, oracle_c, n):
    pc.check_fps_noise(gpu, oracle_c, n)
    if n >= 2048:
        gpu.clear_plans()


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_rhs_noise(gpu, oracle_c, n):
    pc.check_rhs(gpu, oracle_c, noise_field(n, seed=n))
    if n >= 2048:
        gpu.clear_plans()


@pytest.mark.parametrize("tag", pc.REF_PY)
def test_ref_py_fixtures(gpu, tag):
    """The CUDA path against vectors computed by the REFERENCE'S OWN code: its two Python twins of script 19, run
    unmodified (tests/golden/make_ref_fixtures.py): vorticity and streamfunction after 10-50 RK3 steps from the twin's
    initial field, plus one rhs / Poisson solve on white noise.  Tolerance 1e-12 (parity_cases.TOL_REF_PY)."""
    pc.check_ref_py_lib(gpu, tag)


def test_rhs_vm_ic(gpu, oracle_c):
    pc.check_rhs(gpu, oracle_c, vm_field(128))


@pytest.mark.parametrize("n,nt", [(32, 50), (64, 50), (128, 50), (256, 20), (512, 20), (1024, 100), (2048, 10),
                                   (4096, 3)])
def test_numerical_vm(gpu, oracle_c, n, nt):
    pc.check_numerical(gpu, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    if n >= 2048:
        gpu.clear_plans()


def test_numerical_noise(gpu, oracle_c):
    pc.check_numerical(gpu, oracle_c, noise_field(64, 3), 5, 1e-3, 100.)


def test_vm_defaults_full_run(gpu, oracle_c):
    """Config 1: vm.jl defaults, 128^2, Re=1000, dt=.01, all 2000 steps."""
    n = 128
    out = pc.check_numerical(gpu, oracle_c, vm_field(n), 2000, .01, 1000.)
    dx = 2 * np.pi / n
    # invariants recorded by the survey probe: circulation is conserved, enstrophy decays
    assert abs(out[:n, :n].sum() * dx * dx - 1.999999996340) < 1e-10
    assert abs(0.5 * (out[:n, :n]**2).sum() * dx * dx - 0.397839858490) < 1e-9


def test_tgv_1024(gpu, oracle_c):
    """Config 3 (shortened): tgv at 1024^2, Re=10, dt=1e-4 (script default .01 is unstable there), 50 steps."""
    from cfd_julia_b200.common import compute_l2norm_bnds, exact_tgv
    n = 1024
    dx, dy, x, y = grid(n)
    out = pc.check_numerical(gpu, oracle_c, tgv_field(n), 50, 1e-4, 10.)
    ue = exact_tgv(n, n, x, y, 50 * 1e-4, 10.)
    # error against the analytic decay (tgv.jl:131-139); the oracle gives 3.16479e-6 for this configuration
    assert abs(compute_l2norm_bnds(n, n, out - ue) - 3.1647889878e-6) < 1e-12


def test_tgv_1024_full_config(gpu):
    """Config 3 in full: tgv.jl at 1024^2, Re=10, nq=4, tf=1 with dt=1e-4 (10 000 steps; the script's dt=.01 is unstable
    at this resolution).  No oracle run (it would take minutes): the analytic Taylor-Green decay
    (tgv.jl:82-90) is the reference, as in the script's own self-check (tgv.jl:131-139); second-order accuracy of the
    scheme gives an L2 error ~ (64/1024)^2 of the 64^2 default-case error."""
    from cfd_julia_b200.common import compute_l2norm_bnds, exact_tgv
    n, nt, dt, re = 1024, 10000, 1e-4, 10.
    dx, dy, x, y = grid(n)
    wn = tgv_field(n)
    out = gpu.numerical_tgv(n, n, nt, dx, dy, dt, re, wn)
    ue = exact_tgv(n, n, x, y, nt * dt, re)
    l2 = compute_l2norm_bnds(n, n, out - ue)
    assert np.isfinite(out).all()
    assert l2 < 5e-5 and np.max(np.abs(out - ue)) < 1e-4  # 64^2 case: L2 6.9e-3 (dt error included)
    assert abs(np.max(np.abs(out)) - 8. * np.exp(-32. / re)) < 1e-3  # amplitude 2 nq exp(-2 nq^2 t / re)


@pytest.mark.parametrize("n,nt,ns", [(32, 20, 4), (64, 20, 2), (128, 50, 5), (512, 10, 1), (1024, 10, 2), (2048, 3, 1),
                                      (4096, 2, 1)])
def test_hybrid_solver(gpu, oracle_np, n, nt, ns):
    """20_NS2D_Hybrid_Solver/hybrid.jl (RK3 / Crank-Nicolson in Fourier space) against the numpy oracle"""
    pc.check_hybrid(gpu, oracle_np, n, nt, ns=ns)
    if n >= 2048:
        gpu.clear_plans()


def test_hybrid_defaults_500_steps(gpu, oracle_np):
    """hybrid.jl's own configuration (128^2, dt = .01, Re = 1000), first 500 of its 2000 steps"""
    pc.check_hybrid(gpu, oracle_np, 128, 500, dt=.01, ns=10)


def test_hybrid_8192_properties(gpu):
    """full size, no oracle run (minutes of numpy FFTs): finite, mean-free, periodic duplicates, enstrophy decays"""
    n = 8192
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    ut = gpu.numerical_hybrid(n, n, 2, dx, dy, 1e-4, 1000., x, y, w, 1)
    assert np.isfinite(ut).all() and abs(ut[:n, :n].mean()) < 1e-12
    assert np.array_equal(ut[n, :], ut[0, :]) and np.array_equal(ut[:, n], ut[:, 0])
    w0 = w[1:n + 1, 1:n + 1] - w[1:n + 1, 1:n + 1].mean()
    e0, e1 = float((w0**2).sum()), float((ut[:n, :n]**2).sum())
    assert e1 < e0 and (e0 - e1) / e0 < 1e-3
    assert rel_l2(ut[:n, :n], w0) < 1e-3  # two steps of dt = 1e-4 barely move the field
    gpu.clear_plans()


@pytest.mark.parametrize("n,nt", [(16, 20), (64, 20), (256, 10), (1024, 5), (4096, 2)])
def test_lid_driven_cavity(gpu, oracle_np, n, nt):
    """18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl against the numpy/scipy oracle"""
    pc.check_ldc(gpu, oracle_np, n, nt)
    if n >= 1024:
        gpu.clear_plans()


def test_lid_driven_cavity_script_config(gpu, oracle_np):
    """the script's own configuration (64^2, dt = .001, Re = 100, from rest), first 300 of its 10 000 steps"""
    wn, sn, rms = pc.check_ldc(gpu, oracle_np, 64, 300, dt=.001, from_rest=True)
    assert sn.min() < -0.04 and rms[0] > rms[-1] > 0


def test_golden(gpu):
    pc.check_golden(gpu)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512])
def test_order_jl(gpu, n):
    pc.check_order_jl(gpu, n)


def test_ps_fft_4096(gpu, oracle_c):
    """Config 2: fft_p.jl's solve at 4096^2 against the oracle."""
    from helpers import mms
    n = 4096
    dx, f, _ = mms(n)
    u = gpu.ps_fft(n, n, dx, dx, f)
    assert rel_l2(u, oracle_c.ps_fft(n, n, dx, dx, f)) < 1e-12
    gpu.clear_plans()


def test_tgv_defaults(gpu):
    pc.check_tgv_defaults(gpu)


def test_snapshots(gpu, oracle_c, tmp_path):
    pc.check_snapshots(gpu, oracle_c, tmp_path)


def test_errors(gpu):
    pc.check_errors(gpu)


def test_graph_and_plain_launch_agree(gpu):
    n = 256
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n)
    res = []
    for graph in (1, 0):
        p = gpu.plan(n, n)
        p.set_option("graph", graph)
        p.upload(w0)
        p.step(dx, dy, 1e-3, 1000., 7)
        wn = np.zeros_like(w0)
        p.download(wn)
        res.append(wn)
        assert p.step_elapsed_ms() > 0
    assert np.array_equal(res[0], res[1])
    gpu.plan(n, n).set_option("graph", 1)


# ---- the two forms of the solve along j (option "fps_mode"): K2's FFT pair, and the cyclic tridiagonal solve by
# two-sided recurrences of csrc/vmk_tri.cuh (the default from 2048^2 up, so the sized tests above already run it) ----
@pytest.mark.parametrize("n,k0", [(64, 0), (128, 1), (256, 0), (512, 7), (1024, 0), (2048, 64)])
def test_tri_fps_small_sizes(gpu, oracle_c, n, k0):
    gpu.clear_plans()
    p = gpu.plan(n, n)
    p.set_option("fps_mode", 1)
    p.set_option("tri_k0", k0)
    pc.check_fps_noise(gpu, oracle_c, n, seed=n)
    gpu.clear_plans()


@pytest.mark.parametrize("n,nt", [(64, 20), (256, 10), (512, 10)])
def test_tri_rhs_and_numerical(gpu, oracle_c, n, nt):
    gpu.clear_plans()
    gpu.plan(n, n).set_option("fps_mode", 1)
    pc.check_rhs(gpu, oracle_c, noise_field(n, seed=n + 7))
    pc.check_numerical(gpu, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    gpu.clear_plans()


@pytest.mark.parametrize("n", [2048, 8192])
def test_fft_form_along_j_still_served(gpu, oracle_c, n):
    """fps_mode 0 (K1 -> K2 -> K3 with the transposed / PIECES layouts) against the oracle, and against fps_mode 1"""
    gpu.clear_plans()
    dx, dy, _, _ = grid(n)
    f = np.asfortranarray(np.random.default_rng(n).uniform(-1, 1, (n, n)))
    out = []
    for mode in (0, 1):
        p = gpu.plan(n, n)
        p.set_option("fps_mode", mode)
        s = np.zeros((n + 2, n + 2), order="F")
        l0 = p.launch_count
        gpu.fps(n, n, dx, dy, None, None, None, None, f, s)
        assert p.launch_count - l0 == (6 if mode else 3)
        out.append(s)
    ref = np.zeros((n + 2, n + 2), order="F")
    oracle_c.fps(n, n, dx, dy, f, ref)
    for s in out:
        assert rel_l2(s[1:n + 1, 1:n + 1], ref[1:n + 1, 1:n + 1]) < 1e-12
    assert rel_l2(out[0], out[1]) < 1e-13
    gpu.clear_plans()


# ---- the fused form of the recurrences (fps_mode 2, one GPU; the default at 8192^2): forward recurrence in K1's
# epilogue, backward recurrence in K3's load stage, per-slot state in tensor memory (tcgen05.ld / .st) ----
@pytest.mark.parametrize("n,k0,grid_ctas", [(512, 0, 0), (512, 3, 37), (1024, 0, 0), (1024, 64, 101), (2048, 0, 0),
                                            (4096, 0, 0), (4096, 16, 123)])
def test_fused_fps_sizes(gpu, oracle_c, n, k0, grid_ctas):
    gpu.clear_plans()
    p = gpu.plan(n, n)
    p.set_option("fps_mode", 2)
    p.set_option("tri_k0", k0)
    if grid_ctas:  # ragged blocks of row pairs per unit
        p.set_option("fz_grid", grid_ctas)
    l0 = p.launch_count
    pc.check_fps_noise(gpu, oracle_c, n, seed=n + 1)
    assert p.launch_count - l0 == 4  # K1, K2 on the rows kx < K0, scan, K3
    gpu.clear_plans()


@pytest.mark.parametrize("n,nt", [(512, 10), (1024, 5), (2048, 3)])
def test_fused_rhs_and_numerical(gpu, oracle_c, n, nt):
    gpu.clear_plans()
    gpu.plan(n, n).set_option("fps_mode", 2)
    pc.check_rhs(gpu, oracle_c, noise_field(n, seed=n + 9))
    pc.check_numerical(gpu, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    gpu.clear_plans()


def test_fused_is_the_default_at_8192_and_agrees_with_the_other_forms(gpu, oracle_c):
    n = 8192
    gpu.clear_plans()
    dx, dy, _, _ = grid(n)
    f = np.asfortranarray(np.random.default_rng(n + 3).uniform(-1, 1, (n, n)))
    out = {}
    for mode in (-1, 1, 2):
        p = gpu.plan(n, n)
        p.set_option("fps_mode", mode)
        s = np.zeros((n + 2, n + 2), order="F")
        l0 = p.launch_count
        gpu.fps(n, n, dx, dy, None, None, None, None, f, s)
        assert p.launch_count - l0 == (6 if mode == 1 else 4)
        out[mode] = s
    ref = np.zeros((n + 2, n + 2), order="F")
    oracle_c.fps(n, n, dx, dy, f, ref)
    for s in out.values():
        assert rel_l2(s[1:n + 1, 1:n + 1], ref[1:n + 1, 1:n + 1]) < 1e-12
    assert np.array_equal(out[-1], out[2]) and rel_l2(out[1], out[2]) < 1e-13
    gpu.clear_plans()


# ---- full-size (8192^2) properties: the oracle takes ~6 s per step there, so one step is compared directly
# and longer runs are checked through size-independent properties ------------------------------------
def test_full_size_two_steps_vs_oracle(gpu, oracle_c):
    n = 8192
    pc.check_numerical(gpu, oracle_c, vm_field(n), 2, 1e-4, 1000.)
    gpu.clear_plans()


def test_full_size_properties(gpu):
    n = 8192
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n)
    p = gpu.plan(n, n)
    p.upload(w0)
    p.step(dx, dy, 1e-4, 1000., 10)
    wn = np.zeros_like(w0)
    psi = np.zeros_like(w0)
    p.download(wn, psi)
    wi = wn[1:n + 1, 1:n + 1]
    # periodic ghosts valid on download (vm.jl:68-76)
    assert np.array_equal(wn, ghost_fill(n, wn.copy(order="F")))
    assert np.array_equal(psi, ghost_fill(n, psi.copy(order="F")))
    # circulation is conserved by the Arakawa Jacobian + periodic Laplacian
    assert abs(wi.sum() - w0[1:n + 1, 1:n + 1].sum()) / w0[1:n + 1, 1:n + 1].sum() < 1e-12
    # psi of the last rhs call solves the discrete Poisson problem for the stage field: check the operator identity
    # lap(psi) = -(w_stage - mean) through linearity on a fresh solve instead (fps is linear in f)
    rng = np.random.default_rng(1)
    f1 = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
    f2 = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
    s1 = np.zeros_like(w0)
    s2 = np.zeros_like(w0)
    s3 = np.zeros_like(w0)
    gpu.fps(n, n, dx, dy, None, None, None, None, f1, s1)
    gpu.fps(n, n, dx, dy, None, None, None, None, f2, s2)
    gpu.fps(n, n, dx, dy, None, None, None, None, np.asfortranarray(2. * f1 - 3. * f2), s3)
    assert rel_l2(s3, 2. * s1 - 3. * s2) < 1e-12
    # and the 5-point Laplacian of the solution gives back f minus its mean (the zero mode is dropped, Common.jl:118;
    # modes in spectral row/column 0 carry the eps quirk, a 5e-13 relative perturbation)
    ghost_fill(n, s1)
    lap = ((s1[2:, 1:-1] - 2 * s1[1:-1, 1:-1] + s1[:-2, 1:-1]) / dx**2 +
           (s1[1:-1, 2:] - 2 * s1[1:-1, 1:-1] + s1[1:-1, :-2]) / dy**2)
    assert rel_l2(lap, f1 - f1.mean()) < 1e-7  # conditioning of the second difference at 8192^2 (~ 1e-16 * N^2)
    gpu.clear_plans()


def test_bench_contract_line(tmp_path):
    """bench.py prints ONE JSON line with the keys the driver reads (run on a small grid to keep it short)."""
    import json
    import subprocess
    import sys
    from helpers import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", "1024", "--steps", "3", "--warmup", "3",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    # 3 steps x 3 stages x (K1, K2, K3, K4), or x (K1, totals, scan, K2 on the low rows, solve, K3, K4) with the
    # recurrence form of the solve along j (the default from 2048^2 up), or x (K1, K2 on the low rows, scan, K3, K4) fused
    form = d["roofline"]["solve_along_j"]
    per_stage = 5 if "fused form" in form else 7 if form.startswith("recurrences") else 4
    assert d["value"] > 0 and d["gpu_launches"] == 9 * per_stage and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["peak"] > 0 and 0 < rf["frac"] < 1.2 and rf["achieved"] > 0
