"""CPU suite: the kernel bodies (same source as the CUDA build) executed by the host emulator
(tests/emul) against the oracle.  Checks the index arithmetic of every FFT size configuration, the
transposed/packed spectrum layout, halo handling and the slab decomposition without a GPU.
The parity claims proper are made by tests/test_gpu_parity.py on the real library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity_cases as pc
from helpers import ROOT, grid, noise_field, rel_l2, stable_dt, tgv_field, vm_field


@pytest.fixture(scope="module")
def emul():
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "-s"])
    return Common(VmkLibrary(os.path.join(d, "libvmk_emul.so"), "vmke_"))


@pytest.fixture(scope="module")
def emul_split(emul):
    """The N = 8192 code path (split re/im exchange, separate landing buffer, octant twiddle table) compiled for
    512 .. 4096 as well, so that it is covered at sizes the CPU suite can afford."""
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    return Common(VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul_split.so"), "vmke_"))


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_split_path_fps(emul_split, oracle_c, n):
    pc.check_fps_noise(emul_split, oracle_c, n)
    emul_split.clear_plans()


@pytest.mark.parametrize("n,nt", [(512, 3), (1024, 2), (2048, 1)])
def test_split_path_numerical(emul_split, oracle_c, n, nt):
    pc.check_rhs(emul_split, oracle_c, noise_field(n, seed=n + 1))
    pc.check_numerical(emul_split, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    emul_split.clear_plans()


def test_split_path_slab(emul_split, oracle_c):
    _slab_run(emul_split, oracle_c, 512, 4)
    _slab_run(emul_split, oracle_c, 1024, 2)


def test_full_size_8192_fps(emul, oracle_c):
    """BASELINE's grid size through the emulator (one Poisson solve; ~20 s)."""
    emul.clear_plans()
    pc.check_fps_noise(emul, oracle_c, 8192)
    emul.clear_plans()


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048])
def test_fps_noise(emul, oracle_c, n):
    pc.check_fps_noise(emul, oracle_c, n)


def test_fps_4096(emul, oracle_c):
    pc.check_fps_noise(emul, oracle_c, 4096)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024])
def test_rhs_noise(emul, oracle_c, n):
    pc.check_rhs(emul, oracle_c, noise_field(n, seed=n))


def test_rhs_vm_ic(emul, oracle_c):
    pc.check_rhs(emul, oracle_c, vm_field(128))


@pytest.mark.parametrize("n,nt", [(32, 20), (64, 10), (128, 10), (256, 5), (512, 3), (1024, 2)])
def test_numerical_vm(emul, oracle_c, n, nt):
    pc.check_numerical(emul, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)


def test_numerical_noise(emul, oracle_c):
    pc.check_numerical(emul, oracle_c, noise_field(64, 3), 5, 1e-3, 100.)


@pytest.mark.parametrize("n,nt,ns", [(32, 6, 3), (64, 4, 2), (128, 3, 1), (256, 2, 1), (512, 2, 2), (1024, 1, 1)])
def test_hybrid_solver(emul, oracle_np, n, nt, ns):
    """hybrid.jl's RK3/CN spectral-space solver (kernel KH + the finite-difference path's K1/K3/K4)"""
    pc.check_hybrid(emul, oracle_np, n, nt, ns=ns)


def test_hybrid_solver_errors(emul):
    from cfd_julia_b200.common import Plan, VmkError
    n = 64
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    with pytest.raises(ZeroDivisionError):
        emul.numerical_hybrid(n, n, 3, dx, dy, .01, 1000., x, y, w, 5)
    with pytest.raises(VmkError):  # wavespace aliases ky = kx: dx != dy is not representable
        emul.numerical_hybrid(n, n, 2, dx, 2 * dy, .01, 1000., x, y, w, 1)
    p = Plan(emul.lib, n, n, 0, 2)
    ut = np.zeros((n + 1, n + 1), order="F")
    from cfd_julia_b200._lib import SNAPSHOT_FN
    assert emul.lib.hybrid_numerical(p.handle, 1, dx, dy, .01, 1000., w.ctypes.data, ut.ctypes.data, 0, SNAPSHOT_FN(),
                                     None) != 0
    p.close()


@pytest.mark.parametrize("n,nt,ns,noise", [(32, 6, 3, 1.), (64, 4, 2, 1.), (128, 3, 1, .05), (256, 2, 2, .5), (512, 2, 1, .05),
                                            (1024, 1, 1, .05)])
def test_pseudospectral_23_rule(emul, oracle_np, n, nt, ns, noise):
    """pseudospectral_23_rule.jl (kernel KP + kp_product + the finite-difference path's K1/K3) against
    the literal full-spectrum numpy restatement; noise = 1 is a white-noise start (every mode, incl. the half-weighted
    edge of the asymmetric retained band)"""
    pc.check_ps23(emul, oracle_np, n, nt, dt=1e-3 if noise >= .5 else None, ns=ns, noise=noise)


def test_pseudospectral_23_rule_defaults_60_steps(emul, oracle_np):
    """the script's own configuration (128^2, dt = .01, Re = 1000, vm_ic), its first 60 steps, snapshots every 20"""
    pc.check_ps23(emul, oracle_np, 128, 60, dt=.01, ns=3, noise=0.)


@pytest.mark.parametrize("n,nt,ns,noise", [(64, 6, 3, 1.), (128, 3, 1, 1.), (256, 2, 2, .5), (512, 2, 1, .05), (1024, 1, 1, .05)])
def test_pseudospectral_32_rule(emul, oracle_np, n, nt, ns, noise):
    """pseudospectral_32_rule.jl -- the 1.5n-point transforms as radix-3 splits into n/2-point ones on
    9 sub-grids (vmk_pseudo32.cuh) against the literal numpy restatement (numpy does the 1.5n-point transforms directly)"""
    pc.check_ps32(emul, oracle_np, n, nt, dt=1e-3 if noise >= .5 else None, ns=ns, noise=noise)
    emul.clear_plans()


def test_pseudospectral_32_rule_defaults_40_steps(emul, oracle_np):
    """the script's own configuration (128^2, dt = .01, Re = 1000, vm_ic), its first 40 steps, snapshots every 20"""
    pc.check_ps32(emul, oracle_np, 128, 40, dt=.01, ns=2, noise=0.)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,nt", [(64, 3), (128, 2), (512, 1)])
def test_pseudospectral_32_rule_fused_option(emul, oracle_np, n, nt, mode):
    pc.check_ps32_fused(emul, oracle_np, n, nt, mode=mode)


def test_pseudospectral_32_rule_fused_white_noise_and_option_range(emul, oracle_np):
    from cfd_julia_b200.common import VmkError
    pc.check_ps32_fused(emul, oracle_np, 64, 2, noise=1., mode=2)
    with pytest.raises(VmkError):
        emul.plan(64, 64).set_option("ps32_fuse", 3)


def test_pseudospectral_32_rule_errors(emul):
    from cfd_julia_b200.common import VmkError
    dx, dy, x, y = grid(32)
    with pytest.raises(VmkError) as e:  # the sub-grids are (n/2)^2 and the smallest transform is 32 points
        emul.numerical_ps32(32, 32, 1, dx, dy, .01, 1000., x, y, vm_field(32), 1)
    assert e.value.code == 1
    dx, dy, x, y = grid(64)
    with pytest.raises(VmkError):
        emul.numerical_ps32(64, 64, 1, dx, 2 * dy, .01, 1000., x, y, vm_field(64), 1)


@pytest.mark.parametrize("which", ["hybrid", "ps23", "ps32"])
def test_spectral_solvers_tgv_known_answer(emul, which):
    """closed-form answer (decaying Taylor-Green eigenfunction), no oracle involved"""
    pc.check_spectral_tgv(emul, which, 64, 25)
    pc.check_spectral_tgv(emul, which, 256, 3)


def test_pseudospectral_errors(emul):
    from cfd_julia_b200.common import Plan, VmkError
    n = 64
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    with pytest.raises(ZeroDivisionError):
        emul.numerical_ps23(n, n, 3, dx, dy, .01, 1000., x, y, w, 5)
    with pytest.raises(VmkError):  # ky = kx (pseudospectral_23_rule.jl:108): dx != dy is not representable
        emul.numerical_ps23(n, n, 2, dx, 2 * dy, .01, 1000., x, y, w, 1)
    p = Plan(emul.lib, n, n, 0, 2)  # slab plans: not supported by the spectral-space solvers
    ut = np.zeros((n + 1, n + 1), order="F")
    from cfd_julia_b200._lib import SNAPSHOT_FN
    assert emul.lib.ps23_numerical(p.handle, 1, dx, dy, .01, 1000., w.ctypes.data, ut.ctypes.data, 0, SNAPSHOT_FN(),
                                   None) != 0
    p.close()


def test_pseudospectral_then_finite_difference(emul, oracle_c, oracle_np):
    """the solvers share buffers and tables inside one plan: interleaving them must not disturb either"""
    pc.check_ps23(emul, oracle_np, 64, 2)
    pc.check_ps32(emul, oracle_np, 64, 2)
    pc.check_numerical(emul, oracle_c, vm_field(64), 3, .01, 1000.)
    pc.check_hybrid(emul, oracle_np, 64, 2)
    pc.check_ps23(emul, oracle_np, 64, 2)


@pytest.mark.parametrize("n,nt", [(16, 4), (32, 3), (64, 3), (128, 2), (256, 1)])
def test_lid_driven_cavity(emul, oracle_np, n, nt):
    """lid_driven_cavity.jl (sine-transform Poisson solve as the periodic solve of the odd extension)"""
    pc.check_ldc(emul, oracle_np, n, nt)
    emul.clear_plans()


def test_lid_driven_cavity_from_rest(emul, oracle_np):
    wn, sn, rms = pc.check_ldc(emul, oracle_np, 32, 20, dt=.001, from_rest=True)
    assert sn.min() < 0 and rms[0] > rms[-1] > 0  # the primary vortex spins up, the change per step decays


def test_lid_driven_cavity_errors(emul):
    from cfd_julia_b200.common import Plan, VmkError
    import ctypes as C
    n = 32
    w = np.zeros((n + 1, n + 1), order="F")
    rms = np.zeros(2)
    p = emul.plan(n, n)  # wrong plan size: the cavity needs 2n x 2n
    assert emul.lib.ldc_numerical(p.handle, n, n, 1, 1. / n, 1. / n, 1e-3, 100., w.ctypes.data, w.ctypes.data,
                                  rms.ctypes.data) == 1
    with pytest.raises(IndexError):
        emul.numerical_ldc(n, n, 5, 1. / n, 1. / n, 1e-3, 100., w, w.copy(order="F"), rms)
    with pytest.raises(IndexError):
        emul.numerical_ldc(n, n, 1, 1. / n, 1. / n, 1e-3, 100., np.zeros((n, n), order="F"), w, rms)


def test_plan_is_serialised_across_host_threads(emul, oracle_c):
    """a plan shared by two host threads (ctypes releases the GIL around the calls) -- the per-plan mutex
    must serialise the entry points; without it the two solves interleave on the same device buffers"""
    import threading
    n = 256
    dx, dy, x, y = grid(n)
    rng = np.random.default_rng(1)
    fs = [np.asfortranarray(rng.uniform(-1, 1, (n, n))) for _ in range(2)]
    refs = []
    for f in fs:
        s = np.zeros((n + 2, n + 2), order="F")
        oracle_c.fps(n, n, dx, dy, f, s)
        refs.append(s)
    emul.plan(n, n)  # created once, shared
    outs = [[np.zeros((n + 2, n + 2), order="F") for _ in range(4)] for _ in range(2)]

    def work(i):
        for s in outs[i]:
            emul.fps(n, n, dx, dy, None, None, None, None, fs[i], s)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(2):
        for s in outs[i]:
            assert rel_l2(s[1:n + 1, 1:n + 1], refs[i][1:n + 1, 1:n + 1]) < 1e-12


def test_profile_option_counts_every_kernel_class(emul):
    """set_option("profile", 1) + vmk_profile_read: per-class launch counts of one step of each spectral-space solver
    (k1 row-forward, k2 spectrum-row kernels, k3 row-inverse, k4 pointwise), summed over the plan and its inner plan"""
    n = 64
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    p = emul.plan(n, n)
    p.set_option("profile", 1)
    p.profile_read()
    emul.numerical_hybrid(n, n, 1, dx, dy, .01, 1000., x, y, w, 1)
    got = {k: v["launches"] for k, v in p.profile_read().items()}
    assert got == {"k1": 4, "k2": 4, "k3": 7, "k4": 3}  # init K1+KH; per stage 2 K3, K4, K1, KH; final K3
    emul.numerical_ps23(n, n, 1, dx, dy, .01, 1000., x, y, w, 1)
    got = {k: v["launches"] for k, v in p.profile_read().items()}
    assert got == {"k1": 4, "k2": 5, "k3": 13, "k4": 3}  # per stage 4 K3, product, K1, KP; final KP(4) + K3
    emul.numerical_ps32(n, n, 1, dx, dy, .01, 1000., x, y, w, 1)
    got = p.profile_read()
    # init: K1, KX, E0; per stage: E1, KX, E2, 4 K3, product, K1, E3, KX, E4; final: E5, KX, K3
    assert {k: v["launches"] for k, v in got.items()} == {"k1": 4, "k2": 8, "k3": 13, "k4": 17}
    assert all(v["ms"] >= 0 for v in got.values())
    p.set_option("profile", 0)
    emul.numerical_ps23(n, n, 1, dx, dy, .01, 1000., x, y, w, 1)
    assert sum(v["launches"] for v in p.profile_read().values()) == 0


def test_golden(emul):
    pc.check_golden(emul)


def test_golden_f_rows(emul):
    pc.check_golden_f_rows(emul)


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512])
def test_order_jl(emul, n):
    pc.check_order_jl(emul, n)


def test_tgv_defaults(emul):
    pc.check_tgv_defaults(emul)


def test_snapshots(emul, oracle_c, tmp_path):
    pc.check_snapshots(emul, oracle_c, tmp_path)


def test_errors(emul):
    pc.check_errors(emul)


def test_device_path_and_options(emul, oracle_c):
    n = 64
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n)
    p = emul.plan(n, n)
    p.set_option("k4_rows", 5)  # ragged last row block
    p.upload(w0)
    p.step(dx, dy, .01, 1000., 2)
    p.step(dx, dy, .01, 1000., 1)
    wn = np.zeros_like(w0)
    psi = np.zeros_like(w0)
    p.download(wn, psi)
    ref = w0.copy(order="F")
    _, s = oracle_c.numerical(n, n, 3, dx, dy, .01, 1000., ref)
    assert rel_l2(wn, ref) < 1e-12 and rel_l2(psi, s) < 1e-12
    assert p.launch_count >= 36 and p.device_bytes > 0
    prof = p.profile_steps(dx, dy, .01, 1000., 1)
    assert all(prof[k]["launches"] == 3 for k in ("k1", "k2", "k3", "k4"))
    p.set_option("k4_rows", 32)


@pytest.mark.parametrize("n,nranks", [(64, 2), (64, 4), (128, 8), (256, 2)])
def test_slab_decomposition(emul, oracle_c, n, nranks):
    """Ranks as plans in one process (vmk_peer_attach_local); the emulator runs launches synchronously,
    so stepping the ranks kernel by kernel in lock-step stands in for the cross-rank barrier."""
    _slab_run(emul, oracle_c, n, nranks)


@pytest.mark.parametrize("n,nranks,opts", [
    (128, 4, {"a2a_engine": 0, "a2a_chunks": 4, "a2a_order": 1}),   # column-chunked K1, SM push with interleaved peers
    (128, 8, {"k2_push": 2, "k2_chunks": 4, "a2a_chunks": 2}),      # K2 staged + SM push of row chunks
    (256, 2, {"k2_push": 2, "a2a_engine": 0, "a2a_order": 1, "a2a_ctas": 3}),
    (64, 4, {"k2_push": 0}),                                       # K2 staged + copy engines
])
def test_slab_transpose_options(emul, oracle_c, n, nranks, opts):
    """every way the two transposes of the distributed FFT can be scheduled (vmk_set_option) gives the same fields"""
    _slab_run(emul, oracle_c, n, nranks, opts)


def _slab_run(emul, oracle_c, n, nranks, opts=None):
    from cfd_julia_b200.common import Plan
    from cfd_julia_b200._lib import BARRIER_FN
    lib = emul.lib
    plans = [Plan(lib, n, n, r, nranks) for r in range(nranks)]
    for p in plans:
        for k, v in (opts or {}).items():
            p.set_option(k, v)
    arr = (C.c_void_p * nranks)(*[p.handle for p in plans])
    for p in plans:
        lib.check(lib.peer_attach_local(p.handle, arr))
    # lock-step execution: each rank runs on its own thread, the barrier hook is a host barrier
    import threading
    bar = threading.Barrier(nranks)
    hook = BARRIER_FN(lambda _u: bar.wait())
    for p in plans:
        lib.check(lib.barrier_hook(p.handle, hook, None))
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n) + 0.1 * noise_field(n, 5)
    outs = [np.zeros_like(w0) for _ in plans]
    psis = [np.zeros_like(w0) for _ in plans]
    errs = []

    def run(r):
        try:
            p = plans[r]
            p.upload(w0)
            bar.wait()
            p.step(dx, dy, 1e-3, 1000., 3)
            bar.wait()
            p.download(outs[r], psis[r])
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)
            bar.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    ref = w0.copy(order="F")
    _, s = oracle_c.numerical(n, n, 3, dx, dy, 1e-3, 1000., ref)
    nj = n // nranks
    for r in range(nranks):
        rows = slice(r * nj, (r + 1) * nj + 2)  # the rank's ghosted rows j0 .. j0+NJ+1
        assert rel_l2(outs[r][:, rows], ref[:, rows]) < 1e-12
        assert rel_l2(psis[r][:, rows], s[:, rows]) < 1e-12
    for p in plans:
        p.close()


@pytest.mark.parametrize("tag", pc.REF_PY)
def test_ref_py_fixtures(emul, tag):
    """vectors computed by the reference's own Python twins of script 19 (tests/golden/make_ref_fixtures.py)"""
    pc.check_ref_py_lib(emul, tag)


@pytest.mark.parametrize("n,nt", [(32, 7), (64, 5), (128, 4), (256, 3)])
def test_fused_small_grid_step(emul, oracle_c, n, nt):
    """N <= 256: the whole step loop as one cluster launch (ks_body) is bit-identical to the 12-launch path"""
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n) + 0.1 * noise_field(n, 9)
    dt = stable_dt(n, 1000.)
    outs = []
    for fused in (1, 0):
        emul.clear_plans()
        p = emul.plan(n, n)
        p.set_option("fuse_small", fused)
        p.set_option("k4_rows", 5 if n == 64 else 32)
        p.upload(w0)
        l0 = p.launch_count
        p.step(dx, dy, dt, 1000., nt - 1)
        p.step(dx, dy, dt, 1000., 1)
        assert p.launch_count - l0 == (2 if fused else 12 * nt)
        wn, psi = np.zeros_like(w0), np.zeros_like(w0)
        p.download(wn, psi)
        outs.append((wn, psi))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    ref = w0.copy(order="F")
    _, s = oracle_c.numerical(n, n, nt, dx, dy, dt, 1000., ref)
    assert rel_l2(outs[0][0], ref) < 1e-12 and rel_l2(outs[0][1], s) < 1e-12
    emul.clear_plans()


# ---- the solve along j as a cyclic tridiagonal solve by two-sided recurrences (csrc/vmk_tri.cuh) -----------------------
@pytest.mark.parametrize("n,k0", [(64, 0), (128, 1), (256, 0), (512, 5), (1024, 0), (2048, 64)])
def test_tri_fps(emul, oracle_c, n, k0):
    """fps with fps_mode = 1 (FFT along i, recurrences along j, rows kx < K0 by K2) against the oracle's FFT x FFT"""
    emul.clear_plans()
    p = emul.plan(n, n)
    p.set_option("fps_mode", 1)
    p.set_option("tri_k0", k0)
    pc.check_fps_noise(emul, oracle_c, n, seed=n)
    emul.clear_plans()


@pytest.mark.parametrize("n,nt", [(64, 10), (128, 6), (256, 4), (512, 3)])
def test_tri_rhs_and_numerical(emul, oracle_c, n, nt):
    emul.clear_plans()
    emul.plan(n, n).set_option("fps_mode", 1)
    pc.check_rhs(emul, oracle_c, noise_field(n, seed=n + 7))
    pc.check_numerical(emul, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    emul.clear_plans()


def test_tri_default_by_size(emul, oracle_c):
    """the recurrence form is the default from 2048^2 up: 6 launches per Poisson solve instead of 3"""
    emul.clear_plans()
    n = 2048
    dx, dy, _, _ = grid(n)
    p = emul.plan(n, n)
    p.upload(vm_field(n))
    l0 = p.launch_count
    p.step(dx, dy, stable_dt(n, 1000.), 1000., 1)
    assert p.launch_count - l0 == 3 * (6 + 1)  # K1, totals, scan, K2 on the rows kx < K0, solve, K3; K4
    p.set_option("fps_mode", 0)
    l0 = p.launch_count
    p.step(dx, dy, stable_dt(n, 1000.), 1000., 1)
    assert p.launch_count - l0 == 12
    with pytest.raises(Exception):
        emul.plan(32, 32).set_option("fps_mode", 1)
    emul.clear_plans()


@pytest.mark.parametrize("n,nranks,k0", [(64, 2, 0), (128, 4, 3), (256, 8, 0), (256, 2, 16), (512, 4, 0)])
def test_tri_slab(emul, oracle_c, n, nranks, k0):
    """slab decomposition without transposes: three complex numbers per kx and rank cross the ranks"""
    _slab_run(emul, oracle_c, n, nranks, {"fps_mode": 1, "tri_k0": k0})


@pytest.mark.parametrize("n,nt", [(128, 5), (512, 2)])
def test_tri_zigzag_is_bit_identical(emul, n, nt):
    """option "zigzag": consecutive streaming kernels sweep the rows in alternating directions (each starts on what
    its predecessor left in L2); the arithmetic per point does not change"""
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n) + 0.1 * noise_field(n, 4)
    outs = []
    for zz in (0, 1):
        emul.clear_plans()
        p = emul.plan(n, n)
        p.set_option("fps_mode", 1)
        p.set_option("zigzag", zz)
        p.set_option("k4_rows", 8)
        p.upload(w0)
        p.step(dx, dy, stable_dt(n, 1000.), 1000., nt)
        wn, psi = np.zeros_like(w0), np.zeros_like(w0)
        p.download(wn, psi)
        outs.append((wn, psi))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    emul.clear_plans()


def test_tri_zigzag_slab(emul, oracle_c):
    _slab_run(emul, oracle_c, 256, 4, {"fps_mode": 1, "zigzag": 1})


# ---- the fused form of the recurrences (fps_mode 2): forward recurrence in K1's epilogue, backward in K3's load stage,
# per-slot state in tensor memory on the GPU (a per-thread array here) ----
@pytest.mark.parametrize("n,k0,grid_ctas", [(512, 0, 0), (512, 3, 5), (1024, 0, 0), (1024, 64, 7), (2048, 0, 0)])
def test_fused_fps(emul, oracle_c, n, k0, grid_ctas):
    """fps with fps_mode = 2 against the oracle's FFT x FFT; grid_ctas != 0: ragged blocks of row pairs per unit"""
    emul.clear_plans()
    p = emul.plan(n, n)
    p.set_option("fps_mode", 2)
    p.set_option("tri_k0", k0)
    if grid_ctas:
        p.set_option("fz_grid", grid_ctas)
    l0 = p.launch_count
    pc.check_fps_noise(emul, oracle_c, n, seed=n + 1)
    assert p.launch_count - l0 == 4  # K1, K2 on the rows kx < K0, scan, K3
    emul.clear_plans()


@pytest.mark.parametrize("n,nt,grid_ctas", [(512, 3, 0), (512, 2, 3), (1024, 2, 0)])
def test_fused_rhs_and_numerical(emul, oracle_c, n, nt, grid_ctas):
    emul.clear_plans()
    p = emul.plan(n, n)
    p.set_option("fps_mode", 2)
    if grid_ctas:
        p.set_option("fz_grid", grid_ctas)
    pc.check_rhs(emul, oracle_c, noise_field(n, seed=n + 9))
    pc.check_numerical(emul, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    emul.clear_plans()


def test_fused_form_refused_where_it_does_not_apply(emul):
    emul.clear_plans()
    p = emul.plan(256, 256)  # fewer than 32 threads per transform: the per-warp state does not apply
    with pytest.raises(Exception):
        p.set_option("fps_mode", 2)
    emul.clear_plans()


@pytest.mark.parametrize("mode,ratio,grid_ctas", [(1, 1.7, 0), (2, 1.7, 0), (2, 0.4, 0), (2, 60.0, 0), (2, 60.0, 1)])
def test_recurrence_forms_anisotropic_grid(emul, oracle_c, mode, ratio, grid_ctas):
    """dy != dx: r, the block tables and (fused form) the range guard of the (1/r)^M Horner sum follow the grid; at
    dy / dx = 60 with blocks of 130 rows ((1/r)^M = 1e540) the guard sends fps_mode 2 back to the separate kernels"""
    n = 512
    emul.clear_plans()
    p = emul.plan(n, n)
    p.set_option("fps_mode", mode)
    if grid_ctas:
        p.set_option("fz_grid", grid_ctas)
    dx = 2 * np.pi / n
    dy = ratio * dx
    f = np.asfortranarray(np.random.default_rng(17).uniform(-1, 1, (n, n)))
    s = np.zeros((n + 2, n + 2), order="F")
    ref = np.zeros((n + 2, n + 2), order="F")
    l0 = p.launch_count
    emul.fps(n, n, dx, dy, None, None, None, None, f, s)
    assert p.launch_count - l0 == (6 if mode == 1 or grid_ctas else 4)
    oracle_c.fps(n, n, dx, dy, f, ref)
    assert rel_l2(s[1:n + 1, 1:n + 1], ref[1:n + 1, 1:n + 1]) < 1e-12
    emul.clear_plans()
