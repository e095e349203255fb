"""CPU suite for the thread-block-cluster kernels (csrc/vmk_cluster.cuh): the same kernel bodies executed by the host
emulator, built with -DVMK_CLUSTER_TEST so that the cluster code path (production sizes 16384 / 32768) runs at
64 .. 4096 -- clusters of 2 and 4 CTAs, several transforms per CTA, the packed DC/Nyquist row, both V layouts and the
slab decomposition.  The emulator lets each CTA of a cluster run ahead to its next cluster barrier (alternating CTA
order), so a missing barrier around a distributed-shared-memory access fails these tests.
Also validates the size-independent property checks that tests/test_gpu_cluster.py applies at 16384^2 / 32768^2."""
import os
import subprocess

import numpy as np
import pytest

import parity_cases as pc
import test_gpu_cluster as tg
from helpers import ROOT, grid, noise_field, rel_l2, stable_dt, vm_field
from test_emul import _slab_run


@pytest.fixture(scope="module")
def emul_cl():
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "-s"])
    cm = Common(VmkLibrary(os.path.join(d, "libvmk_emul_cluster.so"), "vmke_"))
    yield cm
    cm.clear_plans()


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
def test_cluster_fps_noise(emul_cl, oracle_c, n):
    pc.check_fps_noise(emul_cl, oracle_c, n)
    emul_cl.clear_plans()


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_cluster_rhs_noise(emul_cl, oracle_c, n):
    pc.check_rhs(emul_cl, oracle_c, noise_field(n, seed=n))


@pytest.mark.parametrize("n,nt", [(64, 10), (128, 10), (256, 5), (512, 3), (1024, 2), (2048, 1)])
def test_cluster_numerical_vm(emul_cl, oracle_c, n, nt):
    pc.check_numerical(emul_cl, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    emul_cl.clear_plans()


@pytest.mark.parametrize("n,nranks", [(64, 2), (128, 4), (256, 8), (1024, 2)])
def test_cluster_slab_decomposition(emul_cl, oracle_c, n, nranks):
    _slab_run(emul_cl, oracle_c, n, nranks)


def test_cluster_lid_driven_cavity(emul_cl, oracle_np):
    """the cavity solver on top of the cluster kernels (production: nx = 8192 -> 16384-point transforms)"""
    pc.check_ldc(emul_cl, oracle_np, 64, 2)
    emul_cl.clear_plans()


def test_cluster_v_layouts_agree(emul_cl):
    n = 256
    dx, dy, _, _ = grid(n)
    w0 = noise_field(n, 11)
    res = []
    for pieces in (1, 0):
        p = emul_cl.plan(n, n)
        p.set_option("v_pieces", pieces)
        p.upload(w0)
        p.step(dx, dy, 1e-4, 1000., 2)
        wn = np.zeros_like(w0)
        p.download(wn)
        res.append(wn)
    emul_cl.plan(n, n).set_option("v_pieces", 1)
    assert rel_l2(res[0], res[1]) < 1e-14


def test_property_checks_used_at_full_size(emul_cl, oracle_c):
    """the analytic-mode and periodic-tiling checks of the GPU suite, at a size the emulator can afford"""
    n = 512
    hi, lo = tg._mode_sets(n)
    tg._check_modes_residual(emul_cl, n, hi, 1e-13)
    tg._check_modes(emul_cl, n, lo, 1e-13)
    tg._tiled_run(emul_cl, oracle_c, 512, 64, 2)


def test_spectral_solvers_refuse_cluster_sizes(emul_cl):
    """the hybrid / pseudo-spectral solvers have no cluster variant (rows of 16384 / 32768 points; here: the cluster
    test build at 128): VMK_ESIZE with a message, not a crash"""
    from cfd_julia_b200.common import VmkError
    n = 128
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    for fn in (emul_cl.numerical_hybrid, emul_cl.numerical_ps23, emul_cl.numerical_ps32):
        with pytest.raises(VmkError) as e:
            fn(n, n, 1, dx, dy, .01, 1000., x, y, w, 1)
        assert e.value.code == 1 and "8192" in str(e.value)
    emul_cl.clear_plans()


# ---- recurrence form of the solve along j (csrc/vmk_tri.cuh) on top of the cluster kernels' natural-layout K1 / K3 ----
@pytest.mark.parametrize("n,k0", [(64, 0), (128, 3), (256, 0), (512, 0), (1024, 0)])
def test_cluster_tri_fps(emul_cl, oracle_c, n, k0):
    emul_cl.clear_plans()
    p = emul_cl.plan(n, n)
    p.set_option("fps_mode", 1)
    p.set_option("tri_k0", k0)
    pc.check_fps_noise(emul_cl, oracle_c, n, seed=n + 1)
    emul_cl.clear_plans()


@pytest.mark.parametrize("n,nt", [(64, 8), (256, 4), (512, 2)])
def test_cluster_tri_numerical(emul_cl, oracle_c, n, nt):
    emul_cl.clear_plans()
    emul_cl.plan(n, n).set_option("fps_mode", 1)
    pc.check_rhs(emul_cl, oracle_c, noise_field(n, seed=n + 3))
    pc.check_numerical(emul_cl, oracle_c, vm_field(n), nt, stable_dt(n, 1000.), 1000.)
    emul_cl.clear_plans()


@pytest.mark.parametrize("n,nranks", [(128, 2), (256, 4), (512, 8)])
def test_cluster_tri_slab(emul_cl, oracle_c, n, nranks):
    _slab_run(emul_cl, oracle_c, n, nranks, {"fps_mode": 1})
