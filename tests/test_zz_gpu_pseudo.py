"""GPU suite (-m gpu), SURVEY 8f row f3: the pseudo-spectral solvers with the 2/3 rule
(22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl) and the 3/2 rule (21_NS2D_PseudoSpectral_32_Rule/
pseudospectral_32_rule.jl) through the C ABI of libvmk.so against the literal numpy
restatement (oracle_np.ps_numerical: full complex spectra, numpy C2C transforms).  Tolerance: relative L2 <= 1e-10.

(The file sorts last on purpose: it was written when two GPU-minutes of the round's budget were left.  Its kernels were
developed on the host emulator -- tests/test_emul.py::test_pseudospectral_* -- and confirmed on a B200 with the torch-
free scripts tests/quick/gpu_quick_f3*.py (profiles/r01_f3_*_gpu.txt); of this file the golden, closed-form and 64^2 cases
have run under pytest on a B200 (6 passed), the larger sizes and the 500-step runs had not when the budget ended.)"""
import numpy as np
import pytest

import parity_cases as pc
from helpers import grid, rel_l2, vm_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU suite needs a CUDA device"
    import cfd_julia_b200
    from cfd_julia_b200.common import Common
    lib = cfd_julia_b200.default_library()  # raises if libvmk.so is missing: no fallback
    assert lib.prefix == "vmk_" and lib.path.endswith("libvmk.so")
    cm = Common(lib)
    yield cm
    cm.clear_plans()


@pytest.mark.parametrize("n,nt,ns,noise", [(32, 20, 4, 1.), (64, 20, 2, 1.), (128, 50, 5, .05), (256, 10, 2, .5),
                                            (512, 10, 1, .05), (1024, 10, 2, .05), (2048, 3, 1, .05), (4096, 2, 1, .05)])
def test_pseudospectral_23_rule(gpu, oracle_np, n, nt, ns, noise):
    pc.check_ps23(gpu, oracle_np, n, nt, dt=1e-3 if noise >= .5 else None, ns=ns, noise=noise)
    if n >= 2048:
        gpu.clear_plans()


def test_pseudospectral_23_rule_defaults_500_steps(gpu, oracle_np):
    """the script's own configuration (128^2, dt = .01, Re = 1000, vm_ic), first 500 of its 2000 steps"""
    pc.check_ps23(gpu, oracle_np, 128, 500, dt=.01, ns=10, noise=0.)


def test_pseudospectral_8192_properties(gpu):
    """full size, no oracle run (minutes of numpy FFTs): finite, mean-free, periodic duplicates, enstrophy decays"""
    n = 8192
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    ut = gpu.numerical_ps23(n, n, 2, dx, dy, 1e-4, 1000., x, y, w, 1)
    assert np.isfinite(ut).all() and abs(ut[:n, :n].mean()) < 1e-12
    assert np.array_equal(ut[n, :], ut[0, :]) and np.array_equal(ut[:, n], ut[:, 0])
    w0 = w[1:n + 1, 1:n + 1] - w[1:n + 1, 1:n + 1].mean()
    e0, e1 = float((w0**2).sum()), float((ut[:n, :n]**2).sum())
    assert e1 < e0 and (e0 - e1) / e0 < 1e-3
    assert rel_l2(ut[:n, :n], w0) < 1e-3  # two steps of dt = 1e-4 barely move the field
    gpu.clear_plans()


# (the oracle's 1.5n-point numpy transforms dominate the run time of these cases: 2 steps at 4096^2 are ~100 s of host
# FFTs, so the largest oracle comparison is 2048^2 and 8192^2 is covered by test_pseudospectral_rules_agree_8192)
@pytest.mark.parametrize("n,nt,ns,noise", [(64, 20, 4, 1.), (128, 50, 5, .05), (256, 10, 2, .5), (512, 10, 1, .05),
                                            (1024, 6, 2, .05), (2048, 2, 1, .05)])
def test_pseudospectral_32_rule(gpu, oracle_np, n, nt, ns, noise):
    pc.check_ps32(gpu, oracle_np, n, nt, dt=1e-3 if noise >= .5 else None, ns=ns, noise=noise)
    if n >= 1024:
        gpu.clear_plans()


def test_pseudospectral_32_rule_defaults_500_steps(gpu, oracle_np):
    """the script's own configuration (128^2, dt = .01, Re = 1000, vm_ic), first 500 of its 2000 steps"""
    pc.check_ps32(gpu, oracle_np, 128, 500, dt=.01, ns=10, noise=0.)


def test_pseudospectral_rules_agree_8192(gpu):
    """full size (12288-point padded transforms as 3 x 4096), no oracle run: on a well-resolved field both de-aliasing
    rules compute the same Jacobian up to the field's spectral tail (the periodic image of vm_ic has a 1e-7 kink in its
    derivative at the box edge; 256^2, two steps of dt = 1e-3: 1.1e-12 on the emulator), so the two solvers must agree"""
    n = 8192
    dx, dy, x, y = grid(n)
    w = vm_field(n)
    u32 = gpu.numerical_ps32(n, n, 2, dx, dy, 1e-4, 1000., x, y, w, 1)
    gpu.clear_plans()
    u23 = gpu.numerical_ps23(n, n, 2, dx, dy, 1e-4, 1000., x, y, w, 1)
    gpu.clear_plans()
    assert np.isfinite(u32).all() and abs(u32[:n, :n].mean()) < 1e-12
    assert np.array_equal(u32[n, :], u32[0, :]) and np.array_equal(u32[:, n], u32[:, 0])
    assert rel_l2(u32, u23) < 1e-9


@pytest.mark.parametrize("which", ["hybrid", "ps23", "ps32"])
def test_spectral_solvers_tgv_known_answer(gpu, which):
    """closed-form answer (decaying Taylor-Green eigenfunction), no oracle involved; 1024^2 x 100 steps"""
    pc.check_spectral_tgv(gpu, which, 64, 100)
    pc.check_spectral_tgv(gpu, which, 1024, 100, tol=1e-11)
    gpu.clear_plans()


def test_golden_f_rows(gpu):
    """committed fixtures (tests/golden/spectral_64_20.npz, ldc_32_20.npz)"""
    pc.check_golden_f_rows(gpu)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,nt", [(64, 5), (256, 3), (1024, 2)])
def test_pseudospectral_32_rule_fused_option(gpu, oracle_np, n, nt, mode):
    """opt-in kernels (default off; emulator-validated only when they were written): mode 1 must equal the default path
    bit for bit, mode 2 to rounding"""
    pc.check_ps32_fused(gpu, oracle_np, n, nt, mode=mode)
    gpu.clear_plans()
