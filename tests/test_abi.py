"""CPU suite: the product library builds, loads, exports every symbol include/vmk.h declares, and fails
loudly (never falls back to a CPU path) when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def lib():
    import cfd_julia_b200
    cfd_julia_b200.build()
    return cfd_julia_b200.default_library()


def _declared():
    src = open(os.path.join(ROOT, "include", "vmk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vmk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    raw = C.CDLL(lib.path)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/vmk.h but not exported by libvmk.so"


def test_binding_covers_header():
    from cfd_julia_b200 import SYMBOLS
    assert sorted(SYMBOLS) == _declared()


def test_product_does_not_link_oracle_or_emulator(lib):
    import subprocess
    out = subprocess.run(["ldd", lib.path], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out and "fftw" not in out and "cufft" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "cfd_julia_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                if f.endswith(".py"):  # the package never names (let alone loads) the emulator library
                    assert "libvmk_emul.so\"" not in txt and "emul/" not in txt.replace("tests/emul/libvmk_emul.so,", ""), f


def test_size_errors_without_device(lib):
    h = C.c_void_p()
    assert lib.plan_create(48, 48, C.byref(h)) == 1 and b"power of two" in lib.last_error()
    assert lib.plan_create(64, 128, C.byref(h)) == 1
    assert lib.plan_create(65536, 65536, C.byref(h)) == 1  # 16384 and 32768 are served by the cluster kernels
    assert lib.plan_create_slab(64, 64, 3, 2, C.byref(h)) == 3
    assert lib.version() >= 100


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.plan_create(64, 64, C.byref(h))
    assert rc == 2 and h.value is None
    assert b"cuda" in lib.last_error().lower()
    from cfd_julia_b200.common import Common, VmkError
    import numpy as np
    with pytest.raises(VmkError):
        Common(lib).fps(32, 32, .1, .1, None, None, None, None, np.zeros((32, 32), order="F"),
                        np.zeros((34, 34), order="F"))


def test_fused_kernels_use_tensor_memory():
    """the fused K1 / K3 (fps_mode 2) keep their per-slot state in tensor memory: their SASS holds tcgen05.ld / .st
    (LDTM / STTM) and the allocation (UTCATOMSWS) -- and no tensor-core MMA; checked on the built library, no GPU needed"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import cfd_julia_b200
    so = cfd_julia_b200.build()
    txt = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    seen = 0
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0]
        if ("K1FBody" in name or "K3FBody" in name) and "FftCfgILi13" in name:
            seen += 1
            assert len(re.findall(r"\bLDTM\.", f)) >= 16 and len(re.findall(r"\bSTTM\.", f)) >= 32, name
            assert "UTCATOMSWS" in f and "UTCMMA" not in f and "UTCHMMA" not in f, name
        elif "Body" in name:
            assert "LDTM" not in f or "FBody" in name, name  # no other kernel touches tensor memory
    assert seen == 2
