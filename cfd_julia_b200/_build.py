"""Builds cfd_julia_b200/libvmk.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU.  The library is git-ignored but travels to the GPU box with
the repo snapshot, so importing the package there does not need to rebuild.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libvmk.so")
SRC = os.path.join(HERE, "csrc", "vmk.cu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference arithmetic is unfused (oracle: -ffp-contract=off); FMAs are written explicitly where wanted
    "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


# Test build of the SAME CUDA sources with the thread-block-cluster kernels (production: 16384, 32768) enabled for
# 64 .. 8192, so that the GPU suite can compare them with the CPU oracle at sizes the oracle can afford.
SO_CLUSTER_TEST = os.path.join(HERE, "libvmk_cltest.so")


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libvmk.so")
    return cand


def sources():
    return [SRC] + sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "*.hpp"))) + [os.path.join(ROOT, "include", "vmk.h")]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def _compile(out: str, extra) -> None:
    """nvcc into a private temporary file, then an atomic rename: several ranks of one torchrun job that all find the
    library stale each produce a complete file, and no process ever maps a half-written one."""
    tmp = f"{out}.tmp.{os.getpid()}"
    try:
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + list(extra) + ["-o", tmp, SRC], cwd=ROOT)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        _compile(SO, ["-Xptxas", "-v"] if verbose else [])
    return SO


def build_cluster_test(force: bool = False) -> str:
    if force or not os.path.exists(SO_CLUSTER_TEST) or any(
            os.path.getmtime(x) > os.path.getmtime(SO_CLUSTER_TEST) for x in sources()):
        _compile(SO_CLUSTER_TEST, ["-DVMK_CLUSTER_TEST"])
    return SO_CLUSTER_TEST


if __name__ == "__main__":
    print(build(force=True, verbose=True))
