"""GPU suite, needs >= 2 devices: the slab decomposition (one process per GPU, peer loads/stores over NVLink inside
the kernels, device-side flag barrier) against the CPU oracle.  The same decomposition logic is covered on the CPU by
tests/test_emul.py::test_slab_decomposition and tests/test_dist_cpu.py."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_parity(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,cases", [(2, "16384@2048:1"), (8, "32768@1024:1")])
def test_slab_parity_cluster_sizes(world, cases):
    """rows of 16384 / 32768 points (thread-block-cluster kernels) across GPUs; (8, 32768) is BASELINE config 5"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, VMK_MG_CASES=cases))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_process_two_devices():
    """The Julia host model: ONE process and one host thread drive two devices through
    vmk_peer_attach_local; vmk_step is asynchronous, so stepping rank 0 and then rank 1 runs them concurrently."""
    import ctypes as C
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import cfd_julia_b200 as vm
    from cfd_julia_b200.common import Plan
    from helpers import grid, noise_field, rel_l2, vm_field
    from oracle import oracle_c as oc
    oc.build()
    lib = vm.default_library()
    n, nt, P = 1024, 3, 2
    plans = []
    for r in range(P):
        torch.cuda.set_device(r)
        plans.append(Plan(lib, n, n, r, P))
    arr = (C.c_void_p * P)(*[p.handle for p in plans])
    for p in plans:
        lib.check(lib.peer_attach_local(p.handle, arr))
    dx, dy, _, _ = grid(n)
    w0 = vm_field(n) + 0.05 * noise_field(n, 3)
    outs = [w0.copy(order="F") for _ in plans]
    psis = [np.zeros_like(w0) for _ in plans]
    for p, o in zip(plans, outs):
        p.upload(o)
    for p in plans:
        p.step(dx, dy, 1e-3, 1000., nt)
    for p, o, s in zip(plans, outs, psis):
        p.download(o, s)
    ref = w0.copy(order="F")
    _, sref = oc.numerical(n, n, nt, dx, dy, 1e-3, 1000., ref)
    nj = n // P
    for r in range(P):
        rows = slice(r * nj, (r + 1) * nj + 2)
        assert rel_l2(outs[r][:, rows], ref[:, rows]) < 1e-10
        assert rel_l2(psis[r][:, rows], sref[:, rows]) < 1e-10
    for p in plans:
        p.close()
