"""CPU suite: the N > 1 path as separate processes over gloo (see tests/dist_cpu_worker.py)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT


@pytest.mark.parametrize("world", [2, 4])
def test_slab_processes_gloo(world):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "-s"])
    from oracle import oracle_c
    oracle_c.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
