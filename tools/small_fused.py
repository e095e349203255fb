"""Small grids: the fused cluster kernel (ks_body, one launch for the whole step loop) against the CUDA graph of 12
launches per step; device time per step (CUDA events) and parity with the oracle.  usage: python tools/small_fused.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_julia_b200 as vm  # noqa: E402
from bench import vm_initial_condition  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402

for n in (128, 256):
    dx, w = vm_initial_condition(n)
    dt = min(.01, 1e-4 * (8192. / n)**2)
    res = {}
    for fused in (0, 1):
        p = vm.plan(n, n)
        p.set_option("fuse_small", fused)
        p.upload(w)
        p.step(dx, dx, dt, 1000., 50)
        p.sync()
        p.upload(w)
        nt = 2000
        p.step(dx, dx, dt, 1000., nt)
        p.sync()
        out = np.zeros_like(w)
        p.download(out)
        res[fused] = (p.step_elapsed_ms() / nt * 1e3, out)
    ref = w.copy(order="F")
    oc.numerical(n, n, 2000, dx, dx, dt, 1000., ref)
    err = [float(np.linalg.norm(res[f][1] - ref) / np.linalg.norm(ref)) for f in (0, 1)]
    print(f"n={n:4d}  graph of 12 launches {res[0][0]:7.2f} us/step   fused cluster kernel {res[1][0]:7.2f} us/step   "
          f"identical={np.array_equal(res[0][1], res[1][1])}  rel-L2 vs oracle after 2000 steps {err[0]:.2e} / {err[1]:.2e}",
          flush=True)
