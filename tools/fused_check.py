"""Recurrences along j as separate kernels (fps_mode 1) against the fused form (fps_mode 2: inside K1 / K3, state in
tensor memory) on a GPU: parity of one Poisson solve and of a short run with the C oracle, per-class kernel times
(un-graphed pass) and the graph-replayed step time.
usage: python tools/fused_check.py [sizes...]   (VMK_LIB selects the library; FZ_GRIDS="0 148 132" sweeps the grid)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cfd_julia_b200 as vm  # noqa: E402
from bench import vm_initial_condition  # noqa: E402
from helpers import rel_l2  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402

oc.build()
sizes = [int(a) for a in sys.argv[1:]] or [1024, 8192]
grids = [int(g) for g in os.environ.get("FZ_GRIDS", "0").split()]
for n in sizes:
    dx, w = vm_initial_condition(n)
    dt = min(.01, 1e-4 * (8192. / n)**2)
    f = np.asfortranarray(np.random.default_rng(n).uniform(-1, 1, (n, n)))
    ref = np.zeros((n + 2, n + 2), order="F")
    oc.fps(n, n, dx, dx, f, ref)
    nt = 2
    wr = w.copy(order="F")
    _, psir = oc.numerical(n, n, nt, dx, dx, dt, 1000., wr)
    for mode, fz in [(1, 0)] + [(2, g) for g in grids]:
        p = vm.plan(n, n)
        p.set_option("fps_mode", mode)
        if fz:
            p.set_option("fz_grid", fz)
        s = np.zeros((n + 2, n + 2), order="F")
        vm.fps(n, n, dx, dx, None, None, None, None, f, s)
        e_fps = rel_l2(s[1:n + 1, 1:n + 1], ref[1:n + 1, 1:n + 1])
        p.upload(w)
        p.step(dx, dx, dt, 1000., nt)
        wo, po = np.zeros_like(w), np.zeros_like(w)
        p.download(wo, po)
        e_w, e_p = rel_l2(wo, wr), rel_l2(po, psir)
        p.upload(w)
        p.step(dx, dx, dt, 1000., 3)
        p.sync()
        p.step(dx, dx, dt, 1000., 20)
        p.sync()
        ms = p.step_elapsed_ms() / 20
        prof = p.profile_steps(dx, dx, dt, 1000., 3)
        per = {k: round(v["ms"] / 3, 4) for k, v in prof.items()}
        print(f"n={n} fps_mode={mode} fz_grid={fz}: fps rel-L2 {e_fps:.2e}  run({nt}) w {e_w:.2e} psi {e_p:.2e}  "
              f"step {ms:.4f} ms  per-step class ms {per}", flush=True)
