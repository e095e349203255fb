"""Per-kernel times of the cluster sizes (16384, 32768) with the L2 prefetch on / off.  GPU box only."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import cfd_julia_b200 as vm  # noqa: E402
from cfd_julia_b200.common import Plan  # noqa: E402

lib = vm.default_library()
for n in [int(a) for a in sys.argv[1:]] or [16384]:
    dx, w0 = bench.vm_initial_condition(n)
    dt = 1e-4 * (8192. / n)**2
    p = Plan(lib, n, n)
    p.upload(w0)
    for pf in (-1, 0, 7):
        p.set_option("cl_prefetch", pf)
        p.step(dx, dx, dt, 1000., 3)
        p.sync()
        p.step(dx, dx, dt, 1000., 5)
        p.sync()
        ms = p.step_elapsed_ms() / 5
        prof = p.profile_steps(dx, dx, dt, 1000., 2)
        print(json.dumps({"n": n, "cl_prefetch": pf, "ms_per_step": ms,
                          "frac_hbm": 232. * n * n / (ms * 1e-3) / 1e9 / bench.measured_peak()[0],
                          **{k: round(v["ms"] / v["launches"], 4) for k, v in prof.items()}}), flush=True)
    p.close()
