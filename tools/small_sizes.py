"""Step time of the device-resident loop at the small grids (launch-bound regime), incl. config 1 (vm.jl defaults)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_julia_b200 as vm  # noqa: E402
from bench import vm_initial_condition  # noqa: E402

for n in (32, 64, 128, 256, 512, 1024, 2048):
    dx, w = vm_initial_condition(n)
    p = vm.plan(n, n)
    p.upload(w)
    dt = min(.01, 1e-4 * (8192. / n)**2)
    p.step(dx, dx, dt, 1000., 50)
    p.sync()
    nt = 2000 if n <= 512 else 200
    t0 = time.perf_counter()
    p.step(dx, dx, dt, 1000., nt)
    p.sync()
    wall = time.perf_counter() - t0
    print(f"n={n:5d}  device {p.step_elapsed_ms() / nt * 1e3:8.1f} us/step   wall {wall / nt * 1e6:8.1f} us/step   "
          f"{n * n * nt / wall:.3e} grid-point-steps/s")
t0 = time.perf_counter()
dx, w = vm_initial_condition(128)
out = vm.numerical_tgv(128, 128, 2000, dx, dx, .01, 1000., w)
print(f"config 1 (vm.jl defaults: 128^2, 2000 steps) through numerical(): {time.perf_counter() - t0:.3f} s")
