"""Smallest end-to-end case for compute-sanitizer: 128^2 vortex merger, 2 RK3 steps + one vm_rhs + one fps call."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_julia_b200 as vm  # noqa: E402
from bench import vm_initial_condition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dx, w = vm_initial_condition(n)
p = vm.plan(n, n)
p.set_option("graph", 0)
out = vm.numerical_tgv(n, n, 2, dx, dx, .01, 1000., w)
r = np.zeros_like(w)
s = np.zeros_like(w)
f = np.zeros((n, n), order="F")
vm.vm_rhs(n, n, dx, dx, 1000., w, None, None, None, None, r, s, f)
vm.fps(n, n, dx, dx, None, None, None, None, f, s)
print("ok", float(np.abs(out).max()), float(np.abs(s).max()))
