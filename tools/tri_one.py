"""One un-graphed vortex-merger step with the recurrence form of the solve along j, for `ncu` launch lists.
usage: python tools/tri_one.py [n] [fps_mode] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_julia_b200 as vm  # noqa: E402
from bench import vm_initial_condition  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dx, w = vm_initial_condition(n)
p = vm.plan(n, n)
p.set_option("fps_mode", mode)
p.set_option("graph", 0)
p.upload(w)
p.step(dx, dx, 1e-4 * (8192. / n)**2, 1000., steps)
p.sync()
print("launches", p.launch_count)
