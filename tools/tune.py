"""Per-kernel timing sweep over the plan's tuning knobs (vmk_set_option) on the bench workload.
usage (GPU box): python tools/tune.py [n] -- prints ms per launch of K1..K4 for each setting."""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfd_julia_b200 as vm  # noqa: E402
from bench import DT, RE, vm_initial_condition  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dx, w0 = vm_initial_condition(n)
    p = vm.plan(n, n)
    p.upload(w0)

    def measure(tag):
        p.step(dx, dx, DT, RE, 2)
        p.sync()
        prof = p.profile_steps(dx, dx, DT, RE, 3)
        p.step(dx, dx, DT, RE, 5)
        p.sync()
        ms = p.step_elapsed_ms() / 5
        print(f"{tag:48s} " + " ".join(f"{k}={v['ms'] / max(v['launches'], 1):.4f}" for k, v in prof.items()) +
              f"  step={ms:.3f} ms", flush=True)

    measure("defaults")
    for arg in sys.argv[2:]:
        for kv in arg.split(","):
            k, v = kv.split("=")
            p.set_option(k, int(v))
        measure(arg)


if __name__ == "__main__":
    main()
