"""Per-kernel-class device time of the spectral-space solvers (hybrid, ps23, ps32) from the library's own event
profiling (set_option("profile", 1) + vmk_profile_read), torch-free: ms per RK3 step and class as the difference of a
3-step and a 1-step call.  usage: python tools/profile_solvers.py [n ...]   (default 2048 8192)
Writes gpurun_out/profile_solvers.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402,F401


def main():
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    from helpers import grid, vm_field
    emul = bool(os.environ.get("VMK_QUICK_EMUL"))
    cm = Common(VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul.so"), "vmke_") if emul else
                VmkLibrary(os.path.join(ROOT, "cfd_julia_b200", "libvmk.so"), "vmk_"))
    sizes = [int(a) for a in sys.argv[1:]] or ([64] if emul else [2048, 8192])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "profile_solvers.txt"), "w") as log:
        for n in sizes:
            dx, dy, x, y = grid(n)
            w = vm_field(n)
            dt = 1e-4 if n >= 4096 else 1e-3
            for which in ("hybrid", "ps23", "ps32"):
                fn = {"hybrid": cm.numerical_hybrid, "ps23": cm.numerical_ps23, "ps32": cm.numerical_ps32}[which]
                p = cm.plan(n, n)
                fn(n, n, 1, dx, dy, dt, 1000., x, y, w, 1)  # first call: lazy allocations, not profiled
                p.set_option("profile", 1)
                p.profile_read()
                fn(n, n, 1, dx, dy, dt, 1000., x, y, w, 1)
                a = p.profile_read()
                fn(n, n, 3, dx, dy, dt, 1000., x, y, w, 1)
                b = p.profile_read()
                p.set_option("profile", 0)
                parts = {k: ((b[k]["ms"] - a[k]["ms"]) / 2, (b[k]["launches"] - a[k]["launches"]) // 2) for k in a}
                line = f"{which} {n}^2 per RK3 step: " + ", ".join(
                    f"{k} {ms:.3f} ms ({cnt} launches)" for k, (ms, cnt) in parts.items()) + \
                    f"  | sum {sum(ms for ms, _ in parts.values()):.3f} ms"
                print(line, flush=True)
                log.write(line + "\n")
            cm.clear_plans()


if __name__ == "__main__":
    main()
