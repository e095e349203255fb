"""Torch-free full-size check and first timings of the spectral-space solvers through libvmk.so:
hybrid / ps23 / ps32 at 2048^2 and 8192^2, device time per step from the plan's CUDA events (difference of a 3-step and
a 1-step call, so the initial transform and the final field cancel), finite / mean-free / periodic-duplicate checks and
the agreement of the two de-aliasing rules.  Writes gpurun_out/f3_big.txt.  Seconds of GPU time."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "f3_big.txt"), "w")


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


def main():
    t0 = time.time()
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    from helpers import grid, rel_l2, vm_field
    if os.environ.get("VMK_QUICK_EMUL"):
        cm = Common(VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul.so"), "vmke_"))
        sizes = (64, 128)
    else:
        cm = Common(VmkLibrary(os.path.join(ROOT, "cfd_julia_b200", "libvmk.so"), "vmk_"))
        sizes = (2048, 8192)
    for n in sizes:
        dx, dy, x, y = grid(n)
        w = vm_field(n)
        dt = 1e-4 if n >= 4096 else 1e-3
        res = {}
        for which in ("hybrid", "ps23", "ps32"):
            try:
                fn = {"hybrid": cm.numerical_hybrid, "ps23": cm.numerical_ps23, "ps32": cm.numerical_ps32}[which]
                p = cm.plan(n, n)
                l0 = p.launch_count
                fn(n, n, 1, dx, dy, dt, 1000., x, y, w, 1)
                t1, l1 = p.step_elapsed_ms(), p.launch_count
                ut = fn(n, n, 3, dx, dy, dt, 1000., x, y, w, 1)
                t3, l3 = p.step_elapsed_ms(), p.launch_count
                ok = bool(np.isfinite(ut).all()) and abs(float(ut[:n, :n].mean())) < 1e-12 and \
                    np.array_equal(ut[n, :], ut[0, :]) and np.array_equal(ut[:, n], ut[:, 0])
                res[which] = ut
                say(f"{which} {n}^2: {(t3 - t1) / 2:.3f} ms per RK3 step (3-step call {t3:.2f} ms, 1-step call {t1:.2f} ms),"
                    f" {((l3 - l1) - (l1 - l0)) // 2} launches per step, device memory {p.device_bytes / 2**30:.2f} GiB,"
                    f" checks {'ok' if ok else 'FAILED'}")
            except Exception:
                say(f"{which} {n}^2: FAILED\n" + traceback.format_exc())
        if "ps23" in res and "ps32" in res:
            say(f"ps23 vs ps32 {n}^2 after 3 steps: rel-L2 = {rel_l2(res['ps32'], res['ps23']):.3e}")
        if "ps23" in res and "hybrid" in res:
            say(f"hybrid vs ps23 {n}^2 after 3 steps: rel-L2 = {rel_l2(res['hybrid'], res['ps23']):.3e} (different discretisations)")
        cm.clear_plans()
    say("total", round(time.time() - t0, 2), "s")


if __name__ == "__main__":
    main()
