"""A torch-free, seconds-long GPU check of SURVEY 8f rows f3/f4 through libvmk.so (ctypes) against the numpy oracle:
meant for a `gpurun` call with very little budget left.  Writes gpurun_out/f3_quick.txt."""
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
log = open(os.path.join(out_dir, "f3_quick.txt"), "w")


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


def main():
    t0 = time.time()
    from cfd_julia_b200._lib import VmkLibrary
    from cfd_julia_b200.common import Common
    import parity_cases as pc
    from helpers import grid, noise_field, rel_l2, vm_field
    from oracle import oracle_np as onp
    if os.environ.get("VMK_QUICK_EMUL"):  # dry run of this script on the host emulator
        cm = Common(VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul.so"), "vmke_"))
    else:
        cm = Common(VmkLibrary(os.path.join(ROOT, "cfd_julia_b200", "libvmk.so"), "vmk_"))
    say("library loaded", round(time.time() - t0, 2), "s")
    cases = [(23, 64, 5, 1.), (32, 64, 5, 1.), (23, 256, 3, .5), (32, 256, 3, .5), (23, 1024, 2, .05), (32, 1024, 2, .05)]
    for rule, n, nt, noise in cases:
        try:
            t = time.time()
            dx, dy, x, y = grid(n)
            w = vm_field(n) + noise * noise_field(n, 5)
            fn = cm.numerical_ps23 if rule == 23 else cm.numerical_ps32
            ut = fn(n, n, nt, dx, dy, 1e-3, 1000., x, y, w, 1)
            tg = time.time() - t
            ref = onp.ps_numerical(rule, n, n, nt, dx, dy, 1e-3, 1000., w)
            say(f"ps{rule} n={n} nt={nt} noise={noise}: rel-L2 vs oracle = {rel_l2(ut, ref):.3e}  (gpu call {tg:.2f} s,"
                f" launches {cm.plan(n, n).launch_count})")
        except Exception:
            say(f"ps{rule} n={n}: FAILED\n" + traceback.format_exc())
    for which in ("hybrid", "ps23", "ps32"):
        try:
            pc.check_spectral_tgv(cm, which, 64, 100)
            say(f"closed-form Taylor-Green, {which}, 64^2 x 100 steps: ok (<= 1e-12)")
        except Exception:
            say(f"closed-form Taylor-Green {which}: FAILED\n" + traceback.format_exc())
    try:
        pc.check_golden_f_rows(cm)
        say("golden fixtures (hybrid, ps23, ps32, cavity): ok")
    except Exception:
        say("golden fixtures: FAILED\n" + traceback.format_exc())
    say("total", round(time.time() - t0, 2), "s")


if __name__ == "__main__":
    main()
