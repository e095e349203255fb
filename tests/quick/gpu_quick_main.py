"""Torch-free regression check of the main path after library changes (seconds of GPU time): the smoke case against the
C oracle, the 8192^2 step time from the plan's CUDA events, and two host threads sharing one plan (per-plan mutex) on
the real library.  Writes gpurun_out/main_quick.txt."""
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "main_quick.txt"), "w")


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
    from oracle import oracle_c as oc
    oc.build()
    emul = bool(os.environ.get("VMK_QUICK_EMUL"))
    cm = Common(VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul.so"), "vmke_") if emul else
                VmkLibrary(os.path.join(ROOT, "cfd_julia_b200", "libvmk.so"), "vmk_"))
    try:  # __graft_entry__.smoke() without torch
        n, nt = 128, 5
        dx, dy, x, y = grid(n)
        wn = vm_field(n)
        ref_w = wn.copy(order="F")
        out = cm.numerical_tgv(n, n, nt, dx, dy, .01, 1000., wn)
        ref, _ = oc.numerical(n, n, nt, dx, dy, .01, 1000., ref_w)
        say(f"smoke case 128^2 x 5 steps: rel-L2 vs C oracle = {rel_l2(out, ref):.3e}")
    except Exception:
        say("smoke case FAILED\n" + traceback.format_exc())
    try:  # two host threads, one plan
        n = 256
        dx, dy, x, y = grid(n)
        rng = np.random.default_rng(1)
        fs = [np.asfortranarray(rng.uniform(-1, 1, (n, n))) for _ in range(2)]
        refs = []
        for f in fs:
            s = np.zeros((n + 2, n + 2), order="F")
            oc.fps(n, n, dx, dy, f, s)
            refs.append(s)
        cm.plan(n, n)
        outs = [[np.zeros((n + 2, n + 2), order="F") for _ in range(8)] for _ in range(2)]

        def work(i):
            for s in outs[i]:
                cm.fps(n, n, dx, dy, None, None, None, None, fs[i], s)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        worst = max(rel_l2(s[1:n + 1, 1:n + 1], refs[i][1:n + 1, 1:n + 1]) for i in range(2) for s in outs[i])
        say(f"two host threads on one plan, 16 fps calls: worst rel-L2 vs C oracle = {worst:.3e}")
    except Exception:
        say("thread case FAILED\n" + traceback.format_exc())
    try:  # BASELINE's grid: device-resident steps
        n = 64 if emul else 8192
        dx, dy, x, y = grid(n)
        p = cm.plan(n, n)
        p.upload(vm_field(n))
        p.step(dx, dy, 1e-4, 1000., 3)
        p.sync()
        times = []
        for _ in range(3):
            p.step(dx, dy, 1e-4, 1000., 20)
            p.sync()
            times.append(p.step_elapsed_ms() / 20)
        say(f"{n}^2 RK3 step (20-step calls, CUDA events): {' / '.join(f'{t:.3f}' for t in times)} ms per step"
            f"  -> {n * n / (min(times) * 1e-3):.4e} grid-point-steps/s")
    except Exception:
        say("step timing FAILED\n" + traceback.format_exc())
    say("total", round(time.time() - t0, 2), "s")


if __name__ == "__main__":
    main()
