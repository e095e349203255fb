#!/usr/bin/env python
"""bench.py -- vortex-merger grid-point-steps/s at 8192^2 FP64 on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (libvmk.so through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU algorithm on the host cores

One "step" = one SSP-RK3 time step of the 8192^2 periodic vortex merger (vm.jl:24-76): 3 x (Poisson solve +
Arakawa/Laplacian rhs + stage combine) = 12 kernel launches.  Re = 1000, dt = 1e-4 (the scripts' dt = .01 is
unstable at this resolution).  The working set (~3.2 GB) is far larger than the 126 MB L2, so no
flush is needed between timed steps.

JSON line keys beyond the base contract:
  roofline     the slowest kernel of the step: algorithmic bytes per launch / mean launch duration (CUDA
               events on the plan's stream, measured here) against MEASURED_PEAKS.json's HBM copy bandwidth;
               `step` holds the same for the whole step (232 B per grid-point-step)
  cpu_baseline the oracle port (oracle/vm_oracle.c, OpenMP) timed on this box's host cores on a bounded sample
  e2e          vmk_numerical(nt = K) on a pinned HOST array: upload + K steps + download inside the timed region
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def metric_name(n):
    """BASELINE.json's metric, named after the grid the run actually used."""
    return f"vortex_merger_grid_point_steps_per_s_{n}x{n}_fp64"


UNIT = "grid-point-steps/s"
BYTES_PER_POINT_STEP = 232.0
# algorithmic bytes per grid point per launch: K1/K2/K3 read 8 + write 8; K4 reads w, psi (+wn) and writes
KERNEL_BYTES = {"k1": 16.0, "k2": 16.0, "k3": 16.0, "k4": (24.0 + 32.0 + 32.0) / 3.0,
                # recurrence form of the solve along j (csrc/vmk_tri.cuh): totals read the spectrum; the scan reads 3 and
                # writes 2 complex numbers per chunk of 32 points; the solve reads and writes the spectrum; K2 transforms
                # K0 = 64 of the N/2 spectrum rows (latency, not bandwidth: it runs beside the others)
                "kt_totals": 8.0 + 48.0 / 32, "kt_scan": 80.0 / 32, "kt_solve": 16.0 + 32.0 / 32, "k2_low_rows": 0.25}
RE, DT = 1000., 1e-4


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in rows:
            p = [q.strip() for q in l.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except Exception:  # noqa: BLE001
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def vm_initial_condition(n, j0=0, nj=None):
    """vm_ic + main()'s ghost fill (Common.jl:208-219, vm.jl:107-128) -- synthetic input, no files.
    Only the ghosted columns j0 .. j0+nj+1 (a rank's slab and its halo columns) are evaluated; the rest of the
    (lazily committed) array stays untouched, so that 8 ranks at 32768^2 do not each build an 8.6 GB field."""
    dx = 2 * np.pi / n
    nj = n if nj is None else nj
    w = np.zeros((n + 2, n + 2), order="F")
    xg = dx * ((np.arange(n + 2) - 1) % n)[:, None]          # ghosted index -> periodic node coordinate
    # the slab's ghosted columns, plus the INTERIOR columns its halos wrap to (vmk_upload imposes periodicity from
    # the interior of the caller's array, not from its ghost cells)
    cols = np.unique(np.concatenate([np.arange(j0, j0 + nj + 2), [(j0 - 1) % n + 1, (j0 + nj) % n + 1]]))
    for jb in range(0, len(cols), 512):                       # column blocks bound the temporaries
        cb = cols[jb:jb + 512]
        y = dx * ((cb - 1) % n)[None, :]
        w[:, cb] = (np.exp(-np.pi * ((xg - 3 * np.pi / 4)**2 + (y - np.pi)**2)) +
                    np.exp(-np.pi * ((xg - 5 * np.pi / 4)**2 + (y - np.pi)**2)))
    return dx, w


def host_threads():
    """Hardware threads this process may use (affinity mask), ignoring OMP_NUM_THREADS: torch.distributed.run
    exports OMP_NUM_THREADS=1 to its workers, which made round 1's reference arm run on one thread."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return max(1, os.cpu_count() or 1)


def cpu_sample(n, steps, warmup=0, threads=None):
    """The oracle port on the host cores with an explicit OpenMP thread count:
    returns (grid-point-steps/s, threads, seconds per step)."""
    from oracle import oracle_c as oc
    oc.build()
    oc.set_num_threads(threads or host_threads())
    dx, w = vm_initial_condition(n)
    dt = min(.01, DT * (8192. / n)**2)
    if warmup:
        oc.numerical(n, n, warmup, dx, dx, dt, RE, w)
    t0 = time.perf_counter()
    oc.numerical(n, n, steps, dx, dx, dt, RE, w)
    el = time.perf_counter() - t0
    return n * n * steps / el, oc.num_threads(), el / steps


def julia_probe(n, dt):
    """the real reference, if this box has it.  `julia` on PATH + CFD_JULIA_REFERENCE pointing at a
    checkout of the reference -> time vm.jl's own `numerical` through oracle/julia_shim/time_vm.jl (1 Julia thread,
    as the reference runs).  Otherwise say what is missing; never fatal."""
    exe = shutil.which("julia")
    if not exe:
        return {"julia": "absent", "which": None}
    ref = os.environ.get("CFD_JULIA_REFERENCE", "")
    if not os.path.isdir(ref):
        return {"julia": "present, reference checkout absent (set CFD_JULIA_REFERENCE)", "which": exe}
    try:
        out = subprocess.run([exe, "-t", "1", os.path.join(ROOT, "oracle", "julia_shim", "time_vm.jl"), str(n), "1",
                              repr(dt)], capture_output=True, text=True, timeout=900)
        last = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode == 0 and last:
            return {"julia": "ran", "which": exe, **json.loads(last[-1])}
        return {"julia": f"failed rc={out.returncode}", "which": exe, "stderr": out.stderr[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"julia": f"failed: {e}", "which": exe}


def run_reference(args):
    """--impl reference: the reference's algorithm on the CPU at the SAME grid as the b200 arm.  Julia + FFTW.jl are
    not in the image, so this is the oracle port (kind "port") with all host threads (OpenMP, count set explicitly);
    a 1-thread figure (the reference itself is single-threaded, vm.jl:24) is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n
    nthr = host_threads()
    # bounded: at 8192^2 one RK3 step of the port is ~2 s on 16 threads; cap the run at ~3 minutes
    dt = min(.01, DT * (8192. / n)**2)
    v1 = t1 = None
    if not args.no_cpu_1thread:
        # one step on ONE thread, on a grid small enough to stay below ~30 s (per-point cost is size-independent
        # to within the log N of the FFT)
        n1 = min(n, 4096)
        v1, _, t1 = cpu_sample(n1, 1, 0, threads=1)
    steps, warm = args.steps, min(args.warmup, 1)
    v, threads, s_per_step = cpu_sample(n, 1, 0, threads=nthr)  # probe one step
    budget_s = 170.
    steps = int(max(1, min(steps, budget_s // max(s_per_step, 1e-9))))
    if steps > 1:
        v, threads, s_per_step = cpu_sample(n, steps, warm, threads=nthr)
    line = {
        "impl": "reference", "metric": metric_name(n), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, dt),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"oracle/vm_oracle.c (OpenMP, {threads} threads set explicitly; OMP_NUM_THREADS of the "
                                   f"launcher ignored), {steps} RK3 steps of the same {n}x{n} workload, "
                                   f"{s_per_step:.2f} s/step; Julia/FFTW.jl absent from the image",
                         "value_1thread": v1,
                         "sample_1thread": None if v1 is None else
                         f"1 RK3 step at {min(n, 4096)}x{min(n, 4096)} on 1 thread ({t1:.1f} s); the Julia reference is "
                         "single-threaded (vm.jl:24)",
                         "host_threads": nthr, **julia_probe(min(n, 2048), min(.01, DT * (8192. / min(n, 2048))**2))},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n, dt):
    return {"workload": f"vortex merger {n}x{n} periodic, Re=1000, dt={dt:.3g}, RK3 + FFT Poisson "
                        + ("(BASELINE configs[3])" if n == 8192 else
                           "(BASELINE configs[4])" if n == 32768 else "(not the headline size)")}


def parity_probe(lib, Plan, n, world, rank, gather, steps=3, fps_mode=None):
    """Outside the timed region: `steps` RK3 steps of an n x n vortex merger on the SAME rank layout as the benchmark,
    compared with the C oracle on every rank's slab -- so that the scaling runs carry multi-GPU correctness."""
    from oracle import oracle_c as oc
    oc.build()
    oc.set_num_threads(max(1, host_threads() // max(world, 1)))
    dx, w0 = vm_initial_condition(n)
    dt = min(.01, 1e-4 * (8192. / n)**2)
    plan = Plan(lib, n, n, rank, world)
    if world > 1:
        plan.attach_peers(gather)
    if fps_mode is not None:
        plan.set_option("fps_mode", fps_mode)
    plan.upload(w0)
    plan.step(dx, dx, dt, RE, steps)
    plan.sync()
    w = np.zeros_like(w0)
    psi = np.zeros_like(w0)
    plan.download(w, psi)
    ref = w0.copy(order="F")
    _, psi_ref = oc.numerical(n, n, steps, dx, dx, dt, RE, ref)
    nj = n // world
    sl = (slice(None), slice(rank * nj + 1, rank * nj + nj + 1))  # the rank's own interior columns, all i
    ew = float(np.linalg.norm(w[sl] - ref[sl]) / np.linalg.norm(ref[sl]))
    ep = float(np.linalg.norm(psi[sl] - psi_ref[sl]) / np.linalg.norm(psi_ref[sl]))
    return plan, {"n": n, "steps": steps, "rel_l2_w": ew, "rel_l2_psi": ep}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=8192,
                    help="grid size (default: the BASELINE workload); use --size under torchrun, whose own parser "
                         "takes a bare --n for an abbreviation of its options")
    ap.add_argument("--no-cpu-1thread", action="store_true", help="reference arm: skip the 1-thread sample")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed 2048^2 x 3-step oracle comparison")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import cfd_julia_b200 as vm
    from cfd_julia_b200.common import Plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    lib = vm.default_library()
    n, K, W = args.n, args.steps, args.warmup
    world_ = int(os.environ.get("WORLD_SIZE", "1"))
    rank_ = int(os.environ.get("RANK", "0"))
    dx, w0 = vm_initial_condition(n, rank_ * (n // world_), n // world_)
    global DT
    if n > 8192:
        DT = DT * (8192. / n)**2  # diffusive RK3 limit ~ dx^2 (1.155e-5 at 32768^2)

    dist = None
    keep = []
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        plan = Plan(lib, n, n, rank, world)

        def _gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        plan.attach_peers(_gather)  # CUDA IPC handles; the data path itself uses no NCCL call
        dist.barrier()
    else:
        plan = Plan(lib, n, n)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    plan.upload(w0)
    plan.step(dx, dx, DT, RE, W)  # warm-up (also builds the CUDA graph)
    plan.sync()

    # ---- timed region: K steps, device-resident, CUDA events on the plan's stream -----------------------------
    clocks = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    l0 = plan.launch_count
    t0 = time.time()
    plan.step(dx, dx, DT, RE, K)
    plan.sync()
    barrier()
    t1 = time.time()
    ms = plan.step_elapsed_ms()
    launches = plan.launch_count - l0
    clk = clocks.stop(t0, t1) if clocks else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tl = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(tl)
        launches = int(tl.item())
    value = float(n) * n * K / (ms * 1e-3)

    # sanity: the field is finite after the run (a wrong dt would give NaNs and a meaningless number)
    chk = np.zeros_like(w0)
    plan.download(chk)
    j0, nj = rank * (n // world), n // world
    assert np.isfinite(chk[:, j0:j0 + nj + 2]).all(), "non-finite vorticity after the timed steps"

    # ---- per-kernel durations (events around every launch; separate, untimed pass) ---------------------------------
    peak, peak_src = measured_peak()
    prof = plan.profile_steps(dx, dx, DT, RE, 3)
    pts = float(n) * (n // world)
    kern = {}
    tri = plan.profile_tri()
    fused = tri["kt_scan"]["launches"] and not tri["kt_solve"]["launches"]
    fps_mode = ("recurrences along j inside K1 / K3, per-slot state in tensor memory (csrc/vmk_tri.cuh, fused form)" if fused
                else "recurrences along j (csrc/vmk_tri.cuh)" if tri["kt_solve"]["launches"] else "FFT along j (K2)")
    if tri["kt_scan"]["launches"]:
        # the "k2" class of the recurrence form = chunk totals (reads the spectrum: 8 B/point) + scan (chunk totals and
        # carries: 80 B per 32 points) + in-place solve (16 B/point) + K2 on the rows kx < K0 (beside them, on a
        # second stream; timed in line here): list the three streaming kernels separately
        rest = prof["k2"]["ms"] - sum(v["ms"] for v in tri.values())
        nlow = prof["k2"]["launches"] - sum(v["launches"] for v in tri.values())
        prof = dict(prof)
        prof.pop("k2")
        prof.update({k: v for k, v in tri.items() if v["launches"]})
        prof["k2_low_rows"] = {"ms": rest, "launches": nlow}
    for k, v in prof.items():
        if v["launches"]:
            dur = v["ms"] / v["launches"]
            kb = KERNEL_BYTES[k]
            if fused and k == "kt_scan":  # 3 complex totals read, 2 carries written per block of rows (one per CTA) and kx
                kb = 80.0 * (tri["kt_scan"].get("units") or 148) / (2.0 * n)
            ach = kb * pts / (dur * 1e-3) / 1e9
            kern[k] = {"ms_per_launch": dur, "achieved_gbs": ach, "frac": ach / peak,
                       "share": v["ms"] / sum(q["ms"] for q in prof.values())}
    top = max(kern, key=lambda k: kern[k]["ms_per_launch"]) if kern else None
    step_gbs = BYTES_PER_POINT_STEP * value / world / 1e9
    roofline = {"bound": "hbm", "kernel": top, "achieved": kern[top]["achieved_gbs"] if top else None, "peak": peak,
                "unit": "GB/s", "frac": kern[top]["frac"] if top else None, "traffic": None, "peak_source": peak_src,
                "kernels": kern, "solve_along_j": fps_mode,
                "step": {"achieved": step_gbs, "frac": step_gbs / peak, "frac_of_8TBs": step_gbs / 8000.,
                         "bytes_per_point_step": BYTES_PER_POINT_STEP,
                         "bytes_note": "232 = the FFT x FFT formulation's algorithmic traffic (SURVEY 8d), kept as the "
                                       "common numerator; the recurrence form as separate kernels moves 8 B/point more "
                                       "per solve (256 per step), the fused form 16 B/point less (184 per step)"}}
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this command at 8192^2 on one
    # GPU (profiles/traffic.json, tools/ncu_summary.py); a capture exists for that configuration only
    roofline["kernel_times"] = ("CUDA events around every launch in a separate un-graphed pass (vmk_profile_steps); "
                                "their sum can exceed the graph-replayed step, use them as shares")
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if top and world == 1 and n == 8192 and os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(top)
        except Exception:  # noqa: BLE001
            pass

    # ---- e2e: the reference-facing call on a pinned host array ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n + 2) * (n + 2), dtype=torch.float64).pin_memory()
        hv = host.numpy().reshape((n + 2, n + 2), order="F")
        hv[...] = w0
        barrier()
        te = time.perf_counter()
        lib.check(lib.numerical(plan.handle, K, dx, dx, DT, RE, host.data_ptr(), None, 0,
                                vm._lib.SNAPSHOT_FN(), None))
        barrier()
        el = time.perf_counter() - te
        if dist is not None:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        slab_bytes = 8 * (n + 2) * (n // world + 2)
        e2e = {"value": float(n) * n * K / el, "unit": UNIT, "h2d_bytes_per_step": slab_bytes * world / K,
               "d2h_bytes_per_step": slab_bytes * world / K,
               "call": f"vmk_numerical(nt={K}) on a pinned host array: upload + {K} steps + download"}

    # ---- parity, outside the timed region, on the same rank layout (all ranks take part) -------------------------------
    parity = None
    if not args.no_parity:
        def _gather2(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        pn = min(n, 2048)  # the smallest grid that takes the same default code path as the benchmarked one
        # (the fused form of the recurrences is the default at 8192^2 on one GPU only: ask for it at the probe's size)
        pmode = 2 if (world == 1 and pn >= 512 and "fused form" in roofline.get("solve_along_j", "")) else None
        pplan, parity = parity_probe(lib, Plan, pn, world, rank, _gather2 if dist is not None else None, fps_mode=pmode)
        parity["solve_along_j"] = roofline.get("solve_along_j")
        if dist is not None:
            t = torch.tensor([parity["rel_l2_w"], parity["rel_l2_psi"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity["rel_l2_w"], parity["rel_l2_psi"] = float(t[0].item()), float(t[1].item())
            dist.barrier()
        pplan.close()
        parity["against"] = "oracle/vm_oracle.c, max over ranks of each rank's own slab"
        parity["ok"] = bool(parity["rel_l2_w"] < 1e-10 and parity["rel_l2_psi"] < 1e-10)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, sps = cpu_sample(n, 2)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle/vm_oracle.c (OpenMP, {threads} threads set explicitly): 2 RK3 steps of the same {n}x{n} "
                         f"workload, {sps:.2f} s/step; Julia/FFTW.jl absent from the image",
               **julia_probe(min(n, 2048), min(.01, 1e-4 * (8192. / min(n, 2048))**2))}

    if rank == 0:
        line = {
            "metric": metric_name(n), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {**workload_config(n, DT),
                       "l2": f"working set {plan.device_bytes / 1e9:.1f} GB >> 126 MB L2, no flush needed",
                       "parallelism": f"slab{world}" if world > 1 else "single GPU",
                       "exchange": (("no transposes: per Poisson solve every rank stores 3 complex numbers per kx (rank "
                                     "totals of the recurrences) and its columns of the K0 = 64 FFT-form rows into each "
                                     "peer over NVLink, + halo rows of psi and w; "
                                     if roofline.get("solve_along_j", "").startswith("rec") else
                                     "all-to-all transposes as peer stores over NVLink inside K1/K2, halo rows; ")
                                    + "device-side flag barriers, no NCCL call on the data path")
                       if world > 1 else None,
                       "cuda_graph": True},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "parity": parity,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        plan.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
