"""Turn an .ncu-rep (ncu --set full) into the committed evidence: profiles/<tag>_ncu_summary.md and traffic.json.
usage: python tools/ncu_summary.py gpurun_out/r01_prof.ncu-rep r01 ["title"]   (a title also keeps traffic.json as is)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1TEX LSU wavefronts % (shared+global)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  of which shared memory %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, body = rows[0], rows[1], rows[2:]
    names = []
    for r in body:
        m = re.search(r"K(\d)(C?)Body(?:<\(int\)(\d)>)?", r[h.index("Kernel Name")])
        names.append(f"k{m.group(1)}" + ("c" if m.group(2) else "") + (f" (stage {m.group(3)})" if m.group(3) else ""))
    title = sys.argv[3] if len(sys.argv) > 3 else "one RK3 step at 8192^2 (12 launches), `python bench.py --steps 2 --warmup 3`"
    out = [f"# ncu --set full, {tag}: {title}",
           "", "Cold-cache, serialised launches under the profiler: compare shares and counters, not absolute times.", "",
           "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for key, label in WANT:
        if key not in h:
            continue
        i = h.index(key)
        out.append(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in body) + " |")
    traffic = {}
    ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    for nm, r in zip(names, body):
        k = nm.split()[0]
        traffic.setdefault(k, []).append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
    traffic = {k: sum(v) / len(v) for k, v in traffic.items()}
    out += ["", "DRAM traffic per launch (read + write, mean over the captured launches), bytes:", "",
            "```", json.dumps(traffic, indent=1), "```"]
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    if len(sys.argv) <= 3:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    print("\n".join(out[-8:]))


if __name__ == "__main__":
    main()
