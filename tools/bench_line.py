"""Condense bench.py's JSON line (stdin) to one readable row."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        k = {q: round(v["ms_per_launch"], 4) for q, v in d["roofline"]["kernels"].items()}
        print(f"gpus={d['n_gpus']} ms/step={d['ms_per_step']:.3f} value={d['value']:.4g} step_frac={d['roofline']['step']['frac']:.3f} "
              f"kernels={k} e2e={d['e2e'] and round(d['e2e']['value'] / 1e9, 2)} launches={d['gpu_launches']} clocks={d['clocks']}")
