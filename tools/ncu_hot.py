"""Summarise `ncu --page source --csv --print-source sass` output: per kernel, where the warp-stall samples sit.

usage: python tools/ncu_hot.py <source.csv> [kernel-substring] [top-N]
Prints, for every kernel section, the stall samples grouped by opcode class, and the top-N instructions with a
window of the surrounding SASS so that the phase of the kernel can be recognised.
"""
import csv
import re
import sys
from collections import Counter


def sections(path):
    cur, name = [], None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if row and row[0] == "Kernel Name":
                if cur:
                    yield name, cur
                name, cur = row[1], []
            elif row:
                cur.append(row)
    if cur:
        yield name, cur


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    seen = set()
    for name, rows in sections(path):
        if want not in name or name in seen:
            continue
        seen.add(name)
        hdr, body = rows[0], rows[1:]
        isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        tot = sum(int(r[isamp] or 0) for r in body)
        print("=" * 100)
        print(name[:140])
        print(f"instructions {len(body)}, samples {tot}")
        byop = Counter()
        for r in body:
            op = r[isrc].split()
            op = [o for o in op if not o.startswith("@")]
            byop[re.sub(r"\..*", "", op[0]) if op else "?"] += int(r[isamp] or 0)
        print("samples by opcode:", ", ".join(f"{k}:{v}" for k, v in byop.most_common(14)))
        order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
        for i in sorted(order):
            r = body[i]
            print(f"  [{i:5d}] samples {int(r[isamp]):6d} exec {r[iexec]:>8s}  {r[isrc].strip()[:90]}")


if __name__ == "__main__":
    main()
