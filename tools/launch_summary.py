#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for one step
(steps are delimited by the byte_presence kernel that opens every pipeline step).
usage: python tools/launch_summary.py launches.csv [step_index] > profiles/...txt"""
import collections
import csv
import re
import sys


def short(k):
    k = k.replace("void ", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([\w:~]+)(<[^>]{0,48})?", k)
    return ((m.group(1) + (m.group(2) or "")) if m else k)[:90]


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]
    marks = [i for i, (k, _) in enumerate(rows) if "byte_presence" in k] + [len(rows)]
    a, b = marks[step], marks[step + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows[a:b]:
        agg[short(k)][0] += 1
        agg[short(k)][1] += v
    tot = sum(v for _, v in agg.values())
    ours = sum(v for k, (_, v) in agg.items() if not k.startswith(("at::", "cutlass", "cublas", "epilogue", "nccl")))
    print(f"step {step}: {b - a} launches, {tot / 1e6:.3f} ms GPU time (ncu gpu__time_duration.sum; cold-cache, serialised => compare shares)")
    print(f"libpgb200 kernels: {ours / 1e6:.3f} ms ({100 * ours / tot:.1f}%)")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v / 1e3:10.1f} us {c:5d}x {100 * v / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
