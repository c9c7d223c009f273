#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python scripts/launch_summary.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui].replace("second", "s"), v / 1e3)
        agg[r[ki][:100]][0] += 1
        agg[r[ki][:100]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time "
          "(ncu-serialised, cold caches: compare SHARES, not absolutes)")
    print(f"{'us':>10s} {'calls':>6s} {'share':>6s}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{t:10.1f} {c:6d} {100 * t / tot:5.1f}%  {n}")


if __name__ == "__main__":
    main()
