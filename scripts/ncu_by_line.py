#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

    python scripts/ncu_by_line.py <report.ncu-rep> <kernel regex> <cubin> <mangled name> [top]

Joins `ncu --page source --csv` (SASS order, with "Instructions Executed" and stall samples) with `nvdisasm -g`
line annotations of the same function (instruction order is identical)."""
import collections
import csv
import io
import re
import subprocess
import sys

SORT = 1


def main():
    rep, kre, cubin, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    global SORT
    SORT = 0 if (len(sys.argv) > 6 and sys.argv[6] == "inst") else 1
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel instance only
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_i[0]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    sass = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    lines, cur, inside = [], ("?", 0), False
    for ln in dis.splitlines():
        if ln.lstrip().startswith(".section"):
            inside = (".text." + mangled) in ln and "," in ln and ln.split(",")[0].strip().endswith(mangled)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if len(lines) != len(sass):
        print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for loc, r in zip(lines, sass):
        inst = int(r[ix["Instructions Executed"]] or 0)
        smp = int(r[ix["# Samples"]] or 0)
        agg[loc][0] += inst
        agg[loc][1] += smp
        tot_i += inst
        tot_s += smp
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for loc, (inst, smp) in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
        print(f"{loc[0]:18s}:{loc[1]:4d}  inst {inst:12d} {100 * inst / max(tot_i, 1):5.1f}%   samples {smp:7d} {100 * smp / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main()
