#!/usr/bin/env python
"""HBM-bound primitives against the measured HBM roofline: stand-alone message-input gather, deterministic
segmented reduce, graph plan build and the residue-graph featurizer.

    python scripts/prim_microbench.py [--edges 10000000] [--dims ck|mb]

Algorithmic bytes per edge are those of SURVEY.md §8(d) / DESIGN.md §3.5; time = CUDA events around `iters`
back-to-back calls on inputs far larger than the 126 MB L2.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import caster_dta_b200 as cg
from caster_dta_b200 import ops, synth

DIMS = {"ck": ((16, 4), (32, 1)), "mb": ((100, 16), (32, 1))}


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, default=10_000_000)
    ap.add_argument("--dims", default="ck")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--featurize-pairs", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda")
    peak = 6650.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    (ns, nv), (es, ev) = DIMS[args.dims]
    ei_np, n = synth.conv_microbench_graph(args.edges, 30)
    ei = torch.from_numpy(ei_np).to(dev)
    e = ei.shape[1]
    kbar = e / n
    R, Q, I = 4 * ns + 12 * nv, 4 * es + 12 * ev, 16
    x = (torch.randn(n, ns, device=dev), torch.randn(n, nv, 3, device=dev))
    ea = (torch.randn(e, es, device=dev), torch.randn(e, ev, 3, device=dev))
    res = {"dims": args.dims, "nodes": n, "edges": e, "hbm_peak_GBps": peak, "kernels": {}}

    def rec(name, ms, nbytes, units=e):
        res["kernels"][name] = {"ms": ms, "units_per_s": units / (ms * 1e-3), "algorithmic_GBps": nbytes / (ms * 1e-3) / 1e9,
                                "frac_of_measured_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak}

    ms = timed(lambda: ops.gather_message_input(ei, x, ea), args.iters)
    rec("gather_message_input", ms, e * (I + 2 * Q + 2 * R + R / kbar))
    plan = ops.GraphPlan(ei, n)
    ms = timed(lambda: ops.GraphPlan(ei, n), args.iters)
    rec("plan_build", ms, e * (16 + 6 * 4) + 2 * 4 * (n + 1))
    rows = torch.randn(e, ns + 3 * nv, device=dev)
    out = torch.empty(n, ns + 3 * nv, device=dev)
    ms = timed(lambda: ops.segment_reduce(rows, plan, "sum", use_perm=False, out=out), args.iters)
    rec("segment_reduce(sorted rows)", ms, e * (R + R / kbar) + 4 * n)
    ms = timed(lambda: ops.segment_reduce(rows, plan, "mean", use_perm=True, out=out), args.iters)
    rec("segment_reduce(via perm)", ms, e * (R + 4 + R / kbar) + 4 * n)
    # featurizer: Davis-shape proteins, kNN-30 and radius 4 A
    pb = synth.protein_batch_coords("davis", args.featurize_pairs, 9)
    coords, ptr = torch.from_numpy(pb["coords"]).to(dev), torch.from_numpy(pb["ptr"]).to(dev)
    for tt, th in (("num", 30), ("dist", 4.0)):
        eif, _, _ = cg.residue_graph_batch(coords, ptr, th, tt, True)
        ef = eif.shape[1]
        ms = timed(lambda: cg.residue_graph_batch(coords, ptr, th, tt, True), 3)
        rec(f"featurize({tt}={th})", ms, 12 * coords.shape[0] + ef * (Q + I), units=ef)
        res["kernels"][f"featurize({tt}={th})"]["residues"] = int(coords.shape[0])
        res["kernels"][f"featurize({tt}={th})"]["edges"] = int(ef)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
