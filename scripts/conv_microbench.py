#!/usr/bin/env python
"""Isolated GVPConvLayer micro-benchmark (BASELINE.json config 5) and kernel-profiling target.

    python scripts/conv_microbench.py --dims ck --edges 684000 --iters 5 [--backward] [--graph knn|micro]

Reports edges/s of the fused conv kernels (device time from the in-library event bracketing) and their fraction
of the HBM roofline with the algorithmic bytes of DESIGN.md §4.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

import caster_dta_b200 as cg
from caster_dta_b200 import _lib, synth

DIMS = {"ck": ((16, 4), (32, 1)), "mb": ((100, 16), (32, 1))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="ck")
    ap.add_argument("--edges", type=int, default=684000)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--backward", action="store_true")
    ap.add_argument("--aggr", default="sum")
    ap.add_argument("--tc", action="store_true", help="tcgen05 bf16 message GEMMs (config-5 dims)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    nd, ed = DIMS[args.dims]
    ei_np, n = synth.conv_microbench_graph(args.edges, args.k)
    ei = torch.from_numpy(ei_np).to(dev)
    e = ei.shape[1]
    torch.manual_seed(9)
    layer = cg.GVPConvLayer(nd, ed, drop_rate=0.0, activations=(F.relu, None), vector_gate=True, aggr=args.aggr).to(dev)
    x = (torch.randn(n, nd[0], device=dev, requires_grad=args.backward), torch.randn(n, nd[1], 3, device=dev, requires_grad=args.backward))
    ea = (torch.randn(e, ed[0], device=dev, requires_grad=args.backward), torch.randn(e, ed[1], 3, device=dev, requires_grad=args.backward))

    def run():
        out = layer(x, ei, ea)
        if args.backward:
            (out[0].sum() + out[1].sum()).backward()

    _lib.set_tensor_cores(args.tc)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.iters):
        run()
    t1.record()
    torch.cuda.synchronize()
    layer_ms = t0.elapsed_time(t1) / args.iters
    prof = _lib.profile_collect()
    peak = 6555.8
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    r, q, i, kbar = 4 * nd[0] + 12 * nd[1], 4 * ed[0] + 12 * ed[1], 16, e / n
    alg = {"conv_fwd": q + i + 2 * r / kbar, "conv_bwd": 2 * q + i + 3 * r / kbar}
    flops_edge = {"ck": 5342, "mb": 132820}[args.dims]     # GVPConvLayer forward FLOP per edge at kbar = 30
    res = {"dims": args.dims, "nodes": n, "edges": e, "tensor_cores": bool(args.tc), "backward": bool(args.backward),
           "layer": {"ms": layer_ms, "edges_per_s": e / (layer_ms * 1e-3),
                     "reference_TFLOPs": flops_edge * (3 if args.backward else 1) * e / (layer_ms * 1e-3) / 1e12,
                     "note": "whole GVPConvLayer call incl. weight packing and Python launch overhead, CUDA events"},
           "kernels": {}}
    for name, (ms, cnt) in prof.items():
        if not cnt:
            continue
        avg = ms / cnt
        d = {"avg_ms": avg, "launches": cnt}
        if name in alg:
            d["edges_per_s"] = e / (avg * 1e-3)
            d["hbm_GBps_algorithmic"] = alg[name] * e / (avg * 1e-3) / 1e9
            d["frac_of_measured_hbm_peak"] = d["hbm_GBps_algorithmic"] / peak
        res["kernels"][name] = d
    print(json.dumps(res))


if __name__ == "__main__":
    main()
