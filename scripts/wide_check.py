#!/usr/bin/env python
"""One-call GPU check of the wide-dims GEMM formulation (caster_dta_b200/wide.py): the parity tests of tests/test_gpu_wide.py,
then the config-5 measurement block of bench.py.  Results are appended to gpurun_out/wide_check.jsonl as they arrive, so a
call that is cut off still leaves what it finished.

    python scripts/wide_check.py [--budget SECONDS] [--edges N]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = time.time()
OUT = os.path.join(ROOT, "gpurun_out", "wide_check.jsonl")


def put(obj):
    obj["t"] = round(time.time() - T0, 2)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as fh:
        fh.write(json.dumps(obj) + "\n")
        fh.flush()
        os.fsync(fh.fileno())
    print(json.dumps(obj), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=35.0, help="stop starting new stages after this many seconds")
    ap.add_argument("--edges", type=int, default=1_000_000)
    args = ap.parse_args()
    import torch
    put({"stage": "import", "cuda": torch.cuda.is_available()})
    if not torch.cuda.is_available():
        return 1
    import pytest
    rc = pytest.main(["-x", "-q", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_wide.py"), "-p", "no:cacheprovider", "--tb=short"])
    put({"stage": "pytest tests/test_gpu_wide.py", "rc": int(rc)})
    if time.time() - T0 > args.budget:
        put({"stage": "stopped", "why": "time budget"})
        return int(rc)
    import bench
    res = bench.config5_block(torch.device("cuda", 0), edges=args.edges, iters=3)
    put({"stage": "config5", "result": res})
    if time.time() - T0 > args.budget:
        put({"stage": "stopped", "why": "time budget"})
        return int(rc)
    try:                                    # the checkpoint-dims path through the same autograd Functions
        import __graft_entry__ as g
        g.smoke()
        put({"stage": "smoke", "ok": True})
    except Exception as exc:                # noqa: BLE001
        put({"stage": "smoke", "ok": False, "error": f"{type(exc).__name__}: {str(exc)[:300]}"})
        return 1
    return int(rc)


if __name__ == "__main__":
    sys.exit(main())
