"""Probe: can an NCCL all-reduce be captured in a CUDA graph and replayed on this box (torchrun, one rank per GPU)?
Prints a marker per stage to stderr; run under `timeout`."""
import os
import sys
import time

import torch
import torch.distributed as dist


def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time():.1f}", *a, file=sys.stderr, flush=True)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    log("init done")
    x = torch.full((800000,), float(dist.get_rank() + 1), device=dev)
    dist.all_reduce(x)
    torch.cuda.synchronize()
    log("eager all_reduce ok", float(x[0]))
    y = torch.zeros(800000, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y.add_(1.0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    log("capturing")
    with torch.cuda.graph(g):
        y.mul_(2.0)
        dist.all_reduce(y)
        y.add_(1.0)
    log("captured")
    for i in range(3):
        y.fill_(float(dist.get_rank() + 1))
        g.replay()
        torch.cuda.synchronize()
        log("replay", i, float(y[0]))
    dist.barrier()
    log("done")
    if os.environ.get("PROBE_KEEP_GRAPH") != "1":
        del g                       # a live graph that captured collectives keeps the communicator busy: destroy hangs
        torch.cuda.synchronize()
    log("destroying")
    dist.destroy_process_group()
    log("destroyed")


if __name__ == "__main__":
    main()
