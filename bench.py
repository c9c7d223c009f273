#!/usr/bin/env python
"""Headline benchmark: CASTER-DTA(2,2) training step (forward + backward + Adam) on synthetic Davis-shape batches.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the whole model over one batch of 32 protein-ligand pairs (proteins U[300,1000] residues,
kNN k=30 residue graphs incl. self loops, ligands U[20,46] atoms).  `value` = pairs/s with the batch resident in HBM;
`e2e` = the same step fed from pinned HOST buffers through the public module API (H2D of the whole graph batch and a
D2H read of the loss inside the timed region).  Data parallel (N > 1): one process per GPU, per-GPU batch fixed (weak
scaling), one flat-bucket NCCL all-reduce of the 764 396 gradients per step.

`--impl reference` times the CPU port of the reference (`oracle/`, the reference itself is Python that cannot travel
to the GPU box) on the host cores, same batch, same step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PAIRS, KNN = 32, 30
METRIC = "protein-ligand pairs/sec (fwd+bwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="davis")
    ap.add_argument("--pairs", type=int, default=PAIRS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
def host_batch(shape, pairs, seed):
    """Synthetic batch on the host (numpy): protein backbones + node features, ligand graphs, targets."""
    from caster_dta_b200 import synth
    pb = synth.protein_batch_coords(shape, pairs, seed)
    mol = synth.molecule_batch(pairs, seed)
    y = np.random.default_rng(seed + 1).normal(size=(pairs,)).astype(np.float32)
    return pb, mol, y


def oracle_graph(pb, k):
    """Residue graph on the CPU (oracle featurizer) -- reference arm only."""
    from oracle import featurizer_oracle
    eis, ess, evs = [], [], []
    for b in range(len(pb["ptr"]) - 1):
        lo, hi = int(pb["ptr"][b]), int(pb["ptr"][b + 1])
        ei, s, v = featurizer_oracle.residue_graph(pb["coords"][lo:hi], k, "num", True)
        eis.append(ei + lo); ess.append(s); evs.append(v)
    return np.concatenate(eis, 1), np.concatenate(ess), np.concatenate(evs)


def cpu_reference_steps(pb, mol, y, graph, steps, warmup, threads):
    """Forward + backward + Adam of the CPU port on the same batch.  Returns seconds per step."""
    from oracle import gvp_oracle, joint_oracle
    from caster_dta_b200.configs import caster_dta_2_2
    import caster_dta_b200 as cg
    torch.set_num_threads(threads)
    kw = caster_dta_2_2()
    torch.manual_seed(9)
    init = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])   # init only
    p = {k: v.detach().clone().requires_grad_(v.numel() > 0 and v.dtype.is_floating_point) for k, v in init.state_dict().items()}
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4)
    ei, es, ev = graph
    n, e = pb["x_s"].shape[0], ei.shape[1]
    prot = dict(x=(torch.from_numpy(pb["x_s"]), torch.from_numpy(pb["x_v"])), edge_index=torch.from_numpy(ei),
                ntypes=torch.from_numpy(pb["ntypes"]), etypes=torch.zeros(e, dtype=torch.long),
                eattr=(torch.from_numpy(es), torch.from_numpy(ev)), batch=torch.from_numpy(pb["batch"]))
    molt = {k: torch.from_numpy(v) for k, v in mol.items()}
    target = torch.from_numpy(y)
    pk = kw["protein_gnn_kwargs"]
    gen = torch.Generator().manual_seed(1)

    def masks():
        out = []
        for _ in range(pk["num_convs"]):
            pair = []
            for _ in range(2):
                ms = (torch.rand(n, 16, generator=gen) > pk["dropout_rate"]).float() / (1 - pk["dropout_rate"])
                mv = (torch.rand(n, 4, generator=gen) > pk["dropout_rate"]).float() / (1 - pk["dropout_rate"])
                pair.append((ms, mv))
            out.append(tuple(pair))
        return out

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        emb = gvp_oracle.lba_encoder(p, "protein_gnn.gnn_model.", prot["x"], prot["edge_index"], prot["ntypes"],
                                     prot["etypes"], prot["eattr"], pk["num_ntypes"], pk["num_etypes"], pk["num_convs"],
                                     pk["aggr"], drop_masks=masks())
        pred, _ = joint_oracle.joint_forward(p, kw, prot, molt, training=True, protein_embed=emb)
        loss = torch.nn.functional.mse_loss(pred.squeeze(-1), target)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), n, e


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# algorithmic HBM bytes per edge of the fused conv kernels at checkpoint dims:
#   R = 4*ns + 12*nv = 112 B node row, Q = 4*es + 12*ev = 140 B edge row, I = 16 B index pair, kbar = E/N
def conv_bytes_per_edge(kind, kbar, ns=16, nv=4, es=32, ev=1):
    r, q, i = 4 * ns + 12 * nv, 4 * es + 12 * ev, 16
    if kind == "conv_fwd":
        return q + i + 2 * r / kbar               # read edge row + index, read x and write dh once per node
    return 2 * q + i + 3 * r / kbar               # + write d(edge row); read x, d_out and write d_x once per node


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (e.g. "NCCL version ..." from the first collective) must not share stdout with the ONE JSON line:
    route fd 1 to stderr for the whole run and keep the real stdout for `emit`."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    args = parse()
    quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        pb, mol, y = host_batch(args.shape, args.pairs, 9)
        graph = oracle_graph(pb, KNN)
        sec, n, e = cpu_reference_steps(pb, mol, y, graph, args.steps, max(args.warmup, 1), threads)
        val = args.pairs / sec
        sample = f"{args.steps} full steps (fwd+bwd+Adam) on one {args.pairs}-pair {args.shape}-shape batch, N={n}, E={e}"
        emit(({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CASTER-DTA(2,2) train step, {args.shape}-shape, {args.pairs} pairs, kNN k={KNN}",
                       "nodes": n, "edges": e, "note": "CPU port of the reference (oracle/), eager fp32, all host threads"},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch.distributed as dist
    import caster_dta_b200 as cg
    from caster_dta_b200 import _lib, ops, parallel
    from caster_dta_b200.configs import caster_dta_2_2

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    # ---- batch (each rank its own, same shape distribution) ----------------------------------------------------------
    pb, mol, y = host_batch(args.shape, args.pairs, 9 + rank)
    coords = torch.from_numpy(pb["coords"]).to(dev)
    ptr = torch.from_numpy(pb["ptr"]).to(dev)
    ei, (e_s, e_v), etypes = cg.residue_graph_batch(coords, ptr, KNN, "num", True)
    n, e = int(pb["x_s"].shape[0]), int(ei.shape[1])
    max_res = int(np.diff(pb["ptr"]).max())
    max_atoms = int(np.bincount(mol["batch"]).max())
    host = {
        "p_x_s": torch.from_numpy(pb["x_s"]), "p_x_v": torch.from_numpy(pb["x_v"]), "p_ei": ei.cpu(), "p_nt": torch.from_numpy(pb["ntypes"]),
        "p_et": etypes.cpu(), "p_e_s": e_s.cpu(), "p_e_v": e_v.cpu(), "p_batch": torch.from_numpy(pb["batch"]),
        "m_x": torch.from_numpy(mol["x"]), "m_ei": torch.from_numpy(mol["edge_index"]), "m_nt": torch.from_numpy(mol["ntypes"]),
        "m_et": torch.from_numpy(mol["etypes"]), "m_ea": torch.from_numpy(mol["eattr"]), "m_batch": torch.from_numpy(mol["batch"]),
        "y": torch.from_numpy(y),
    }
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def to_device(buf=None):
        if buf is None:
            return {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        for k, v in host.items():
            buf[k].copy_(v, non_blocking=True)
        return buf

    def dicts(d):
        prot = dict(x=(d["p_x_s"], d["p_x_v"]), edge_index=d["p_ei"], ntypes=d["p_nt"], etypes=d["p_et"],
                    eattr=(d["p_e_s"], d["p_e_v"]), batch=d["p_batch"], num_graphs=args.pairs, max_nodes=max_res)
        molg = dict(x=d["m_x"], edge_index=d["m_ei"], ntypes=d["m_nt"], etypes=d["m_et"], eattr=d["m_ea"], batch=d["m_batch"],
                    num_graphs=args.pairs, max_nodes=max_atoms)
        return prot, molg

    # ---- model ---------------------------------------------------------------------------------------------------------
    kw = caster_dta_2_2()
    torch.manual_seed(9)
    model = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"]).to(dev).train()
    model.overlap_encoders = os.environ.get("CGVP_OVERLAP", "1") == "1"
    if model.overlap_encoders and os.environ.get("CGVP_WGRAD_STREAM", "1") == "1":
        ops.set_wgrad_stream(torch.cuda.Stream(device=dev))
    parallel.broadcast_parameters(model, 0)
    bucket = parallel.GradSync(model)
    opt = torch.optim.Adam(bucket.params, lr=1e-4, fused=True, capturable=True)
    n_params = bucket.numel

    def fwd_bwd(d):
        prot, molg = dicts(d)
        pred, _ = model(prot, molg)
        loss = torch.nn.functional.mse_loss(pred.squeeze(-1), d["y"])
        loss.backward()
        ops.join_wgrad_stream()
        return loss.detach()

    graphs = {}          # id(batch dict) -> GraphedStep replaying zero-grad + forward + backward on that batch's buffers
    graph_note = "eager launches (--no-graph)" if args.no_graph else None

    def step(d, eager=False):
        g = None if eager else graphs.get(id(d))
        if g is None:
            bucket.reset()                                  # autograd then assigns the gradients (no accumulate kernels)
            loss = fwd_bwd(d)
        else:
            g.select()
            loss = g.replay()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def capture(d):
        """Capture forward + backward on the (static) buffers of `d`; the all-reduce and Adam stay eager."""
        nonlocal graph_note
        if args.no_graph or graph_note not in (None, "cuda graph"):
            return
        from caster_dta_b200.graphs import GraphedStep
        try:
            graphs[id(d)] = GraphedStep(lambda: fwd_bwd(d), bucket)
            graph_note = "cuda graph"
        except Exception as exc:                       # keep the bench alive; the JSON line says what happened
            import traceback
            traceback.print_exc(file=sys.stderr)
            graphs.clear()
            graph_note = f"eager launches (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            torch.cuda.synchronize()

    resident = to_device()
    torch.cuda.synchronize()
    for _ in range(2):
        step(resident)                                  # eager warm-up (allocator, plan cache, cuBLAS handles)
    launches_per_step0 = _lib.LAUNCHES
    step(resident)
    launches_per_step = _lib.LAUNCHES - launches_per_step0
    torch.cuda.synchronize()
    capture(resident)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also finds the dominant kernel of the step) -----------------------------------------------------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(2):
        step(resident, eager=True)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    dominant = max(("conv_fwd", "conv_bwd", "rows_fwd", "rows_bwd", "segment_reduce"), key=lambda k: prof[k][0])
    _lib.profile_enable(False)

    if os.environ.get("CGVP_BENCH_TRACE") and rank == 0:      # diagnostics only: where does a step spend CPU / GPU time
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as tp:
            for _ in range(3):
                step(resident)
            torch.cuda.synchronize()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "step_trace.txt"), "w") as fh:
            fh.write(tp.key_averages().table(sort_by="self_cpu_time_total", row_limit=45))
            fh.write("\n\n")
            fh.write(tp.key_averages().table(sort_by="cuda_time_total", row_limit=45))
        t0 = time.perf_counter()
        for _ in range(5):
            step(resident)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"[trace] 5 steps: cpu issue {1e3 * (t1 - t0) / 5:.2f} ms/step, incl. drain {1e3 * (t2 - t0) / 5:.2f} ms/step", file=sys.stderr)

    # ---- timed region 1: device-resident inputs ------------------------------------------------------------------------
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    clocks = ClockSampler(local)
    launches0 = _lib.LAUNCHES
    barrier()
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)                        # L2 flush between timed iterations (outside the events)
        a.record()
        step(resident)
        b.record()
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = launches_per_step * args.steps
    # per-kernel device time of the SAME step: the library brackets each main kernel with CUDA events on the launching
    # stream; that needs host calls, so this pass launches eagerly (a graph replay makes none).  L2 flushed as above.
    _lib.profile_enable(True)
    for _ in range(args.steps):
        flush.fill_(1)
        step(resident, eager=True)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = args.pairs * world * args.steps / (ms_total / 1e3)

    # ---- timed region 2: end to end from pinned host buffers, double-buffered H2D on a copy stream ------------------------
    copy_stream = torch.cuda.Stream(dev)
    bufs = [to_device(), to_device()]
    torch.cuda.synchronize()
    for b in bufs:
        capture(b)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[i])          # the step that last used this buffer has finished
            to_device(bufs[i])
            ready[i].record(copy_stream)

    for i in range(2):
        done[i].record()
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    prefetch(0)
    loss_host = 0.0
    # every step's loss is copied to pinned host memory and read by the host; the read of step i happens while step i+1 runs
    # (a training loop that logs its loss one step late), so the host never stalls the device between steps
    loss_pin = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    for it in range(e2e_steps):
        cur = it & 1
        if it + 1 < e2e_steps:
            prefetch(cur ^ 1)
        torch.cuda.current_stream().wait_event(ready[cur])
        loss = step(bufs[cur])
        done[cur].record()
        loss_pin[cur].copy_(loss, non_blocking=True)  # D2H read of the step's result ...
        loss_ev[cur].record()
        if it > 0:
            loss_ev[cur ^ 1].synchronize()            # ... consumed one step later
            loss_host = float(loss_pin[cur ^ 1])
    loss_ev[(e2e_steps - 1) & 1].synchronize()
    loss_host = float(loss_pin[(e2e_steps - 1) & 1])
    barrier()
    e2e_sec = time.perf_counter() - t0
    clock_info = clocks.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and e2e)
    t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.pairs * world * e2e_steps / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    kms, kn = prof[dominant]
    kbar = e / n
    if dominant in ("conv_fwd", "conv_bwd"):
        alg_bytes = conv_bytes_per_edge(dominant, kbar) * e
    else:
        alg_bytes = None
    roof = {"bound": "hbm", "kernel": dominant + "_kernel", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "launches": kn, "avg_ms": kms / max(kn, 1), "share_of_step": kms / max(ms_total, 1e-9),
            "kernel_timing": "CUDA events around each launch of the kernel in an eager pass of the same step (L2 flushed between steps)",
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]}}
    if alg_bytes is not None and kn:
        roof["achieved"] = alg_bytes / (kms / kn * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = alg_bytes
        # context: at checkpoint dims this kernel is fp32-FMA bound (31 FLOP/B > the 11 FLOP/B CUDA-core ridge), so also report
        # the arithmetic side -- ALGORITHMIC FLOPs (5 086 / edge forward; the backward is 2x that, SURVEY 8d -- the forward
        # recompute the kernel also does is not counted) against the FFMA peak
        flops = 5086.0 * e * (2.0 if dominant == "conv_bwd" else 1.0)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_peak = sm_count * 128 * 2 * (clock_info or {}).get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        roof["fp32"] = {"achieved_tflops": flops / (kms / kn * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                        "frac": flops / (kms / kn * 1e-3) / 1e12 / fp32_peak,
                        "note": "algorithmic FLOPs (recompute not counted) against FFMA peak = SMs x 128 lanes x 2 x max SM clock; the kernel is arithmetic / issue bound, which is why its HBM fraction is small"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(traffic_file):
        roof["traffic"] = json.load(open(traffic_file)).get(dominant)
        if dominant == "conv_bwd":
            roof["traffic_note"] = ("DRAM bytes include reading the 224 B/edge training stash written by the forward (a deliberate "
                                    "recompute-for-traffic trade, cgvp_conv_fwd_stash) and the dj round trip; algorithmic bytes are the "
                                    "compulsory traffic without it")

    # ---- CPU baseline (bounded sample) --------------------------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        graph = (host["p_ei"].numpy(), host["p_e_s"].numpy(), host["p_e_v"].numpy())
        sec, _, _ = cpu_reference_steps(pb, mol, y, graph, 2, 1, threads)
        cpu = {"value": args.pairs / sec, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": f"2 full steps (fwd+bwd+Adam) on the same {args.pairs}-pair batch after 1 warm-up; {sec:.2f} s/step"}

    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": warm + 2,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"CASTER-DTA(2,2) train step (fwd+bwd+Adam), {args.shape}-shape, {args.pairs} pairs/GPU, kNN k={KNN} + self loops",
                   "global_batch": args.pairs * world, "nodes_per_gpu": n, "edges_per_gpu": e, "params": n_params,
                   "parallelism": f"dp{world}", "l2": "flushed between timed iterations (256 MB write)",
                   "launch_mode": graph_note or "eager launches",
                   "edges_per_s_conv": e * 2 * args.steps / max(ms_total / 1e3, 1e-9)},
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "note": "whole graph batch copied from pinned host memory each step (double-buffered); every step's loss is copied to pinned host memory and read by the host one step later",
                "last_loss": loss_host},
        "gpu_launches": launches,
        "gpu_launches_note": "C-ABI calls into libcastergvp.so per step x steps (each enqueues 1-6 kernels; replayed from a CUDA graph when launch_mode says so)",
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
