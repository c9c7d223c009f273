#!/usr/bin/env python
"""Headline benchmark: CASTER-DTA(2,2) training step (forward + backward + Adam) on synthetic Davis-shape batches.

    python bench.py --gpus N --steps K --warmup W [--impl reference | reference-gpu]

One "step" = one optimizer step on one mini-batch drawn by the size-capped batch sampler (`PMD_BatchSampler` rule,
`dataset/dual_dataset.py:424-522`): 32 protein-ligand pairs per GPU, proteins U[300,1000] residues as kNN k=30 residue
graphs incl. self loops, ligands U[20,46] atoms.  The steps CYCLE over a pool of distinct batches (different residue /
edge / atom counts), never the same batch twice in a row.

Our arm, per step, inside the timed region: copy of the batch into the step's input buffers, residue-graph featurizer
from backbone coordinates (node features, kNN edges, RBF / positional / direction features), graph-plan build
(`cgvp_plan_build`), JointGNN forward, MSE loss, backward, gradient pack, one NCCL all-reduce (N > 1) and fused Adam.
Launch mode: one CUDA graph per padded-shape bucket (`caster_dta_b200/training.py`), or `--no-graph` for eager launches.
`value` = pairs/s with the pool resident in HBM; `e2e` = the same steps fed from pinned HOST buffers (H2D of each step's
batch and D2H of each step's loss inside the timed region).

`--impl reference` times the CPU port of the reference (`oracle/`; the reference itself is Python with PyG dependencies
and cannot travel to the GPU box) on the host cores over the same pool, featurized graphs precomputed as the reference's
dataset does.  `--impl reference-gpu` runs the same port in eager fp32 on cuda:0 (`allow_tf32=False`): the stock
PyTorch `index_select` / `index_add_` GPU path the reference would take (`train_model.py:561-570`, without autocast).
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
SHAPE_CAPS = {"davis": 16_000_000, "kiba": 8_000_000, "bindingdb": 4_000_000, "tiny": 16_000_000}      # train_model.py:240-248
SHAPE_MAXLEN = {"davis": 1000, "kiba": 2000, "bindingdb": 2000, "tiny": 60}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu", "config5"])
    ap.add_argument("--shape", default="davis")
    ap.add_argument("--pairs", type=int, default=PAIRS, help="pairs per GPU and step (the sampler's max batch size)")
    ap.add_argument("--pool", type=int, default=8, help="distinct batches the steps cycle over")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-config5", action="store_true", help="skip the isolated GVPConvLayer (100,16)/(32,1) measurement")
    ap.add_argument("--no-kiba", action="store_true", help="skip the KIBA-shape (BASELINE config 3 shape, 64 pairs/GPU) sub-run")
    ap.add_argument("--max-seconds", type=float, default=900.0, help="abort a run that takes longer than this (watchdog)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
def build_pool(shape, pairs, pool, rank, world, pin):
    """`pool` fixed-shape host batches for this rank from the size-capped loader (same global batches on every rank)."""
    from caster_dta_b200 import loader
    ds = loader.SyntheticPairDataset(shape, pairs * world * pool, seed=9, edge_thresh=KNN, thresh_type="num")
    ld = loader.PairBatchLoader(ds, max_num=SHAPE_CAPS[shape], max_bsize=pairs, shuffle=True, seed=9, rank=rank, world_size=world,
                                pin=pin)
    out = []
    for t, m in ld:
        out.append((t, m))
        if len(out) == pool:
            break
    return ds, out


def workload_config(args, world):
    """The SAME dict in every arm (the driver compares it)."""
    return {"workload": f"CASTER-DTA(2,2) train step (fwd+bwd+Adam), {args.shape}-shape, {args.pairs} pairs/GPU, kNN k={KNN} + self loops, "
                        f"steps cycle over a pool of {args.pool} distinct batches from the size-capped sampler",
            "global_batch": args.pairs * world, "parallelism": f"dp{world}", "pool": args.pool,
            "l2": "flushed between timed iterations (256 MB write)"}


def real_part(t, m):
    """The real pairs of a padded batch (numpy): what the reference's loader would hand to its model."""
    n, a, me, p = m["nodes"], m["atoms"], m["mol_edges"], m["pairs"]
    return dict(coords=t["coords"][:n].numpy(), ptr=t["ptr"][:p + 1].numpy(), idents=t["idents"][:n].numpy(),
                m_x=t["m_x"][:a], m_ei=t["m_ei"][:, :me], m_ea=t["m_ea"][:me], m_nt=t["m_nt"][:a], m_et=t["m_et"][:me],
                m_batch=t["m_batch"][:a], y=t["y"][:p], w=torch.full((p,), 1.0 / p))


def reference_steps(pool, aa_table, steps, warmup, threads, device="cpu", featurized=None):
    """Forward + backward + Adam of the reference port (oracle/) over the pool, eager fp32.  The featurized graphs are
    built once, outside the timing (the reference's dataset stores them).  Returns (seconds per step, pairs per step)."""
    from oracle import pipeline
    from caster_dta_b200.configs import caster_dta_2_2
    import caster_dta_b200 as cg
    torch.set_num_threads(threads)
    kw = caster_dta_2_2()
    torch.manual_seed(9)
    init = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])   # init only
    dev = torch.device(device)
    p = {k: v.detach().clone().to(dev).requires_grad_(v.numel() > 0 and v.dtype.is_floating_point) for k, v in init.state_dict().items()}
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4)
    pk = kw["protein_gnn_kwargs"]
    batches = []
    for i, (t, m) in enumerate(pool):
        r = real_part(t, m)
        prot = featurized[i] if featurized is not None else pipeline.featurize_batch(r["coords"], r["ptr"], r["idents"], aa_table, KNN, "num", True)
        mol = pipeline.molecule_dict(r)
        mv = lambda d: {k: (tuple(x.to(dev) for x in v) if isinstance(v, tuple) else v.to(dev)) for k, v in d.items()}
        batches.append((mv(prot), mv(mol), r["y"].to(dev), r["w"].to(dev), m["nodes"], m["pairs"]))
    gen = torch.Generator(device=dev).manual_seed(1)
    keep = 1 - pk["dropout_rate"]

    def masks(n):
        out = []
        for _ in range(pk["num_convs"]):
            pair = []
            for _ in range(2):
                ms = (torch.rand(n, 16, generator=gen, device=dev) < keep).float() / keep
                mvv = (torch.rand(n, 4, generator=gen, device=dev) < keep).float() / keep
                pair.append((ms, mvv))
            out.append(tuple(pair))
        return out

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    times, pairs = [], 0
    for it in range(warmup + steps):
        prot, mol, y, w, n, npairs = batches[it % len(batches)]
        sync()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = pipeline.train_loss(p, kw, prot, mol, y, w, masks(n), None, training=True)
        loss.backward()
        opt.step()
        sync()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            pairs += npairs
    return float(np.sum(times)) / max(len(times), 1), pairs / max(len(times), 1), batches


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


def reference_arm(args, world):
    gpu = args.impl == "reference-gpu"
    if gpu and not torch.cuda.is_available():
        emit({"impl": "reference-gpu", "unavailable": "no CUDA device"})
        return
    threads = os.cpu_count() or 1
    ds, pool = build_pool(args.shape, args.pairs, args.pool, 0, 1, pin=False)
    if gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    use = pool if gpu else pool[:min(len(pool), max(2, min(args.steps + args.warmup, 4)))]      # CPU: featurizing a batch costs ~4 s
    sec, pairs, _ = reference_steps(use, ds.aa_table, args.steps, args.warmup, threads, "cuda" if gpu else "cpu")
    val = pairs / sec
    sizes = [(m["nodes"], KNN * m["nodes"]) for _, m in use]
    sample = (f"{args.steps} full steps (fwd+bwd+Adam) after {args.warmup} warm-up, cycling over {len(use)} of the pool's batches "
              f"(N, E) = {sizes}; {sec * 1e3:.1f} ms/step")
    emit({
        "impl": args.impl, "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "note": ("port of the reference (oracle/, a restatement pinned to the unmodified reference at 1e-12 by tests/golden), eager fp32, "
                 + ("on cuda:0 with allow_tf32=False -- the stock PyTorch GPU path" if gpu else f"all {threads} host threads")
                 + "; one rank's batch per step"),
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def config5_block(dev, edges=1_000_000, iters=5):
    """BASELINE config 5 beside the headline (rank 0, N = 1): one isolated `GVPConvLayer` at nodes (100,16) / edges (32,1),
    `vector_gate=True, activations=(ReLU, None), aggr='mean', drop_rate=0` on the synthetic chain-kNN graph of SURVEY.md 8(d),
    E = 1 M: forward and forward+backward per precision mode, CUDA events, and a parity line (GEMM formulation vs the generic
    tile kernels of the same library at E = 60 k).  Never raises: the headline line must not depend on it."""
    import torch.nn.functional as F
    import caster_dta_b200 as cg
    from caster_dta_b200 import _lib, synth, wide
    res = {"dims": "nodes (100,16), edges (32,1)", "flops_per_edge_fwd": 132820}
    prev_wide = wide.ENABLED
    try:
        nd, ed = (100, 16), (32, 1)

        def make(num_edges):
            ei_np, n = synth.conv_microbench_graph(num_edges, 30)
            ei = torch.from_numpy(ei_np).to(dev)
            g = torch.Generator(device=dev).manual_seed(9)
            r = lambda *sh: torch.randn(*sh, generator=g, device=dev)
            e = ei.shape[1]
            return ei, n, e, [r(n, nd[0]), r(n, nd[1], 3), r(e, ed[0]), r(e, ed[1], 3)]

        torch.manual_seed(9)
        layer = cg.GVPConvLayer(nd, ed, drop_rate=0.0, activations=(F.relu, None), vector_gate=True, aggr="mean").to(dev).train()

        def run(ei, data, backward):
            leaves = [t.detach().requires_grad_(backward) for t in data]
            layer.zero_grad(set_to_none=True)
            with torch.set_grad_enabled(backward):
                out = layer((leaves[0], leaves[1]), ei, (leaves[2], leaves[3]))
                if backward:
                    (out[0].square().sum() + out[1].square().sum()).backward()
            return out, leaves

        def timed(ei, data, backward, n_it):
            for _ in range(2):
                run(ei, data, backward)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(n_it):
                run(ei, data, backward)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n_it

        ei, n, e, data = make(edges)
        res["nodes"], res["edges"] = n, e
        modes = {}
        for name, gemm, tc, bwd in (("fwd_tcgen05_bf16", True, True, False), ("fwd_fp32_gemm", True, False, False),
                                    ("fwd_bwd_fp32_gemm", True, False, True), ("fwd_bwd_tcgen05_fwd_tf32_bwd", True, True, True)):
            wide.set_enabled(gemm)
            _lib.set_tensor_cores(tc)
            try:
                ms = timed(ei, data, bwd, iters)
                modes[name] = {"ms": ms, "edges_per_s": e / (ms * 1e-3),
                               "algorithmic_tflops": 132820.0 * (3 if bwd else 1) * e / (ms * 1e-3) / 1e12}
            except Exception as exc:                                   # noqa: BLE001
                modes[name] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
            finally:
                _lib.set_tensor_cores(False)
        res["modes"] = modes
        # the bar on the same box: the reference port (oracle/, eager fp32 torch with autograd -- the stock index_select / cat /
        # cuBLAS / index_add_ path the reference's modules take on a GPU) on the same layer, graph and inputs
        try:
            from oracle import gvp_oracle
            pref = {k: v.detach().clone().requires_grad_(v.numel() > 0) for k, v in layer.state_dict().items()}

            def ref_run(backward):
                leaves = [t.detach().clone().requires_grad_(backward) for t in data]
                for v in pref.values():
                    v.grad = None
                with torch.set_grad_enabled(backward):
                    out = gvp_oracle.gvp_conv_layer(pref, "", (leaves[0], leaves[1]), ei, (leaves[2], leaves[3]), aggr="mean",
                                                    scalar_act="relu", vector_act=None, vector_gate=True)
                    if backward:
                        (out[0].square().sum() + out[1].square().sum()).backward()

            refres = {}
            for name, bwd in (("fwd", False), ("fwd_bwd", True)):
                torch.cuda.reset_peak_memory_stats(dev)
                ref_run(bwd)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(3):
                    ref_run(bwd)
                b.record()
                torch.cuda.synchronize()
                refres[name + "_ms"] = a.elapsed_time(b) / 3
                refres[name + "_peak_memory_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
            res["reference_port_eager_gpu"] = refres
            del pref
        except Exception as exc:                                       # noqa: BLE001
            res["reference_port_eager_gpu"] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
        torch.cuda.empty_cache()
        # the C-ABI segmented reductions inside this training path (aggregation of the [E,148] message rows; R_i / R_j of the
        # [E,200] backward rows over the target / source CSR views): device time from the library's own event bracketing,
        # algorithmic bytes = rows read once + node rows written once + the CSR offsets (+ the index array on the source side)
        try:
            wide.set_enabled(True)
            run(ei, data, True)
            torch.cuda.synchronize()
            _lib.profile_enable(True)
            run(ei, data, True)
            torch.cuda.synchronize()
            ms, cnt = _lib.profile_collect()["segment_reduce"]
            _lib.profile_enable(False)
            wf, wb = 148, 200
            alg = 4.0 * (e * wf + n * wf + n) + 2 * 4.0 * (e * wb + n * wb + n) + 4.0 * e
            peak, src = measured_peaks()
            res["segment_reduce_in_training_path"] = {"launches": cnt, "ms_total": ms, "algorithmic_bytes": alg,
                                                      "achieved_GBps": alg / (ms * 1e-3) / 1e9 if ms > 0 else None,
                                                      "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
                                                      "peak_GBps": peak, "peak_source": src}
        except Exception as exc:                                       # noqa: BLE001
            res["segment_reduce_in_training_path"] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
            _lib.profile_enable(False)
        # parity at the benchmarked size: tcgen05 bf16 forward against the fp32 GEMM forward of the same layer (<= 1e-2 mode)
        try:
            outs = {}
            for name, tc in (("fp32", False), ("tc", True)):
                wide.set_enabled(True)
                _lib.set_tensor_cores(tc)
                try:
                    outs[name] = run(ei, data, False)[0]
                finally:
                    _lib.set_tensor_cores(False)
            cmp = {}
            for i, k in enumerate(("s", "V")):
                a, b = outs["tc"][i].double(), outs["fp32"][i].double()
                cmp[k] = {"max_scale_relative": float((a - b).abs().max() / b.abs().max()),
                          "l2_relative": float((a - b).norm() / b.norm())}
            res["tcgen05_bf16_forward_vs_fp32_forward"] = cmp
            del outs
        except Exception as exc:                                       # noqa: BLE001
            res["tcgen05_bf16_forward_vs_fp32_forward"] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
        del ei, data
        # one larger point of the training path (chunking keeps the intermediates bounded: the time should scale with E)
        try:
            ei, n, e, data = make(10 * edges)
            wide.set_enabled(True)
            torch.cuda.reset_peak_memory_stats(dev)
            ms = timed(ei, data, True, 2)
            res["fwd_bwd_fp32_gemm_10x_edges"] = {"edges": e, "ms": ms, "edges_per_s": e / (ms * 1e-3),
                                                  "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
        except Exception as exc:                                       # noqa: BLE001
            res["fwd_bwd_fp32_gemm_10x_edges"] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
        ei = data = None
        torch.cuda.empty_cache()
        # the generic tile kernels (the only training path of these dims before wide.py), on a graph they finish quickly
        ei, n, e, data = make(60_000)
        small = {"edges": e}
        grads = {}
        for name, gemm in (("gemm", True), ("tile", False)):
            wide.set_enabled(gemm)
            small[name + "_fwd_bwd_ms"] = timed(ei, data, True, 3)
            out, leaves = run(ei, data, True)
            grads[name] = [out[0].detach(), out[1].detach()] + [t.grad for t in leaves] + [q.grad.clone() for q in layer.parameters() if q.numel()]
        # outputs: max scale-relative difference; gradients: L2-relative per tensor and the share of per-row gradient rows with
        # an entry beyond 1e-4 of the tensor's scale (a ReLU pre-activation within round-off of zero takes the other branch in
        # one of the two arithmetic orders and switches that row's term: tests/helpers.py::assert_rows_close)
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))
        l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))
        small["outputs_max_scale_relative"] = max(rel(a, b) for a, b in zip(grads["gemm"][:2], grads["tile"][:2]))
        small["gradients_l2_relative_max"] = max(l2(a, b) for a, b in zip(grads["gemm"][2:], grads["tile"][2:]))
        bad = 0.0
        for a, b in zip(grads["gemm"][2:6], grads["tile"][2:6]):
            d = (a - b).abs().reshape(a.shape[0], -1).amax(1)
            bad = max(bad, float((d > 1e-4 * b.abs().max()).double().mean()))
        small["gradient_rows_beyond_1e-4_share_max"] = bad
        small["tile_ms_per_1M_edges_extrapolated"] = small["tile_fwd_bwd_ms"] * 1e6 / e
        res["tile_kernels_reference"] = small
    except Exception as exc:                                           # noqa: BLE001
        res["error"] = f"{type(exc).__name__}: {str(exc)[:200]}"
    finally:
        wide.set_enabled(prev_wide)
        _lib.set_tensor_cores(False)
    return res


def run_child(extra, timeout):
    """Run this script again in a CHILD process and return its ONE JSON line (or an {"error": ...} dict): the extras measured
    beside the headline must not be able to crash, hang or poison the CUDA context of the process that prints the line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1"] + list(extra)
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"child exited {r.returncode}", "stderr_tail": r.stderr[-300:]}
        return json.loads(lines[-1])
    except Exception as exc:                                           # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}


def kiba_shape_subrun(args):
    """BASELINE config 3's shape beside the headline (rank 0, N = 1): the same benchmark in a child process with
    `--shape kiba --pairs 64` (proteins up to 2 000 residues, N ~ 46 k, E ~ 1.4 M per step).  Never raises."""
    d = run_child(["--shape", "kiba", "--pairs", "64", "--steps", str(max(3, min(args.steps, 20))),
                   "--warmup", str(max(3, min(args.warmup, 5))), "--pool", str(args.pool), "--no-cpu-baseline",
                   "--no-reference-gpu", "--no-config5", "--no-kiba", "--max-seconds", "240"], timeout=300)
    if "error" in d:
        return d
    roof, e2e = d.get("roofline") or {}, d.get("e2e") or {}
    shapes = ((d.get("batches") or {}).get("per_rank_real_nodes_edges_atoms_padded_nodes") or [[]])[0]
    return {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "warmup": d["warmup"],
            "e2e": {k: e2e.get(k) for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "config": d["config"], "launch_mode": d.get("launch_mode"), "graphs_captured": d.get("graphs_captured"),
            "real_nodes_per_batch": [sh[0] for sh in shapes], "real_edges_per_batch": [sh[1] for sh in shapes],
            "clocks": d.get("clocks"), "kernel_ms_per_step": roof.get("kernel_ms_per_step"),
            "dominant_kernel": roof.get("kernel"), "roofline_frac": roof.get("frac")}


_T0 = time.time()


def mark(rank, what):
    """Progress marker on stderr (one line per stage and rank): a stuck run shows where it stopped."""
    print(f"[bench rank {rank} +{time.time() - _T0:6.1f}s] {what}", file=sys.stderr, flush=True)


def arm_watchdog(rank, seconds):
    """A run that exceeds `seconds` (a hung collective, a stuck box) must not wait for the driver's kill: say so and leave."""
    def fire():
        print(f"[bench rank {rank}] watchdog: still running after {seconds} s -- aborting", file=sys.stderr, flush=True)
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def main():
    args = parse()
    quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    arm_watchdog(rank, args.max_seconds)

    if args.impl == "config5":                       # child mode of the default run (see run_child): the config-5 block alone
        if rank == 0:
            if not torch.cuda.is_available():
                emit({"error": "no CUDA device"})
                return
            torch.cuda.set_device(local)
            torch.backends.cuda.matmul.allow_tf32 = False
            emit(config5_block(torch.device("cuda", local)))
        return
    if args.impl != "ours":
        if rank == 0:
            reference_arm(args, world)
        return

    import torch.distributed as dist
    import caster_dta_b200 as cg
    from caster_dta_b200 import _lib, ops, parallel, training
    from caster_dta_b200.configs import caster_dta_2_2

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    # ---- data: this rank's shard of each global batch, padded to its bucket, pinned on the host and resident on the device --
    mark(rank, "building the batch pool")
    ds, pool = build_pool(args.shape, args.pairs, args.pool, rank, world, pin=True)
    mark(rank, f"pool ready: {[(m['nodes'], m['n_pad']) for _, m in pool]}")
    resident = [({k: v.to(dev) for k, v in t.items()}, m) for t, m in pool]
    h2d_bytes = [sum(v.numel() * v.element_size() for v in t.values()) for t, _ in pool]
    torch.cuda.synchronize()

    # ---- model ---------------------------------------------------------------------------------------------------------
    kw = caster_dta_2_2()
    torch.manual_seed(9)
    model = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"]).to(dev).train()
    model.overlap_encoders = os.environ.get("CGVP_OVERLAP", "1") == "1"
    if model.overlap_encoders and os.environ.get("CGVP_WGRAD_STREAM", "1") == "1":
        ops.set_wgrad_stream(torch.cuda.Stream(device=dev))
    parallel.broadcast_parameters(model, 0)
    opt = parallel.FlatAdam(model, lr=1e-4)
    if world > 1:                                     # connect the communicator before anything is captured
        dist.all_reduce(opt.flat_grad)
        opt.flat_grad.zero_()
    max_len = max(SHAPE_MAXLEN[args.shape], 1024 + 32)             # longest real protein / largest dummy (bucket granularity)
    mk = dict(aa_table=torch.from_numpy(ds.aa_table), edge_thresh=KNN, thresh_type="num", keep_self_loops=True, max_len=max_len,
              max_atoms=128 + 2)
    graph_note = "eager launches (--no-graph)" if args.no_graph else "cuda graph per padded-shape bucket"
    if world > 1 and not args.no_graph and os.environ.get("CGVP_ALLREDUCE_IN_GRAPH", "0") != "1":
        graph_note += " (featurizer .. backward .. gradient pack); NCCL all-reduce + fused Adam launched after each replay"
    stepper = training.BucketedTrainStep(model, opt, launch_mode="eager" if args.no_graph else "graph", **mk)
    eager = training.BucketedTrainStep(model, opt, launch_mode="eager", **mk)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): one eager step per bucket-distinct batch, then capture every bucket of the pool --------------------
    mark(rank, "model ready; first eager step")
    l0 = _lib.LAUNCHES
    eager.step(*resident[0])
    launches_per_step = _lib.LAUNCHES - l0
    torch.cuda.synchronize()
    mark(rank, "capturing the pool's buckets")
    if not args.no_graph:
        try:
            for b, m in resident:
                stepper.prepare(b, m)
        except Exception as exc:                      # keep the bench alive; the JSON line says what happened
            import traceback
            traceback.print_exc(file=sys.stderr)
            graph_note = f"eager launches (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            stepper = eager
            torch.cuda.synchronize()
    mark(rank, f"{len(getattr(stepper, 'graphs', {}))} graphs captured; barrier")
    barrier()
    mark(rank, "warm-up")

    # ---- warm-up: exactly W steps of the timed kind -----------------------------------------------------------------------
    for i in range(args.warmup):
        stepper.step(*resident[i % len(resident)])
    torch.cuda.synchronize()

    if os.environ.get("CGVP_BENCH_TRACE") and rank == 0:      # diagnostics only: where does a step spend CPU / GPU time
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as tp:
            for i in range(4):
                eager.step(*resident[i % len(resident)])
            torch.cuda.synchronize()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "step_trace.txt"), "w") as fh:
            fh.write(tp.key_averages().table(sort_by="cuda_time_total", row_limit=60))

    # ---- timed region 1: device-resident pool --------------------------------------------------------------------------------
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    clocks = ClockSampler(local)
    mark(rank, "timed region 1")
    barrier()
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pairs_done = 0
    for i, (a, b) in enumerate(ev):
        flush.fill_(1)                        # L2 flush between timed iterations (outside the events)
        bt, m = resident[i % len(resident)]
        a.record()
        stepper.step(bt, m)
        b.record()
        pairs_done += m["pairs"]
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms, float(pairs_done)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_total, pairs_total = float(tm[0]), float(t[1])
    else:
        ms_total, pairs_total = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = pairs_total / (ms_total / 1e3)

    # per-kernel device time of the SAME steps: the library brackets each main kernel with CUDA events on the launching
    # stream; that needs host calls, so this pass launches eagerly (a graph replay makes none).  L2 flushed as above.
    mark(rank, "per-kernel pass")
    _lib.profile_enable(True)
    edges_real = 0
    for i in range(args.steps):
        flush.fill_(1)
        bt, m = resident[i % len(resident)]
        eager.step(bt, m)
        edges_real += KNN * m["nodes"]
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    dominant = max(("conv_fwd", "conv_bwd", "rows_fwd", "rows_bwd", "featurize"), key=lambda k: prof[k][0])

    # ---- timed region 2: end to end from pinned host buffers (H2D on a copy stream one step ahead, loss read one step late) --
    copy_stream = torch.cuda.Stream(dev)
    # two device staging slots sized for the largest batch of the pool (no allocation inside the timed loop)
    cap = {k: max(t[k].numel() for t, _ in pool) for k in pool[0][0]}
    stage = [{k: torch.empty(cap[k], dtype=pool[0][0][k].dtype, device=dev) for k in cap} for _ in range(2)]
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    main_stream = torch.cuda.current_stream()

    def prefetch(it):
        s = it & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s])          # the step that last used this slot has finished
            tt, m = pool[it % len(pool)]
            views = {}
            for k, v in tt.items():
                views[k] = stage[s][k][:v.numel()].view(v.shape)
                views[k].copy_(v, non_blocking=True)
            slots[s] = (views, m)
            ready[s].record(copy_stream)

    for s in range(2):
        done[s].record()
    e2e_steps = args.steps
    mark(rank, "timed region 2 (e2e)")
    barrier()
    t0 = time.perf_counter()
    prefetch(0)
    loss_host, e2e_pairs, e2e_h2d = 0.0, 0, 0
    loss_pin = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    for it in range(e2e_steps):
        cur = it & 1
        if it + 1 < e2e_steps:
            prefetch(it + 1)
        main_stream.wait_event(ready[cur])
        bt, m = slots[cur]
        loss = stepper.step(bt, m)
        done[cur].record()
        loss_pin[cur].copy_(loss, non_blocking=True)  # D2H read of the step's result ...
        loss_ev[cur].record()
        e2e_pairs += m["pairs"]
        e2e_h2d += h2d_bytes[it % len(pool)]
        if it > 0:
            loss_ev[cur ^ 1].synchronize()            # ... consumed one step later
            loss_host = float(loss_pin[cur ^ 1])
    loss_ev[(e2e_steps - 1) & 1].synchronize()
    loss_host = float(loss_pin[(e2e_steps - 1) & 1])
    barrier()
    e2e_sec = time.perf_counter() - t0
    clock_info = clocks.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and e2e)
    t = torch.tensor([e2e_sec, float(e2e_pairs)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e_value = float(t[1]) / float(tm[0])
    else:
        e2e_value = float(t[1]) / float(t[0])

    mark(rank, "gathering shapes")
    shapes = [(m["nodes"], KNN * m["nodes"], m["atoms"], m["n_pad"]) for _, m in pool]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, shapes)
    else:
        gathered = [shapes]
    def finish():
        """Leave without hanging: drop the graphs that captured the all-reduce, then tear the process group down."""
        for st in (stepper, eager):
            st.close()
        if world > 1:
            clean = parallel.shutdown()
            sys.stdout.flush(); sys.stderr.flush()
            if not clean:
                os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    kms, kn = prof[dominant]
    roof = {"bound": "hbm", "kernel": dominant + "_kernel", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "launches": kn, "avg_ms": kms / max(kn, 1),
            "kernel_timing": "CUDA events around each launch of the kernel in an eager pass over the same steps (L2 flushed between steps)",
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]}}
    roof["share_of_step"] = roof["kernel_ms_per_step"][dominant] / max(ms_per_step, 1e-9) if kn else None
    if dominant in ("conv_fwd", "conv_bwd") and kn:
        # two launches per step (two conv layers), each over that batch's edges; algorithmic bytes count the REAL edges only
        per_edge = conv_bytes_per_edge(dominant, float(KNN))
        alg_bytes_total = per_edge * edges_real * (kn / args.steps)
        roof["achieved"] = alg_bytes_total / (kms * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = alg_bytes_total / kn
        roof["algorithmic_bytes_per_edge"] = per_edge
        flops = 5086.0 * edges_real * (kn / args.steps) * (2.0 if dominant == "conv_bwd" else 1.0)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_peak = sm_count * 128 * 2 * (clock_info or {}).get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        roof["fp32"] = {"achieved_tflops": flops / (kms * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                        "frac": flops / (kms * 1e-3) / 1e12 / fp32_peak,
                        "note": "algorithmic FLOPs (5 086 / edge forward, 2x for the backward; recompute not counted) against FFMA peak = SMs x 128 lanes x 2 x max SM clock"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(traffic_file):
        tj = json.load(open(traffic_file))
        roof["traffic"] = tj.get(dominant)
        roof["traffic_note"] = tj.get(dominant + "_note")

    # ---- reference baselines on this box (bounded samples, rank 0, N = 1 only) --------------------------------------------------
    cpu = ref_gpu = None
    if world == 1 and not (args.no_cpu_baseline and args.no_reference_gpu):
        from oracle import pipeline
        threads = os.cpu_count() or 1
        use = pool[:2]
        feats = []
        for tt, m in use:
            r = real_part(tt, m)
            feats.append(pipeline.featurize_batch(r["coords"], r["ptr"], r["idents"], ds.aa_table, KNN, "num", True))
        if not args.no_cpu_baseline:
            sec, pairs, _ = reference_steps(use, ds.aa_table, 2, 1, threads, "cpu", feats)
            cpu = {"value": pairs / sec, "unit": "pairs/s", "cores": threads, "kind": "port",
                   "sample": f"2 full steps (fwd+bwd+Adam) after 1 warm-up over the pool's first two batches; {sec:.2f} s/step"}
        if not args.no_reference_gpu:
            sec, pairs, _ = reference_steps(use, ds.aa_table, 6, 3, threads, "cuda", feats)
            ref_gpu = {"value": pairs / sec, "unit": "pairs/s", "ms_per_step": sec * 1e3, "kind": "port",
                       "what": "the reference port (oracle/) in eager fp32 on this GPU, allow_tf32=False: stock index_select / index_add_ / cuBLAS path, featurized graphs resident",
                       "sample": "6 steps after 3 warm-up over the pool's first two batches (wall clock with device sync)"}

    graphs_captured = len(getattr(stepper, "graphs", {}))
    config5 = None
    if world == 1 and not args.no_config5:
        mark(rank, "config 5 (isolated GVPConvLayer at (100,16)/(32,1), child process)")
        stepper.close()
        eager.close()
        torch.cuda.empty_cache()
        config5 = run_child(["--impl", "config5", "--max-seconds", "240"], timeout=300)

    kiba = None
    if world == 1 and args.shape == "davis" and not args.no_kiba:
        mark(rank, "KIBA-shape sub-run (config 3 shape, child process)")
        kiba = kiba_shape_subrun(args)

    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "launch_mode": graph_note, "graphs_captured": graphs_captured,
        "batches": {"per_rank_real_nodes_edges_atoms_padded_nodes": gathered, "params": opt.numel,
                    "edges_per_s_conv": 2 * edges_real * world / max(ms_total / 1e3, 1e-9),
                    "note": "each rank takes the parallel.shard_by_cost share (equal pair counts, balanced edge totals) of every global batch"},
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": e2e_h2d / e2e_steps, "d2h_bytes_per_step": 4,
                "note": "each step's batch (backbone coordinates, residue types, ligand graph, targets) is copied from pinned host memory one step ahead on a copy stream; the residue graph is featurized on the device inside the step; every step's loss is copied to pinned host memory and read by the host one step later",
                "last_loss": loss_host},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_note": "C-ABI calls into libcastergvp.so per step x steps (each enqueues 1-6 kernels; replayed from CUDA graphs when launch_mode says so)",
        "roofline": roof,
        "cpu_baseline": cpu,
        "reference_gpu": ref_gpu,
        "config5_gvpconvlayer": config5,
        "config3_kiba_shape_1gpu": kiba,
    }
    emit(out)
    try:
        finish()
    except Exception:                                                  # noqa: BLE001  -- the line is out: leave cleanly
        import traceback
        traceback.print_exc(file=sys.stderr)
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
