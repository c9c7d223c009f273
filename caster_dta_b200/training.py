"""The training step of CASTER-DTA (`train_model.py:555-575`: forward, MSE loss, backward, optimizer step) on fixed-shape
("bucketed") batches, replayed as ONE CUDA graph per bucket.

Everything between the arrival of a padded batch in device memory and the updated weights is inside the captured
work -- nothing is cached across batches:

    coordinates -> node features + kNN residue graph + edge features   (featurizer.protein_graph_batch)
                -> dst-sorted / src-sorted CSR plan                      (cgvp_plan_build)
                -> JointGNN forward -> weighted MSE over the real pairs -> backward
                -> gradient pack + ONE all-reduce (NCCL) + fused Adam     (parallel.FlatAdam)

A replay after copying another batch of the same bucket into the static input buffers therefore computes that batch's
step (`tests/test_gpu_parity.py::test_bucketed_graph_step_matches_oracle_on_two_batches`).
"""
import os
import types

import torch

from . import ops
from .featurizer import knn_edge_count, protein_graph_batch

INPUT_KEYS = ("coords", "idents", "ptr", "m_x", "m_ei", "m_ea", "m_nt", "m_et", "m_batch", "y", "w")


def bucket_key(meta, edge_thresh, thresh_type, keep_self_loops=True):
    """Everything that fixes the shapes (and launch geometry) of a step: padded sizes, edge count, length bounds."""
    e = knn_edge_count(meta["lengths"], edge_thresh, thresh_type, keep_self_loops)
    return (meta["slots"], meta["n_pad"], e, meta["a_pad"], meta["me_pad"])


class BucketedTrainStep:
    """`step(batch)` runs one optimizer step on a padded device batch (dict of tensors with `INPUT_KEYS`, see
    `loader.pad_pairs`) and returns the loss tensor (device, not synchronised).

    launch_mode "graph": one captured CUDA graph per bucket key (captured on first use, all graphs share one memory
    pool); "eager": the same work launched kernel by kernel."""

    def __init__(self, model, flat_adam, aa_table, edge_thresh=30, thresh_type="num", keep_self_loops=True,
                 max_len=1056, max_atoms=128, launch_mode="graph", update=True, capture_warmup=2, record_masks=False,
                 collective_in_graph=None):
        if thresh_type not in ("num", "prop"):
            raise ValueError("fixed-shape steps need an edge count known on the host: 'num' or 'prop' graphs")
        self.model, self.opt = model, flat_adam
        self.dev = next(model.parameters()).device
        self.aa_table = aa_table.to(self.dev)
        self.edge_thresh, self.thresh_type, self.keep_self_loops = edge_thresh, thresh_type, keep_self_loops
        self.max_len, self.max_atoms = int(max_len), int(max_atoms)
        self.launch_mode, self.update = launch_mode, update      # update=False: gradients only, no optimizer step (tests)
        self.capture_warmup = capture_warmup
        self.graphs = {}            # bucket key -> namespace(graph, static, loss, pred, masks)
        self.pool = None
        self.record_masks = record_masks      # parity tests: every dropout site leaves its mask (ops.MASK_LOG)
        self.last_masks = None
        # Where the all-reduce (and the Adam kernel after it) run when there is more than one rank:
        #   inside each bucket's graph  -- validated on 2 GPUs (tests/test_gpu_multi.py); with 8 ranks whose pools hold different
        #                                  bucket sets the captured-collective path hung on this box (round 2), so it is opt-in
        #                                  (CGVP_ALLREDUCE_IN_GRAPH=1);
        #   eagerly after the replay    -- default for N > 1: one NCCL launch + one fused Adam launch per step outside the graph.
        if collective_in_graph is None:
            collective_in_graph = flat_adam.world() == 1 or os.environ.get("CGVP_ALLREDUCE_IN_GRAPH", "0") == "1"
        self.collective_in_graph = bool(collective_in_graph)

    # ---- the work of one step on the tensors of `b` -------------------------------------------------------------------
    def body(self, b, num_edges, slots, warm=False):
        """`warm`: capture warm-up -- everything except the collective and the weight update, so that ranks which capture
        different buckets at different times neither wait for each other nor let their replicas drift apart."""
        prot = protein_graph_batch(b["coords"], b["ptr"], b["idents"], self.aa_table, self.edge_thresh, self.thresh_type,
                                   self.keep_self_loops, max_len=self.max_len, num_edges=num_edges)
        prot.update(num_graphs=slots, max_nodes=self.max_len)
        mol = dict(x=b["m_x"], edge_index=b["m_ei"], ntypes=b["m_nt"], etypes=b["m_et"], eattr=b["m_ea"], batch=b["m_batch"],
                   num_graphs=slots, max_nodes=self.max_atoms)
        pred, _ = self.model(prot, mol)
        err = pred.squeeze(-1) - b["y"]
        loss = (b["w"] * err * err).sum()          # mean-squared error over the real pairs (dummy pairs weigh 0)
        self.opt.reset()
        loss.backward()
        ops.join_wgrad_stream()
        tail = self.collective_in_graph or self.launch_mode != "graph"       # the collective + Adam belong to this body
        self.opt.sync(collective=tail and not warm)
        if self.update and tail and not warm:
            self.opt.step()
        return loss.detach(), pred.detach()

    def _check(self, meta):
        if meta["max_len"] > self.max_len or meta["max_atoms"] > self.max_atoms:
            raise ValueError(f"batch exceeds the step's bounds: longest protein {meta['max_len']} > {self.max_len} or "
                             f"largest ligand {meta['max_atoms']} > {self.max_atoms}")

    def key_of(self, meta):
        return bucket_key(meta, self.edge_thresh, self.thresh_type, self.keep_self_loops)

    def prepare(self, batch, meta):
        """Capture the graph of this batch's bucket if it does not exist yet (its warm-up runs forward + backward only: no
        collective, no weight update)."""
        self._check(meta)
        key = self.key_of(meta)
        if self.launch_mode == "graph" and key not in self.graphs:
            self._capture(key, batch)
        return key

    def step(self, batch, meta):
        key = self.prepare(batch, meta)
        if self.launch_mode != "graph":
            if self.record_masks:
                ops.MASK_LOG = {}
            try:
                loss, self.last_pred = self.body(batch, key[2], key[0])
                self.last_masks = ops.MASK_LOG
            finally:
                ops.MASK_LOG = None
            return loss
        entry = self.graphs[key]
        for k in INPUT_KEYS:
            entry.static[k].copy_(batch[k], non_blocking=True)
        entry.graph.replay()
        if not self.collective_in_graph:          # N > 1 default: the collective and the fused Adam kernel follow the replay
            self.opt.all_reduce()
            if self.update:
                self.opt.step()
        self.last_pred, self.last_masks = entry.pred, entry.masks
        return entry.loss

    def close(self):
        """Drop the captured graphs.  Call before `torch.distributed.destroy_process_group()`: a live CUDA graph that captured
        NCCL collectives keeps the communicator in use and its destruction blocks."""
        self.graphs.clear()
        self.pool = None
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)

    def _capture(self, key, batch):
        static = {k: batch[k].to(self.dev, copy=True) for k in INPUT_KEYS}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        masks = None
        try:
            if self.record_masks:
                ops.MASK_LOG = {}
            with torch.cuda.stream(side):             # warm-up on a side stream (allocator, cuBLAS handles, lazy inits)
                for _ in range(self.capture_warmup):
                    self.body(static, key[2], key[0], warm=True)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            if self.record_masks:
                ops.MASK_LOG = masks = {}
            graph = torch.cuda.CUDAGraph()
            kw = {} if self.pool is None else {"pool": self.pool}
            with torch.cuda.graph(graph, **kw):
                loss, pred = self.body(static, key[2], key[0])
        finally:
            ops.MASK_LOG = None
        if self.pool is None:
            self.pool = graph.pool()
        self.graphs[key] = types.SimpleNamespace(graph=graph, static=static, loss=loss, pred=pred, masks=masks)
        return self.graphs[key]


class BucketedInferenceStep:
    """Affinity prediction (`inference/evaluation.py:43-46`: eval mode, no gradients) on padded batches, one CUDA graph per
    bucket: featurizer + plan build + protein encoder + ligand encoder + cross attention + head.

    Unique-protein cache: the protein encoder's output depends on the protein only, so a sweep that pairs
    each protein batch with several ligand batches calls `embed_proteins` once (graph A: coordinates -> residue embeddings +
    the packed <-> padded row map) and `predict_cached` per ligand batch (graph B: ligand encoder + cross attention + head on
    the cached embeddings)."""

    def __init__(self, model, aa_table, edge_thresh=30, thresh_type="num", keep_self_loops=True, max_len=2080, max_atoms=130,
                 launch_mode="graph", capture_warmup=1):
        if thresh_type not in ("num", "prop") and launch_mode == "graph":
            raise ValueError("fixed-shape graphs need an edge count known on the host ('num' / 'prop'); use launch_mode='eager'")
        self.model = model
        self.dev = next(model.parameters()).device
        self.aa_table = aa_table.to(self.dev)
        self.edge_thresh, self.thresh_type, self.keep_self_loops = edge_thresh, thresh_type, keep_self_loops
        self.max_len, self.max_atoms = int(max_len), int(max_atoms)
        self.launch_mode, self.capture_warmup = launch_mode, capture_warmup
        self.graphs, self.pool = {}, None

    PROT_KEYS = ("coords", "idents", "ptr")
    MOL_KEYS = ("m_x", "m_ei", "m_ea", "m_nt", "m_et", "m_batch")

    def _edges(self, meta):
        if self.thresh_type == "dist":
            return None
        return knn_edge_count(meta["lengths"], self.edge_thresh, self.thresh_type, self.keep_self_loops)

    def _protein_dict(self, b, num_edges, slots):
        prot = protein_graph_batch(b["coords"], b["ptr"], b["idents"], self.aa_table, self.edge_thresh, self.thresh_type,
                                   self.keep_self_loops, max_len=self.max_len if num_edges is not None else None,
                                   num_edges=num_edges)
        prot.update(num_graphs=slots, max_nodes=self.max_len)
        return prot

    def _mol_dict(self, b, slots):
        return dict(x=b["m_x"], edge_index=b["m_ei"], ntypes=b["m_nt"], etypes=b["m_et"], eattr=b["m_ea"], batch=b["m_batch"],
                    num_graphs=slots, max_nodes=self.max_atoms)

    # ---- bodies ---------------------------------------------------------------------------------------------------------
    def _body_full(self, b, num_edges, slots):
        pred, _ = self.model(self._protein_dict(b, num_edges, slots), self._mol_dict(b, slots))
        return pred

    def _body_embed(self, b, num_edges, slots):
        from .joint import DenseIndex
        prot = self._protein_dict(b, num_edges, slots)
        hints = {k: prot.pop(k) for k in ("num_graphs", "max_nodes")}
        embed = self.model.protein_gnn(**prot)
        dense = DenseIndex(prot["batch"], embed.shape[0], device=embed.device, **hints)
        return types.SimpleNamespace(embed=embed, batch=prot["batch"], dense=dense, slots=slots)

    def _body_cached(self, h, b):
        prot = dict(batch=h.batch, num_graphs=h.slots, max_nodes=self.max_len, protein_embed=h.embed, dense_index=h.dense)
        pred, _ = self.model(prot, self._mol_dict(b, h.slots))
        return pred

    # ---- graph plumbing -----------------------------------------------------------------------------------------------------
    def _run(self, key, keys, batch, fn):
        """Replay (capturing on first use) `fn(static inputs)` for the bucket `key`."""
        if self.launch_mode != "graph":
            with torch.no_grad():
                return fn(batch)
        entry = self.graphs.get(key)
        if entry is None:
            static = {k: batch[k].to(self.dev, copy=True) for k in keys}
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(cur)
            with torch.no_grad():
                with torch.cuda.stream(side):
                    for _ in range(self.capture_warmup):
                        fn(static)
                cur.wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                kw = {} if self.pool is None else {"pool": self.pool}
                with torch.cuda.graph(graph, **kw):
                    out = fn(static)
            if self.pool is None:
                self.pool = graph.pool()
            entry = self.graphs[key] = types.SimpleNamespace(graph=graph, static=static, out=out)
        for k in keys:
            entry.static[k].copy_(batch[k], non_blocking=True)
        entry.graph.replay()
        return entry.out

    def _check(self, meta):
        if meta["max_len"] > self.max_len or meta["max_atoms"] > self.max_atoms:
            raise ValueError(f"batch exceeds the step's bounds: longest protein {meta['max_len']} > {self.max_len} or "
                             f"largest ligand {meta['max_atoms']} > {self.max_atoms}")

    def predict(self, batch, meta):
        """[slots, 1] standardised affinities (rows >= meta['pairs'] belong to the dummy pairs)."""
        self._check(meta)
        e = self._edges(meta)
        key = ("full", meta["slots"], meta["n_pad"], e, meta["a_pad"], meta["me_pad"])
        return self._run(key, self.PROT_KEYS + self.MOL_KEYS, batch, lambda b: self._body_full(b, e, meta["slots"]))

    def embed_proteins(self, batch, meta):
        self._check(meta)
        e = self._edges(meta)
        key = ("embed", meta["slots"], meta["n_pad"], e)
        h = self._run(key, self.PROT_KEYS, batch, lambda b: self._body_embed(b, e, meta["slots"]))
        h.key = key
        return h

    def predict_cached(self, handle, batch, meta):
        self._check(meta)
        key = ("cached", handle.key, meta["a_pad"], meta["me_pad"])
        return self._run(key, self.MOL_KEYS, batch, lambda b: self._body_cached(handle, b))

    def close(self):
        self.graphs.clear()
        self.pool = None
