"""Size-capped dynamic batching and device-side collation for protein / molecule graph pairs.

Host-side mirror of `PMD_BatchSampler` and `PMDCollator` (`dataset/dual_dataset.py:424-547`), working on plain
size arrays and tensors instead of PyG `Data` objects:

* `SizeCappedBatchSampler` packs consecutive (optionally shuffled) pairs into a mini-batch until the summed element
  count -- protein and/or molecule nodes or edges, plus the dense residue x atom attention footprint the batch will
  allocate -- would exceed `max_num` (`:464-522`);
* `collate_graphs` concatenates per-graph tensors into one batched graph with offset edge indices, `batch` and `ptr`
  vectors (what `Batch.from_data_list` produces, `:543-544`), entirely with device ops and no host synchronisation,
  ready for `ops.GraphPlan`.
"""
import torch


class SizeCappedBatchSampler:
    """Yields lists of pair indices.  Arguments follow `PMD_BatchSampler.__init__` (`dual_dataset.py:432-461`); the
    dataset is replaced by four integer sequences (per pair: protein nodes / edges, molecule nodes / edges).

    One deliberate difference: with `skip_too_big=True` the reference `continue`s past an oversized first sample
    without counting it as processed (`:505-507`), which re-visits it forever; here an oversized sample is dropped
    and counted, so the iteration always terminates."""

    def __init__(self, protein_nodes, protein_edges, molecule_nodes, molecule_edges, max_num, count_elem="edge",
                 graph_type="both", include_nodepair=True, shuffle=True, skip_too_big=False, max_bsize=None,
                 generator=None):
        if max_num <= 0:
            raise ValueError(f"`max_num` should be a positive integer value (got {max_num})")
        if count_elem not in ("node", "edge"):
            raise ValueError(f"`max_count` choice should be either 'node' or 'edge' (got '{count_elem}')")
        if graph_type not in ("protein", "molecule", "both"):
            raise ValueError("`graph_type` choice should be one of 'protein', 'molecule', or 'both' "
                             f"(got '{graph_type}')")
        self.p_nodes, self.p_edges = [int(x) for x in protein_nodes], [int(x) for x in protein_edges]
        self.m_nodes, self.m_edges = [int(x) for x in molecule_nodes], [int(x) for x in molecule_edges]
        if not (len(self.p_nodes) == len(self.p_edges) == len(self.m_nodes) == len(self.m_edges)):
            raise ValueError("size sequences must have one entry per pair")
        self.max_num, self.count_elem, self.graph_type = max_num, count_elem, graph_type
        self.include_nodepair, self.shuffle, self.skip_too_big = include_nodepair, shuffle, skip_too_big
        self.max_bsize, self.generator = max_bsize, generator

    def __len__(self):
        return len(self.p_nodes)

    def _cost(self, i):
        p = self.p_nodes[i] if self.count_elem == "node" else self.p_edges[i]
        m = self.m_nodes[i] if self.count_elem == "node" else self.m_edges[i]
        return p if self.graph_type == "protein" else m if self.graph_type == "molecule" else p + m

    def __iter__(self):
        n = len(self)
        order = torch.randperm(n, generator=self.generator).tolist() if self.shuffle else list(range(n))
        pos = 0
        while pos < n:
            samples, current, max_npair = [], 0, 0
            while pos < n:
                i = order[pos]
                num = self._cost(i)
                new_npair = max_npair
                if self.include_nodepair:
                    # the dense [pairs, max residues x max atoms] attention block grows for EVERY member when a larger
                    # pair joins (`:489-497`): charge the difference between the new and the already-charged footprint
                    new_npair = max(max_npair, self.p_nodes[i] * self.m_nodes[i])
                    num += new_npair * (len(samples) + 1) - max_npair * len(samples)
                if current + num > self.max_num:
                    if current == 0:
                        if self.skip_too_big:
                            pos += 1
                            continue
                    else:
                        break
                samples.append(i)
                pos += 1
                current += num
                max_npair = new_npair
                if self.max_bsize is not None and len(samples) >= self.max_bsize:
                    break
            if samples or not self.skip_too_big:
                yield samples


def collate_graphs(graphs, emit_plan=None):
    """Batch a list of graphs (dicts of device tensors) into one disconnected graph.

    Each dict holds `x` (tensor `[n, F]` or tuple `(s [n,S], V [n,C,3])`), `edge_index [2,e]` int64 and any of
    `edge_attr` (tensor or tuple), `node_type [n]`, `edge_type [e]`.  Returns the same keys concatenated, edge indices
    shifted by the node offsets, plus `batch [N]`, `ptr [B+1]`, `num_graphs`, `max_nodes` (python ints known from the
    shapes -- no device read).

    `emit_plan` (default: when the graphs live on a CUDA device and carry (s, V) features): also build the graph plan
    of the batched edge list -- the dst-sorted / src-sorted CSR views the fused GVPConv kernels aggregate over
    (`ops.GraphPlan`, `cgvp_plan_build`) -- and return it under `plan`, so the encoder does not sort the edges again
    (`JointGNN.forward_with_graphs` / `protein_gnn(..., plan=...)` accept it)."""
    if not graphs:
        raise ValueError("collate_graphs needs at least one graph")

    def first(v):
        return v[0] if isinstance(v, (tuple, list)) else v

    def cat(key):
        v0 = graphs[0][key]
        if isinstance(v0, (tuple, list)):
            return tuple(torch.cat([g[key][k] for g in graphs]) for k in range(len(v0)))
        return torch.cat([g[key] for g in graphs])

    dev = first(graphs[0]["x"]).device
    sizes = [int(first(g["x"]).shape[0]) for g in graphs]
    out = {"x": cat("x")}
    ptr_host = [0]
    for n in sizes:
        ptr_host.append(ptr_host[-1] + n)
    ptr = torch.tensor(ptr_host, dtype=torch.int64).to(dev, non_blocking=True)
    n_edges = torch.tensor([int(g["edge_index"].shape[1]) for g in graphs], dtype=torch.int64)
    shift = torch.repeat_interleave(torch.tensor(ptr_host[:-1], dtype=torch.int64), n_edges).to(dev, non_blocking=True)
    out["edge_index"] = torch.cat([g["edge_index"] for g in graphs], 1) + shift
    for key in ("edge_attr", "node_type", "edge_type"):
        if key in graphs[0]:
            out[key] = cat(key)
    out["batch"] = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), torch.tensor(sizes)).to(dev, non_blocking=True)
    out["ptr"] = ptr
    out["num_graphs"] = len(graphs)
    out["max_nodes"] = max(sizes)
    if emit_plan is None:
        emit_plan = dev.type == "cuda" and isinstance(graphs[0]["x"], (tuple, list))
    if emit_plan:
        from . import ops
        out["plan"] = ops.GraphPlan(out["edge_index"], ptr_host[-1])
    return out
