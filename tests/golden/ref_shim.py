"""Test-only dependency shim that lets the UNMODIFIED reference import in this container.

The reference (`/root/reference/models/*.py`, `utils/create_protein_features.py`) hard-imports
`torch_geometric`, `torch_scatter` and `ipdb`, none of which are installed and none of which can be
fetched (no network).  This module registers minimal stand-ins in `sys.modules` that implement exactly
the third-party semantics the hot path relies on:

* `torch_geometric.nn.MessagePassing.propagate`: `*_i` args are gathered with `edge_index[1]`
  (target), `*_j` with `edge_index[0]` (source); messages are summed at `edge_index[1]` with
  `dim_size = N`; `aggr='mean'` divides by `max(in_degree, 1)` (PyG >= 2.5 `scatter(reduce='mean')`).
* `torch_geometric.nn.MLP`, `GINEConv`, `torch_geometric.utils.to_dense_batch`, `degree`,
  `torch_geometric.data.Data`.
* `torch_scatter.scatter_add`.

It is used ONLY by `tests/golden/make_golden.py` (fixture generation) and by CPU tests that are skipped
when `/root/reference` is absent (the GPU box).  Nothing under `caster_dta_b200/` imports it.
"""
import inspect
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class _MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr
        self.explain = False

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v):
                n = v.shape[0]
                break
        names = list(inspect.signature(self.message).parameters)
        args = {}
        for name in names:
            if name.endswith("_i"):
                args[name] = kwargs[name[:-2]].index_select(0, dst)
            elif name.endswith("_j"):
                args[name] = kwargs[name[:-2]].index_select(0, src)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        out = msg.new_zeros((n,) + tuple(msg.shape[1:])).index_add_(0, dst, msg)
        if self.aggr == "mean":
            cnt = torch.bincount(dst, minlength=n).clamp(min=1).to(msg.dtype)
            out = out / cnt.view(-1, *([1] * (msg.dim() - 1)))
        elif self.aggr not in ("add", "sum"):
            raise NotImplementedError(self.aggr)
        return out


class _MLP(nn.Module):
    """PyG `MLP(channel_list, act=..., act_first=False, norm=None)`: plain last layer, no dropout."""

    def __init__(self, channel_list, act=None, act_first=False, norm=None, norm_kwargs=None, **kw):
        super().__init__()
        assert norm is None
        self.lins = nn.ModuleList(
            [nn.Linear(a, b) for a, b in zip(channel_list[:-1], channel_list[1:])])
        self.act = act

    def forward(self, x):
        for lin in self.lins[:-1]:
            x = self.act(lin(x))
        return self.lins[-1](x)


class _GINEConv(_MessagePassing):
    def __init__(self, nn_module, eps=0.0, train_eps=False, edge_dim=None, aggr="add", **kw):
        super().__init__(aggr=aggr)
        self.nn = nn_module
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))
        in_channels = nn_module.lins[0].in_features
        self.lin = nn.Linear(edge_dim, in_channels) if edge_dim is not None else None

    def forward(self, x, edge_index, edge_attr=None):
        e = self.lin(edge_attr) if self.lin is not None else edge_attr
        out = self.propagate(edge_index, x=x, edge_attr=e)
        return self.nn(out + (1 + self.eps) * x)

    def message(self, x_j, edge_attr):
        return (x_j + edge_attr).relu()


def _to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    if batch is None:
        return x.unsqueeze(0), torch.ones(1, x.shape[0], dtype=torch.bool, device=x.device)
    b = int(batch.max()) + 1 if batch_size is None else batch_size
    counts = torch.bincount(batch, minlength=b)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    m = int(counts.max()) if max_num_nodes is None else max_num_nodes
    pos = torch.arange(x.shape[0], device=x.device) - ptr[batch]
    out = x.new_full((b, m) + tuple(x.shape[1:]), fill_value)
    out[batch, pos] = x
    mask = torch.zeros(b, m, dtype=torch.bool, device=x.device)
    mask[batch, pos] = True
    return out, mask


def _degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    return torch.bincount(index, minlength=n).to(dtype or torch.float32)


class _Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def install(reference_root=REFERENCE_ROOT):
    """Register the stand-ins and put the reference on sys.path. Idempotent."""
    if "torch_geometric" not in sys.modules:
        ipdb = types.ModuleType("ipdb")
        ipdb.set_trace = lambda *a, **k: None
        sys.modules["ipdb"] = ipdb

        ts = types.ModuleType("torch_scatter")

        def scatter_add(src, index, dim=0, out=None, dim_size=None):
            size = list(src.shape)
            size[dim] = int(index.max()) + 1 if dim_size is None else dim_size
            return src.new_zeros(size).index_add_(dim, index, src)

        ts.scatter_add = scatter_add
        sys.modules["torch_scatter"] = ts

        pyg = types.ModuleType("torch_geometric")
        pyg_nn = types.ModuleType("torch_geometric.nn")
        pyg_utils = types.ModuleType("torch_geometric.utils")
        pyg_data = types.ModuleType("torch_geometric.data")
        pyg_nn.MessagePassing = _MessagePassing
        pyg_nn.MLP = _MLP
        pyg_nn.GINEConv = _GINEConv
        for unused in ("GATv2Conv", "HEATConv", "GINConv", "GPSConv", "PNAConv", "GATConv",
                       "BatchNorm", "AttentiveFP", "global_mean_pool", "global_add_pool"):
            setattr(pyg_nn, unused, object)
        pyg_nn.models = types.SimpleNamespace(AttentiveFP=object)
        pyg_utils.to_dense_batch = _to_dense_batch
        pyg_utils.degree = _degree
        pyg_data.Data = _Data
        pyg.nn, pyg.utils, pyg.data = pyg_nn, pyg_utils, pyg_data
        sys.modules["torch_geometric"] = pyg
        sys.modules["torch_geometric.nn"] = pyg_nn
        sys.modules["torch_geometric.utils"] = pyg_utils
        sys.modules["torch_geometric.data"] = pyg_data
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def available(reference_root=REFERENCE_ROOT):
    import os
    return os.path.isfile(os.path.join(reference_root, "models", "gvp_layers.py"))
