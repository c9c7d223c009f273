"""Protein residue-graph encoder (the driver of the GVP hot path), drop-in for the reference's
`models/protein_gnn.py`: `SelectableProteinModelWrapper` (`:14-82`) + `VectorProteinGNN_LBAModel` (`:289-388`).

Constructor kwargs are exactly the `protein_gnn_kwargs` of `model_kwargs.json`; `state_dict` keys are identical
(`gnn_model.gvp_node.0.wh.weight`, ...), so the protein slice of the shipped checkpoint loads with strict=True.

The forward pass is 7 fused launches + weight packing: node embed, edge embed (written directly in dst-sorted
order), 2 x (conv, node update), readout.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as gvp
from . import ops


class VectorProteinGNN_LBAModel(nn.Module):
    def __init__(self, in_channels, edge_dim, num_ntypes, num_etypes, ntype_emb_dim, etype_emb_dim, edge_hidden_channels,
                 num_convs=1, hidden_channels=None, out_channels=8, dropout_rate=0.2, activation="relu", aggr="mean"):
        super().__init__()
        self.in_channels, self.edge_dim = tuple(in_channels), tuple(edge_dim)
        self.num_ntypes, self.num_etypes = num_ntypes, num_etypes
        self.num_convs, self.dropout_rate = num_convs, dropout_rate
        hidden_channels = out_channels if hidden_channels is None else hidden_channels
        self.hidden_channels = (hidden_channels, 0) if isinstance(hidden_channels, int) else tuple(hidden_channels)
        self.out_channels = (out_channels, 0) if isinstance(out_channels, int) else tuple(out_channels)
        self.edge_hidden_channels = tuple(edge_hidden_channels)
        # learned type embeddings (nn.Embedding) or one-hot (protein_gnn.py:123-133)
        self.ntype_emb_dim = ntype_emb_dim if ntype_emb_dim is not None else num_ntypes
        self.etype_emb_dim = etype_emb_dim if etype_emb_dim is not None else num_etypes
        if ntype_emb_dim is not None:
            self.ntype_embedding = nn.Embedding(num_ntypes, ntype_emb_dim)
        if etype_emb_dim is not None:
            self.etype_embedding = nn.Embedding(num_etypes, etype_emb_dim)
        self.dropout = nn.Dropout(dropout_rate)

        node_in = (self.in_channels[0] + self.ntype_emb_dim, self.in_channels[1])
        edge_in = (self.edge_dim[0] + self.etype_emb_dim, self.edge_dim[1])
        self.gvp_node = nn.Sequential(gvp.GVP(node_in, self.hidden_channels, activations=(None, None), vector_gate=True),
                                      gvp.LayerNorm(self.hidden_channels))
        self.gvp_edge = nn.Sequential(gvp.GVP(edge_in, self.edge_hidden_channels, activations=(None, None), vector_gate=True),
                                      gvp.LayerNorm(self.edge_hidden_channels))
        self.gvp_relu = nn.ReLU()
        self.conv_list = nn.ModuleList([
            gvp.GVPConvLayer(self.hidden_channels, self.edge_hidden_channels, drop_rate=dropout_rate,
                             activations=(self.gvp_relu, None), vector_gate=True, aggr=aggr)
            for _ in range(num_convs)])
        self.gvp_norm_before_scalar = gvp.LayerNorm(self.hidden_channels)
        self.gvp_to_scalar = gvp.GVP(self.hidden_channels, self.out_channels, activations=(self.gvp_relu, None),
                                     vector_gate=True)

    def _embed(self, seq, feats, types, onehot_classes, embedding, in_index=None):
        """[type embedding ; features] -> GVP -> LayerNorm in one launch (protein_gnn.py:139-152, 375-376)."""
        s, v = feats
        if embedding is not None:                      # learned embedding: concatenated on the host side
            s = torch.cat([embedding(types), s], -1)
            types, onehot_classes = None, 0
        g, ln = seq[0], seq[1]
        prog = gvp._row_program(s.shape[1], g.vi, (g.spec,), onehot=onehot_classes, post_norm=True)
        out = ops.run_rows(prog, s, v if g.vi else None, types=types, in_index=in_index,
                           ln1=(ln.scalar_norm.weight, ln.scalar_norm.bias), weights=g.kernel_weights())
        return out if g.vo else out[0]

    def forward(self, x, edge_index, ntypes, etypes, eattr=None, batch=None, plan=None):
        x_s, x_v = x[0], x[1]
        if eattr is None:
            raise ValueError("the LBA encoder needs edge attributes (s, V)")
        e_s, e_v = eattr[0], eattr[1]
        if plan is None:                     # `plan`: a GraphPlan of this edge_index built earlier (e.g. by the collate)
            plan = ops.get_plan(edge_index, x_s.shape[0])
        h = self._embed(self.gvp_node, (x_s, x_v), ntypes, self.num_ntypes, getattr(self, "ntype_embedding", None))
        # edge embedding is produced directly in dst-sorted order: both conv layers then stream it contiguously
        e = self._embed(self.gvp_edge, (e_s, e_v), etypes, self.num_etypes, getattr(self, "etype_embedding", None),
                        in_index=plan.perm)
        for conv in self.conv_list:
            h = conv(h, edge_index, e, plan=plan, edge_sorted=True)
        g, ln = self.gvp_to_scalar, self.gvp_norm_before_scalar
        prog = gvp._row_program(g.si, g.vi, (g.spec,), pre_norm=True)
        out_s, out_v = ops.run_rows(prog, h[0], h[1] if g.vi else None, ln0=(ln.scalar_norm.weight, ln.scalar_norm.bias),
                                    weights=g.kernel_weights())
        return (out_s, out_v) if g.vo else out_s


class SelectableProteinModelWrapper(nn.Module):
    """`models/protein_gnn.py:14-82`.  Only the GVP ('lbamodel') encoder is in scope of this package."""

    def __init__(self, in_channels, edge_dim, base_conv, **kwargs):
        super().__init__()
        if type(in_channels) is not type(edge_dim):
            raise ValueError("in_channels and edge_dim must be the same type (both ints or both (scalar, vector))")
        if base_conv != "lbamodel":
            raise NotImplementedError(f"base_conv={base_conv!r}: this package accelerates the GVP 'lbamodel' encoder only")
        if isinstance(in_channels, int):
            raise ValueError("the 'lbamodel' encoder needs (scalar, vector) input dims")
        self.base_conv = base_conv
        self.is_scalar_data = False
        self.gnn_model = VectorProteinGNN_LBAModel(in_channels=in_channels, edge_dim=edge_dim, **kwargs)

    def forward(self, x, edge_index, ntypes, etypes, eattr=None, batch=None, plan=None):
        return self.gnn_model(x, edge_index, ntypes, etypes, eattr=eattr, batch=batch, plan=plan)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("gnn_model"), name)
