"""CASTER-DTA joint model around the accelerated protein encoder.

Only `protein_gnn` (the GVP stack) runs on the custom kernels.  The molecule GINE encoder, the cross-attention
block and the MLP head are small dense / scalar graph ops that stay stock PyTorch ("host code stays PyTorch");
they are written here without torch_geometric so that `JointGNN(**model_kwargs.json)` builds and the shipped
checkpoint loads with strict=True:

    JointGNN                <- models/joint_gnn.py:15-288
    CrossAttentionModule    <- models/joint_gnn.py:321-408
    GINE molecule encoder   <- models/molecule_gnn.py:208-280 (+ PyG GINEConv / MLP semantics)
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .encoder import SelectableProteinModelWrapper


def _activation(name):
    table = {"relu": nn.ReLU, "leaky_relu": lambda: nn.LeakyReLU(0.01), "tanh": nn.Tanh, "sigmoid": nn.Sigmoid,
             "gelu": nn.GELU, "elu": nn.ELU, "selu": nn.SELU, "silu": nn.SiLU, "swish": nn.SiLU, "none": nn.Identity}
    if isinstance(name, nn.Module):
        return name
    return table[name.lower()]()


class _MLP(nn.Module):
    """Two-layer perceptron with the parameter names of PyG's `MLP` (`lins.0`, `lins.1`), plain last layer."""

    def __init__(self, dims, act):
        super().__init__()
        self.lins = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.act = act

    def forward(self, x):
        for lin in self.lins[:-1]:
            x = self.act(lin(x))
        return self.lins[-1](x)


class GINEConv(nn.Module):
    """out_i = nn((1 + eps) x_i + sum_j relu(x_j + lin(e_ji)))  (PyG GINEConv)."""

    def __init__(self, mlp, edge_dim, train_eps=True):
        super().__init__()
        self.nn = mlp
        if train_eps:
            self.eps = nn.Parameter(torch.zeros(1))
        else:
            self.register_buffer("eps", torch.zeros(1))
        self.lin = nn.Linear(edge_dim, mlp.lins[0].in_features)

    def forward(self, x, edge_index, edge_attr):
        msg = F.relu(x.index_select(0, edge_index[0]) + self.lin(edge_attr))
        agg = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
        return self.nn(agg + (1 + self.eps) * x)


class HomoMoleculeGNN_GINE(nn.Module):
    def __init__(self, in_channels, edge_dim, num_ntypes, num_etypes, ntype_emb_dim, etype_emb_dim, num_convs=1,
                 hidden_channels=None, out_channels=8, dropout_rate=0.2, activation="relu", aggr="sum",
                 gin_trainable_eps=True, **unused):
        super().__init__()
        if ntype_emb_dim is not None or etype_emb_dim is not None:
            raise NotImplementedError("learned type embeddings are not used by the shipped configuration")
        if aggr not in ("sum", "add"):
            raise NotImplementedError("GINE stand-in implements sum aggregation")
        self.num_ntypes, self.num_etypes, self.out_channels = num_ntypes, num_etypes, out_channels
        hidden = out_channels if hidden_channels is None else hidden_channels
        self.activation = _activation(activation)
        self.dropout = nn.Dropout(dropout_rate)
        dims = [in_channels + num_ntypes] + [hidden] * (num_convs - 1) + [out_channels]
        self.conv_list = nn.ModuleList(
            GINEConv(_MLP([dims[i], dims[i + 1], dims[i + 1]], self.activation), edge_dim + num_etypes, gin_trainable_eps)
            for i in range(num_convs))

    def forward(self, x, edge_index, ntypes, etypes, eattr=None, batch=None):
        x = torch.cat([_one_hot(ntypes, self.num_ntypes, x.dtype), x], -1)
        e = torch.cat([_one_hot(etypes, self.num_etypes, x.dtype), eattr], -1)
        for k, conv in enumerate(self.conv_list[:-1]):
            x = ops.dropout(self.dropout, self.activation(conv(x, edge_index, e)), f"gine.{k}")
        return self.activation(self.conv_list[-1](x, edge_index, e))


def _one_hot(idx, num_classes, dtype):
    """`F.one_hot` without its device->host range check (a sync, and illegal during CUDA-graph capture)."""
    return (idx.unsqueeze(-1) == torch.arange(num_classes, device=idx.device)).to(dtype)


class SelectableMoleculeModelWrapper(nn.Module):
    def __init__(self, base_conv, **kwargs):
        super().__init__()
        if base_conv.lower() != "gine":
            raise NotImplementedError(f"molecule encoder {base_conv!r}: only the shipped 'gine' configuration is provided")
        self.base_conv = base_conv.lower()
        self.gnn_model = HomoMoleculeGNN_GINE(**kwargs)

    def forward(self, x, edge_index, ntypes, etypes, eattr=None, batch=None):
        return self.gnn_model(x, edge_index, ntypes, etypes, eattr=eattr, batch=batch)

    @property
    def out_channels(self):
        return self.gnn_model.out_channels


def to_dense_batch(x, batch, num_graphs=None, max_nodes=None):
    """[total, D] + graph id per row -> ([B, max, D], mask [B, max]).  Pass num_graphs / max_nodes to avoid the
    device->host syncs of computing them."""
    if batch is None:
        return x.unsqueeze(0), torch.ones(1, x.shape[0], dtype=torch.bool, device=x.device)
    b = int(batch[-1]) + 1 if num_graphs is None else int(num_graphs)
    counts = torch.zeros(b, dtype=torch.long, device=x.device).index_add_(0, batch, torch.ones_like(batch))   # bincount syncs
    start = torch.cumsum(counts, 0) - counts
    m = int(counts.max()) if max_nodes is None else int(max_nodes)
    pos = torch.arange(x.shape[0], device=x.device) - start[batch]
    out = x.new_zeros((b, m, x.shape[1]))
    out[batch, pos] = x
    mask = torch.arange(m, device=x.device).unsqueeze(0) < counts.unsqueeze(1)      # [B, max]: row r of graph g is real
    return out, mask


class DenseIndex:
    """Row map between the packed layout [total, D] (graphs back to back) and the padded layout [B, max, D], computed
    once per forward and shared by every pad / unpad.  The row-wise layers of the cross-attention block (LayerNorm,
    projections, feed-forward, residuals) run on packed rows -- the padded layout of the reference
    (`to_dense_batch`, `models/joint_gnn.py:197-204`) spends ~1/3 of its rows on padding at Davis-shape lengths -- and
    only the attention core sees padded tensors.  No device->host sync when num_graphs / max_nodes are given."""

    def __init__(self, batch, total, num_graphs=None, max_nodes=None, device=None):
        if batch is None:
            self.b, self.m = 1, int(total)
            self.idx = torch.arange(total, device=device)
            self.mask = torch.ones(1, total, dtype=torch.bool, device=device)
            self.batch = torch.zeros(total, dtype=torch.long, device=device)
            self.ptr = torch.tensor([0, total], dtype=torch.long, device=device)
            return
        dev = batch.device
        self.b = int(batch[-1]) + 1 if num_graphs is None else int(num_graphs)
        counts = torch.zeros(self.b, dtype=torch.long, device=dev).index_add_(0, batch, torch.ones_like(batch))
        self.ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])           # rows of a graph are contiguous
        start = self.ptr[:-1]
        self.m = int(counts.max()) if max_nodes is None else int(max_nodes)
        self.idx = batch * self.m + (torch.arange(total, device=dev) - start[batch])
        self.mask = torch.arange(self.m, device=dev).unsqueeze(0) < counts.unsqueeze(1)
        self.batch = batch.contiguous()

    def pad(self, x, fill=None):
        """[total, D] -> [B, max, D]; padding rows hold `fill` ([D]) or zeros."""
        d = x.shape[1]
        out = x.new_zeros((self.b * self.m, d)) if fill is None else fill.to(x.dtype).expand(self.b * self.m, d).clone()
        return out.index_copy(0, self.idx, x).view(self.b, self.m, d)

    def unpad(self, dense):
        return dense.reshape(self.b * self.m, dense.shape[-1]).index_select(0, self.idx)


def _lin(mod, x):
    """nn.Linear applied through ops.linear (tensor-core weight gradient for large row counts on CUDA)."""
    return ops.linear(x, mod.weight, mod.bias) if x.is_cuda and x.dim() == 2 else mod(x)


def _seq(mods, x, site=""):
    for m in mods:
        if isinstance(m, nn.Linear):
            x = _lin(m, x)
        elif isinstance(m, nn.Dropout):
            x = ops.dropout(m, x, site + ".inner")
        else:
            x = m(x)
    return x


def _mha_packed(mha, q_in, kv_in, dq, dk, q_fill_in, return_weights, training):
    """`nn.MultiheadAttention.forward(query, key, value, key_padding_mask=~mask)` (`models/joint_gnn.py:350-361`) with
    the in / out projections on packed rows.  q_fill_in: the value every padded QUERY row has in the reference (the
    LayerNorm bias -- LayerNorm of an all-zero row), so that the returned attention map matches on padded rows too."""
    e, h = mha.embed_dim, mha.num_heads
    hd = e // h
    if mha._qkv_same_embed_dim:
        wq, wkv = mha.in_proj_weight[:e], mha.in_proj_weight[e:]
        wk = wv = None
    else:
        wq, wk, wv, wkv = mha.q_proj_weight, mha.k_proj_weight, mha.v_proj_weight, None
    bq = bk = bv = bkv = None
    if mha.in_proj_bias is not None:
        bq, bkv = mha.in_proj_bias[:e], mha.in_proj_bias[e:]
        bk, bv = mha.in_proj_bias[e:2 * e], mha.in_proj_bias[2 * e:]
    lin = ops.linear if q_in.is_cuda else F.linear
    q = lin(q_in, wq, bq)
    if wkv is not None:
        k, v = lin(kv_in, wkv, bkv).split(e, dim=-1)
    else:
        k, v = lin(kv_in, wk, bk), lin(kv_in, wv, bv)
    q_fill = F.linear(q_fill_in, wq, bq) if q_fill_in is not None else None
    p_drop = mha.dropout if training else 0.0
    if q.is_cuda and p_drop == 0.0 and ops.attention_supported(h, hd):
        # fused core on packed rows (csrc/attention.cu): no padded tensors, no [B, H, Lq, Lk] scores in HBM
        o, weights, w_fill = ops.CrossAttnFunction.apply(q, k, v, dq.ptr, dk.ptr, dq.batch, dk.batch, h, dq.m, dk.m,
                                                         q_fill if return_weights else None, return_weights)
        if not return_weights:
            weights = None
        elif q_fill is not None:
            weights = torch.where(dq.mask.unsqueeze(-1), weights, w_fill.unsqueeze(1))
        return _lin(mha.out_proj, o), weights
    qd = dq.pad(q, q_fill).view(dq.b, dq.m, h, hd).transpose(1, 2)                 # [B, H, Lq, hd]
    kd = dk.pad(k).view(dk.b, dk.m, h, hd).transpose(1, 2)
    vd = dk.pad(v).view(dk.b, dk.m, h, hd).transpose(1, 2)
    key_mask = dk.mask[:, None, None, :]                                             # True = real key
    weights = None
    if return_weights:
        scores = torch.matmul(qd * (1.0 / hd) ** 0.5, kd.transpose(-2, -1)).masked_fill(~key_mask, float("-inf"))
        prob = torch.softmax(scores, dim=-1)
        if p_drop > 0.0:
            prob = F.dropout(prob, p=p_drop)
        od = torch.matmul(prob, vd)
        weights = prob.mean(dim=1)                                                   # averaged over heads
    else:
        od = F.scaled_dot_product_attention(qd, kd, vd, attn_mask=key_mask, dropout_p=p_drop)
    o = dq.unpad(od.transpose(1, 2).reshape(dq.b, dq.m, e))
    return mha.out_proj(o), weights


class CrossAttentionModule(nn.Module):
    def __init__(self, embed_dim_1, embed_dim_2, n_attention_heads, attn_dropout, include_residual_stream=True,
                 dim_feedforward_scale=2, feedforward_dropout=0.2):
        super().__init__()
        self.include_residual_stream = include_residual_stream
        self.preattn_norm1, self.preattn_norm2 = nn.LayerNorm(embed_dim_1), nn.LayerNorm(embed_dim_2)
        self.embed1_to_2 = nn.MultiheadAttention(embed_dim_1, n_attention_heads, dropout=attn_dropout, kdim=embed_dim_2,
                                                 vdim=embed_dim_2, batch_first=True)
        self.embed2_to_1 = nn.MultiheadAttention(embed_dim_2, n_attention_heads, dropout=attn_dropout, kdim=embed_dim_1,
                                                 vdim=embed_dim_1, batch_first=True)
        self.ff_norm1, self.ff_norm2 = nn.LayerNorm(embed_dim_1), nn.LayerNorm(embed_dim_2)
        self.ff_dropout = nn.Dropout(feedforward_dropout)
        self.side_stream = None              # set by JointGNN when overlap_encoders is on
        if include_residual_stream:
            def ff(d):
                return nn.Sequential(nn.Linear(d, d * dim_feedforward_scale), nn.ReLU(), nn.Dropout(feedforward_dropout),
                                     nn.Linear(d * dim_feedforward_scale, d))
            self.ff1, self.ff2 = ff(embed_dim_1), ff(embed_dim_2)

    def forward(self, e1, e2, mask1, mask2, return_weights=True):
        n1, n2 = self.preattn_norm1(e1), self.preattn_norm2(e2)
        a1, w1 = self.embed1_to_2(n1, n2, n2, key_padding_mask=~mask2, need_weights=return_weights)
        a2, w2 = self.embed2_to_1(n2, n1, n1, key_padding_mask=~mask1, need_weights=return_weights)
        if self.include_residual_stream:
            e1 = e1 + self.ff_dropout(a1)
            e1 = e1 + self.ff_dropout(self.ff1(self.ff_norm1(e1)))
            e2 = e2 + self.ff_dropout(a2)
            e2 = e2 + self.ff_dropout(self.ff2(self.ff_norm2(e2)))
        else:
            e1, e2 = a1, a2
        return e1, e2, (w1, w2)

    def forward_packed(self, x1, x2, d1, d2, return_weights=True, first=True, site="attn.0"):
        """Same block (`models/joint_gnn.py:321-408`) on packed rows x1 [N1, D1], x2 [N2, D2]; d1 / d2: DenseIndex."""
        n1, n2 = ops.layer_norm(x1, self.preattn_norm1), ops.layer_norm(x2, self.preattn_norm2)
        f1 = self.preattn_norm1.bias if first else None
        f2 = self.preattn_norm2.bias if first else None

        def stream2():                       # everything that only feeds the second stream's output (atoms: many tiny kernels)
            a2, w2 = _mha_packed(self.embed2_to_1, n2, n1, d2, d1, f2, return_weights, self.training)
            if not self.include_residual_stream:
                return a2, w2
            y2 = x2 + ops.dropout(self.ff_dropout, a2, site + ".a2")
            return y2 + ops.dropout(self.ff_dropout, _seq(self.ff2, ops.layer_norm(y2, self.ff_norm2), site + ".ff2"), site + ".ff2.outer"), w2

        side = self.side_stream if x1.is_cuda else None
        if side is not None:                 # the two directions are independent: run the second on the side stream
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                y2, w2 = stream2()
        a1, w1 = _mha_packed(self.embed1_to_2, n1, n2, d1, d2, f1, return_weights, self.training)
        if self.include_residual_stream:
            x1 = x1 + ops.dropout(self.ff_dropout, a1, site + ".a1")
            x1 = x1 + ops.dropout(self.ff_dropout, _seq(self.ff1, ops.layer_norm(x1, self.ff_norm1), site + ".ff1"), site + ".ff1.outer")
        else:
            x1 = a1
        if side is not None:
            main.wait_stream(side)
            for t in (y2, w2):
                if t is not None:
                    t.record_stream(main)
        else:
            y2, w2 = stream2()
        return x1, y2, (w1, w2)


class StackedCrossAttentionModule(nn.Module):
    def __init__(self, make_layer, num_layers):
        super().__init__()
        self.cross_attn_layers = nn.ModuleList(make_layer() for _ in range(num_layers))

    def forward(self, e1, e2, mask1, mask2, return_weights=True):
        weights = []
        for layer in self.cross_attn_layers:
            e1, e2, w = layer(e1, e2, mask1, mask2, return_weights)
            weights.append(w)
        return e1, e2, weights

    def forward_packed(self, x1, x2, d1, d2, return_weights=True):
        """Attention maps of layers after the first differ from the padded formulation on padded QUERY rows only (rows the
        reference computes from padding and never uses)."""
        weights = []
        for i, layer in enumerate(self.cross_attn_layers):
            x1, x2, w = layer.forward_packed(x1, x2, d1, d2, return_weights, first=i == 0, site=f"attn.{i}")
            weights.append(w)
        return x1, x2, weights


def _lin_stack(depth, in_dim, scale=2, norm=None):
    lins, norms, d = [], [], in_dim
    for _ in range(depth):
        o = int(d * scale)
        lins.append(nn.Linear(d, o))
        norms.append(nn.LayerNorm(o) if norm == "layer" else nn.Identity())
        d = o
    return nn.ModuleList(lins), nn.ModuleList(norms), d


class JointGNN(nn.Module):
    """Same constructor and `forward` / `forward_with_graphs` as `models/joint_gnn.py:15-170`; returns
    `(affinity [B,1], attention weights)`.  Set `return_attention=False` to let the attention use the fused SDPA path."""

    def __init__(self, protein_gnn_kwargs, molecule_gnn_kwargs, residue_lin_depth, atom_lin_depth, n_attention_heads,
                 attention_dropout, protein_lin_depth, molecule_lin_depth, pairwise_embedding_dim, out_lin_depth,
                 out_lin_factor=0.5, out_lin_norm_type=None, activation="relu", dropout=0.0, element_pooling="mean",
                 include_residual_stream=True, residual_dim_ff_scale=2, num_cross_attn_layers=1,
                 include_post_pool_layernorm=False):
        super().__init__()
        if out_lin_norm_type == "batch":
            raise NotImplementedError("batch norm in the output stack needs torch_geometric")
        self.element_pooling = element_pooling
        self.num_cross_attn_layers = num_cross_attn_layers
        self.include_post_pool_layernorm = include_post_pool_layernorm
        self.return_attention = True
        self.overlap_encoders = False        # run the molecule encoder on a side stream (set by the training driver)
        self._side_stream = None
        self.activation = _activation(activation)
        self.dropout = nn.Dropout(dropout)
        self.protein_gnn = SelectableProteinModelWrapper(**protein_gnn_kwargs)
        self.molecule_gnn = SelectableMoleculeModelWrapper(**molecule_gnn_kwargs)
        p_out = self.protein_gnn.out_channels
        p_out = p_out[0] if isinstance(p_out, tuple) else p_out
        m_out = self.molecule_gnn.out_channels
        self.residue_lins, self.residue_norms, r_dim = _lin_stack(residue_lin_depth, p_out)
        self.atom_lins, self.atom_norms, a_dim = _lin_stack(atom_lin_depth, m_out)
        if num_cross_attn_layers > 0:
            self.cross_attn_module = StackedCrossAttentionModule(
                lambda: CrossAttentionModule(r_dim, a_dim, n_attention_heads, attention_dropout, include_residual_stream,
                                             residual_dim_ff_scale, dropout), num_cross_attn_layers)
        else:
            self.cross_attn_module = None
        if include_post_pool_layernorm:
            self.protein_post_pool_norm, self.molecule_post_pool_norm = nn.LayerNorm(r_dim), nn.LayerNorm(a_dim)
        self.protein_lins, self.protein_norms, pd = _lin_stack(protein_lin_depth, r_dim)
        self.molecule_lins, self.molecule_norms, md = _lin_stack(molecule_lin_depth, a_dim)
        self.pm_embed_lin = nn.Linear(pd + md, pairwise_embedding_dim)
        self.out_fc_layers, self.out_fc_norms, od = _lin_stack(out_lin_depth, pairwise_embedding_dim, out_lin_factor,
                                                               out_lin_norm_type)
        self.output_layer = nn.Linear(od, 1)

    @staticmethod
    def _graphs_to_dicts(protein_graph, molecule_graph):
        def as_dict(g):
            if isinstance(g, dict):           # the output of `batching.collate_graphs` (PyG `Data` key names + hints + plan)
                d = {"x": g["x"], "edge_index": g["edge_index"], "ntypes": g["node_type"], "etypes": g["edge_type"],
                     "eattr": g["edge_attr"], "batch": g["batch"]}
                d.update({k: g[k] for k in ("num_graphs", "max_nodes", "plan") if k in g})
                return d
            return {"x": g.x, "edge_index": g.edge_index, "ntypes": g.node_type, "etypes": g.edge_type,
                    "eattr": g.edge_attr, "batch": g.batch}
        return as_dict(protein_graph), as_dict(molecule_graph)

    def forward_with_graphs(self, protein_graph, molecule_graph):
        return self.forward(*self._graphs_to_dicts(protein_graph, molecule_graph))

    def _stack(self, x, lins, norms, site):
        for k, (lin, norm) in enumerate(zip(lins, norms)):
            x = ops.dropout(self.dropout, self.activation(norm(_lin(lin, x))), f"{site}.{k}")
        return x

    def _pool(self, x, mask):
        if self.element_pooling == "mean":
            return (x * mask.unsqueeze(-1)).sum(1) / mask.sum(1, keepdim=True)
        if self.element_pooling == "sum":
            return (x * mask.unsqueeze(-1)).sum(1)
        if self.element_pooling == "max":
            return (x - (~mask).unsqueeze(-1) * 1.0e10).max(1).values
        raise ValueError(self.element_pooling)

    def forward(self, protein_graph_data={}, molecule_graph_data={}):
        pg, mg = dict(protein_graph_data), dict(molecule_graph_data)
        hints_p = {k: pg.pop(k, None) for k in ("num_graphs", "max_nodes")}
        hints_m = {k: mg.pop(k, None) for k in ("num_graphs", "max_nodes")}
        mg.pop("plan", None)
        # `protein_embed`: residue embeddings computed earlier for the same protein(s) -- the encoder
        # output depends on the protein only, so an inference sweep over many ligands per protein can reuse it
        embed = pg.pop("protein_embed", None)
        dp_cached = pg.pop("dense_index", None)          # optional: the DenseIndex of the same protein batch (see protein_embed)
        plan = pg.pop("plan", None)                      # optional: the GraphPlan of `edge_index` (batching.collate_graphs emits it)
        # The two encoders are independent (`models/joint_gnn.py:183-190`): the molecule side -- ~60 tiny kernels on ~10^3
        # atoms -- runs on a side stream under the protein encoder's large kernels (fork / join, also inside graph capture;
        # autograd replays each side's backward on the stream of its forward).
        side = None
        x_mol = mg.get("x")
        if self.overlap_encoders and torch.is_tensor(x_mol) and x_mol.is_cuda:
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=x_mol.device)
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                atm = self._stack(self.molecule_gnn(**mg), self.atom_lins, self.atom_norms, "atom_lins")
                dm = DenseIndex(mg.get("batch"), atm.shape[0], device=atm.device, **hints_m)
        if embed is None:
            embed = self.protein_gnn(**pg) if plan is None else self.protein_gnn(plan=plan, **pg)
        res = self._stack(embed, self.residue_lins, self.residue_norms, "residue_lins")
        dp = dp_cached if dp_cached is not None else DenseIndex(pg.get("batch"), res.shape[0], device=res.device, **hints_p)
        if side is None:
            atm = self._stack(self.molecule_gnn(**mg), self.atom_lins, self.atom_norms, "atom_lins")
            dm = DenseIndex(mg.get("batch"), atm.shape[0], device=atm.device, **hints_m)
        else:
            torch.cuda.current_stream().wait_stream(side)
            for t in (atm, dm.idx, dm.mask, dm.batch, dm.ptr):
                t.record_stream(torch.cuda.current_stream())
        weights = None
        if self.cross_attn_module is not None:
            for layer in self.cross_attn_module.cross_attn_layers:
                layer.side_stream = side
            res, atm, weights = self.cross_attn_module.forward_packed(res, atm, dp, dm, self.return_attention)
        pe, me = self._pool(dp.pad(res), dp.mask), self._pool(dm.pad(atm), dm.mask)
        if self.include_post_pool_layernorm:
            pe, me = self.protein_post_pool_norm(pe), self.molecule_post_pool_norm(me)
        pe = self._stack(ops.dropout(self.dropout, self.activation(pe), "pool.protein"), self.protein_lins, self.protein_norms,
                         "protein_lins")
        me = self._stack(ops.dropout(self.dropout, self.activation(me), "pool.molecule"), self.molecule_lins, self.molecule_norms,
                         "molecule_lins")
        x = ops.dropout(self.dropout, self.activation(self.pm_embed_lin(torch.cat([pe, me], -1))), "pm_embed")
        x = self._stack(x, self.out_fc_layers, self.out_fc_norms, "out_fc")
        return self.output_layer(x), weights


def load_state_dict_from_checkpoint(model, state_dict, strict=True):
    """`inference/inference_utils.py:54-67`: checkpoints are saved from the torch.compile'd module, so every key
    carries an `_orig_mod.` prefix."""
    clean = {(k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k): v for k, v in state_dict.items()}
    return model.load_state_dict(clean, strict=strict)
