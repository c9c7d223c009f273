"""CPU oracle (TEST INFRASTRUCTURE) -- functional restatement of the non-GVP remainder of CASTER-DTA.

Restates `models/joint_gnn.py:172-288` (JointGNN.forward), `:389-408` (CrossAttentionModule.forward) and the
GINE molecule encoder `models/molecule_gnn.py:254-280` (with PyG `GINEConv` / `MLP` semantics: out =
MLP((1+eps) x_i + sum_j relu(x_j + lin(e_ji))), MLP = lins.0 -> act -> lins.1) over a flat state_dict `p`
with the reference key names.  Exists so that the whole model's pairs/s can be timed on host cores as
`cpu_baseline` and so the product's end-to-end affinity can be checked; the GVP encoder itself lives in
`gvp_oracle.lba_encoder`.
"""
import math

import torch
import torch.nn.functional as F

from . import gvp_oracle


def _act(name, x):
    if name == "leaky_relu":
        return F.leaky_relu(x, 0.01)
    if name == "relu":
        return torch.relu(x)
    raise ValueError(name)


def _drop(x, prob, training, masks=None, site=None, dense=None):
    """Inverted dropout.  With `masks` (site name -> keep-mask already scaled by 1/(1-p), as recorded from the CUDA
    path) the mask is applied instead of drawing one; `dense = (batch, pos)` converts a mask recorded on packed rows
    [total, D] to this function's padded layout [B, L, D] (padding rows are masked out downstream)."""
    if masks is not None:
        if not training or prob == 0 or site not in masks:
            return x
        m = masks[site].to(x.dtype)
        if dense is not None and m.dim() == 2 and x.dim() == 3:
            batch, pos = dense
            full = torch.ones_like(x)
            full[batch, pos] = m
            m = full
        return x * m
    return F.dropout(x, prob, True) if (training and prob > 0) else x


def gine_encoder(p, prefix, kw, x, edge_index, ntypes, etypes, eattr, training=False, drop_masks=None):
    """`HomoMoleculeGNN_GINE.forward` (`models/molecule_gnn.py:254-268`)."""
    x = torch.cat([F.one_hot(ntypes, kw["num_ntypes"]).to(x.dtype), x], -1)
    e = torch.cat([F.one_hot(etypes, kw["num_etypes"]).to(x.dtype), eattr], -1)
    src, dst = edge_index[0], edge_index[1]
    nconv = kw["num_convs"]
    for k in range(nconv):
        q = f"{prefix}conv_list.{k}."
        el = F.linear(e, p[q + "lin.weight"], p[q + "lin.bias"])
        agg = torch.zeros_like(x).index_add_(0, dst, torch.relu(x.index_select(0, src) + el))
        h = agg + (1 + p[q + "eps"]) * x
        h = _act(kw["activation"], F.linear(h, p[q + "nn.lins.0.weight"], p[q + "nn.lins.0.bias"]))
        x = F.linear(h, p[q + "nn.lins.1.weight"], p[q + "nn.lins.1.bias"])
        x = _act(kw["activation"], x)
        if k < nconv - 1:
            x = _drop(x, kw["dropout_rate"], training, drop_masks, f"gine.{k}")
    return x


def to_dense_batch(x, batch, b):
    """PyG `to_dense_batch` (`models/joint_gnn.py:206-207`)."""
    counts = torch.bincount(batch, minlength=b)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    m = int(counts.max())
    pos = torch.arange(x.shape[0], device=x.device) - ptr[batch]
    out = x.new_zeros((b, m, x.shape[1]))
    out[batch, pos] = x
    mask = torch.zeros(b, m, dtype=torch.bool, device=x.device)
    mask[batch, pos] = True
    return out, mask


def mha(p, prefix, q_in, kv_in, key_real_mask, heads):
    """`nn.MultiheadAttention(batch_first=True)` with kdim == vdim == embed_dim, need_weights=True
    (head-averaged weights), `key_padding_mask = ~mask` (`models/joint_gnn.py:393-394`)."""
    d = q_in.shape[-1]
    w, bias = p[prefix + "in_proj_weight"], p[prefix + "in_proj_bias"]
    q = F.linear(q_in, w[:d], bias[:d])
    k = F.linear(kv_in, w[d:2 * d], bias[d:2 * d])
    v = F.linear(kv_in, w[2 * d:], bias[2 * d:])
    b, lq, lk, hd = q.shape[0], q.shape[1], k.shape[1], d // heads
    q = q.view(b, lq, heads, hd).transpose(1, 2)
    k = k.view(b, lk, heads, hd).transpose(1, 2)
    v = v.view(b, lk, heads, hd).transpose(1, 2)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    att = att.masked_fill(~key_real_mask[:, None, None, :], float("-inf"))
    att = torch.softmax(att, -1)
    o = (att @ v).transpose(1, 2).reshape(b, lq, d)
    return F.linear(o, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"]), att.mean(1)


def cross_attention(p, prefix, e1, e2, m1, m2, heads, drop, training, drop_masks=None, site="attn.0", dense1=None,
                    dense2=None):
    """`CrossAttentionModule.forward` with the residual stream (`models/joint_gnn.py:389-421`)."""
    ln = lambda name, x: F.layer_norm(x, (x.shape[-1],), p[prefix + name + ".weight"], p[prefix + name + ".bias"])
    x1n, x2n = ln("preattn_norm1", e1), ln("preattn_norm2", e2)
    a1, w1 = mha(p, prefix + "embed1_to_2.", x1n, x2n, m2, heads)
    a2, w2 = mha(p, prefix + "embed2_to_1.", x2n, x1n, m1, heads)

    def ff(name, x, dense):
        h = torch.relu(F.linear(x, p[f"{prefix}{name}.0.weight"], p[f"{prefix}{name}.0.bias"]))
        h = _drop(h, drop, training, drop_masks, f"{site}.{name}.inner", dense)
        return F.linear(h, p[f"{prefix}{name}.3.weight"], p[f"{prefix}{name}.3.bias"])

    e1 = e1 + _drop(a1, drop, training, drop_masks, site + ".a1", dense1)
    e1 = e1 + _drop(ff("ff1", ln("ff_norm1", e1), dense1), drop, training, drop_masks, site + ".ff1.outer", dense1)
    e2 = e2 + _drop(a2, drop, training, drop_masks, site + ".a2", dense2)
    e2 = e2 + _drop(ff("ff2", ln("ff_norm2", e2), dense2), drop, training, drop_masks, site + ".ff2.outer", dense2)
    return e1, e2, (w1, w2)


def _dense_index(batch, b):
    counts = torch.bincount(batch, minlength=b)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    return batch, torch.arange(batch.shape[0], device=batch.device) - ptr[batch]


def joint_forward(p, kwargs, prot, mol, training=False, protein_embed=None, drop_masks=None, num_graphs=None):
    """`JointGNN.forward` (`models/joint_gnn.py:172-288`) for the shipped configuration
    (`model_kwargs.json`: depth-1 linear stacks, one cross-attention layer, mean pooling, no norms).

    `prot` / `mol`: dicts with x, edge_index, ntypes, etypes, eattr, batch.  `protein_embed` lets a caller
    substitute residue embeddings computed elsewhere (to isolate the head).  `drop_masks`: site name -> keep-mask
    (see `_drop`), the masks the CUDA path recorded, so that a train-mode step can be compared value for value."""
    pk, mk, jk = kwargs["protein_gnn_kwargs"], kwargs["molecule_gnn_kwargs"], kwargs["joint_gnn_kwargs"]
    a, dr = jk["activation"], jk["dropout"]
    if protein_embed is None:
        protein_embed = gvp_oracle.lba_encoder(
            p, "protein_gnn.gnn_model.", prot["x"], prot["edge_index"], prot["ntypes"], prot["etypes"],
            prot["eattr"], pk["num_ntypes"], pk["num_etypes"], pk["num_convs"], pk["aggr"])
    res = protein_embed
    dm = drop_masks
    atm = gine_encoder(p, "molecule_gnn.gnn_model.", mk, mol["x"], mol["edge_index"], mol["ntypes"],
                       mol["etypes"], mol["eattr"], training, dm)
    for k in range(jk["residue_lin_depth"]):
        res = _drop(_act(a, F.linear(res, p[f"residue_lins.{k}.weight"], p[f"residue_lins.{k}.bias"])), dr, training, dm,
                    f"residue_lins.{k}")
    for k in range(jk["atom_lin_depth"]):
        atm = _drop(_act(a, F.linear(atm, p[f"atom_lins.{k}.weight"], p[f"atom_lins.{k}.bias"])), dr, training, dm,
                    f"atom_lins.{k}")
    b = int(prot["batch"].max()) + 1 if num_graphs is None else int(num_graphs)
    dense_r, dense_a = _dense_index(prot["batch"], b), _dense_index(mol["batch"], b)
    res, rmask = to_dense_batch(res, prot["batch"], b)
    atm, amask = to_dense_batch(atm, mol["batch"], b)
    weights = []
    for k in range(jk["num_cross_attn_layers"]):
        res, atm, w = cross_attention(p, f"cross_attn_module.cross_attn_layers.{k}.", res, atm, rmask, amask,
                                      jk["n_attention_heads"], dr, training, dm, f"attn.{k}", dense_r, dense_a)
        weights.append(w)
    assert jk["element_pooling"] == "mean" and not jk["include_post_pool_layernorm"]
    pe = (res * rmask.unsqueeze(-1)).sum(1) / rmask.sum(1, keepdim=True)
    me = (atm * amask.unsqueeze(-1)).sum(1) / amask.sum(1, keepdim=True)
    pe = _drop(_act(a, pe), dr, training, dm, "pool.protein")
    me = _drop(_act(a, me), dr, training, dm, "pool.molecule")
    for k in range(jk["protein_lin_depth"]):
        pe = _drop(_act(a, F.linear(pe, p[f"protein_lins.{k}.weight"], p[f"protein_lins.{k}.bias"])), dr, training, dm,
                   f"protein_lins.{k}")
    for k in range(jk["molecule_lin_depth"]):
        me = _drop(_act(a, F.linear(me, p[f"molecule_lins.{k}.weight"], p[f"molecule_lins.{k}.bias"])), dr, training, dm,
                   f"molecule_lins.{k}")
    x = torch.cat([pe, me], -1)
    x = _drop(_act(a, F.linear(x, p["pm_embed_lin.weight"], p["pm_embed_lin.bias"])), dr, training, dm, "pm_embed")
    for k in range(jk["out_lin_depth"]):
        x = _drop(_act(a, F.linear(x, p[f"out_fc_layers.{k}.weight"], p[f"out_fc_layers.{k}.bias"])), dr, training, dm,
                  f"out_fc.{k}")
    return F.linear(x, p["output_layer.weight"], p["output_layer.bias"]), weights
