"""CPU oracle (TEST INFRASTRUCTURE) -- functional torch restatement of the reference GVP stack.

Every function takes a flat parameter dict `p` whose keys are the reference `state_dict` keys below a
`prefix` (e.g. ``conv_list.0.conv.message_func.0.ws.weight``), so the shipped checkpoint can be fed in
directly.  All maths is plain eager torch on whatever dtype the inputs have (fp32 for the timed CPU
baseline, fp64 for parity/gradient checks).  Citations are to `/root/reference` (read-only).

Conventions: ``s:[n,S]``, ``V:[n,C,3]`` (xyz innermost); ``edge_index:[2,E]`` with row 0 = source j,
row 1 = target i; messages are aggregated at the target (PyG ``source_to_target`` flow).
"""
import torch
import torch.nn.functional as F

EPS = 1e-8


def act(name, x):
    """Activation by name (the reference passes callables; `models/protein_gnn.py:341-349`)."""
    if name is None:
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(name)


def norm_no_nan(x, axis=-1, keepdims=False, sqrt=True):
    """`models/gvp_layers.py:79-86`: clamp the SQUARED sum at 1e-8, then (optionally) sqrt."""
    q = torch.clamp(torch.sum(x * x, dim=axis, keepdim=keepdims), min=EPS)
    return torch.sqrt(q) if sqrt else q


def gvp(p, prefix, x, scalar_act="relu", vector_act="sigmoid", vector_gate=False, vo_if_scalar_in=0):
    """One Geometric Vector Perceptron, `models/gvp_layers.py:142-175`.

    Dimensions are inferred from the weights: vi>0 iff ``wh.weight`` exists, vo>0 iff ``wv.weight`` exists.
    `vo_if_scalar_in` is the vo of the vi==0 branch (which has no `wv` to infer it from).
    """
    has_vi = (prefix + "wh.weight") in p
    has_vo = (prefix + "wv.weight") in p
    if has_vi:
        s, v = x
        vt = v.transpose(-1, -2)                                   # [n,3,vi]          :151
        vh = vt @ p[prefix + "wh.weight"].t()                      # [n,3,h]           :152
        vn = norm_no_nan(vh, axis=-2)                              # [n,h]             :153
        s = F.linear(torch.cat([s, vn], -1), p[prefix + "ws.weight"], p[prefix + "ws.bias"])   # :154
        if has_vo:
            vo = (vh @ p[prefix + "wv.weight"].t()).transpose(-1, -2)          # [n,vo,3]  :156-157
            if vector_gate:
                g_in = act(vector_act, s) if vector_act else s                  # :159-162
                gate = F.linear(g_in, p[prefix + "wsv.weight"], p[prefix + "wsv.bias"])
                vo = vo * torch.sigmoid(gate).unsqueeze(-1)                     # :163
            elif vector_act:
                vo = vo * act(vector_act, norm_no_nan(vo, axis=-1, keepdims=True))   # :164-166
    else:
        s = F.linear(x, p[prefix + "ws.weight"], p[prefix + "ws.bias"])         # :168
        if vo_if_scalar_in:                                                     # :169-171 (zeros)
            has_vo = True
            vo = s.new_zeros(s.shape[0], vo_if_scalar_in, 3)
    if scalar_act:
        s = act(scalar_act, s)                                                  # :172-173
    return (s, vo) if has_vo else s


def layer_norm(p, prefix, x, eps=1e-5):
    """`models/gvp_layers.py:231-242`: affine LayerNorm on s; vectors divided by the RMS (over channels)
    of their clamped squared norms.  A bare tensor means "scalar channels only" (:237-238)."""
    w, b = p[prefix + "scalar_norm.weight"], p[prefix + "scalar_norm.bias"]
    if torch.is_tensor(x):
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    s, v = x
    vn = norm_no_nan(v, axis=-1, keepdims=True, sqrt=False)        # [n,C,1]           :240
    vn = torch.sqrt(torch.mean(vn, dim=-2, keepdim=True))          # [n,1,1]           :241
    return F.layer_norm(s, (s.shape[-1],), w, b, eps), v / vn


def dropout(x, masks):
    """`models/gvp_layers.py:187-219` with EXTERNALLY supplied keep-masks (already scaled by 1/(1-p)):
    ``masks = (ms [n,S], mv [n,C])``; `None` = eval mode / identity."""
    if masks is None:
        return x
    ms, mv = masks
    s, v = x
    return s * ms, v * mv.unsqueeze(-1)


def aggregate(msg, dst, n, aggr):
    """PyG `MessagePassing` aggregation used by `GVPConv` (`models/gvp_layers.py:267,298`):
    sum at the target with dim_size=N; 'mean' divides by max(in_degree,1)."""
    out = msg.new_zeros((n, msg.shape[1])).index_add_(0, dst, msg)
    if aggr == "mean":
        cnt = torch.bincount(dst, minlength=n).clamp(min=1).to(msg.dtype)
        out = out / cnt.unsqueeze(-1)
    elif aggr not in ("add", "sum"):
        raise ValueError(aggr)
    return out


def gvp_conv(p, prefix, x, edge_index, edge_attr, aggr="mean", n_layers=3,
             scalar_act="relu", vector_act="sigmoid", vector_gate=False):
    """`GVPConv.forward/message`, `models/gvp_layers.py:291-308` (+ constructor :275-289 for which GVPs
    carry activations)."""
    s, v = x
    n, nv = s.shape[0], v.shape[1]
    src, dst = edge_index[0], edge_index[1]
    es, ev = edge_attr
    ms = torch.cat([s.index_select(0, src), es, s.index_select(0, dst)], -1)      # (s_j, e_s, s_i)  :306
    mv = torch.cat([v.index_select(0, src), ev, v.index_select(0, dst)], -2)      # (V_j, e_V, V_i)
    m = (ms, mv)
    for l in range(n_layers):
        last = l == n_layers - 1
        m = gvp(p, f"{prefix}message_func.{l}.", m,
                None if last else scalar_act, None if last else vector_act, vector_gate)   # :277-288
    ms, mv = m
    merged = torch.cat([ms, mv.reshape(mv.shape[0], 3 * nv)], -1)                 # _merge :101-109
    out = aggregate(merged, dst, n, aggr)
    return out[:, : -3 * nv], out[:, -3 * nv:].reshape(n, nv, 3)                  # _split :88-99


def feed_forward(p, prefix, x, n_feedforward=2, scalar_act="relu", vector_act="sigmoid", vector_gate=False):
    """`ff_func` of `GVPConvLayer`, `models/gvp_layers.py:355-364`."""
    for l in range(n_feedforward):
        last = l == n_feedforward - 1
        x = gvp(p, f"{prefix}ff_func.{l}.", x,
                None if last else scalar_act, None if last else vector_act, vector_gate)
    return x


def gvp_conv_layer(p, prefix, x, edge_index, edge_attr, aggr="mean", n_message=3, n_feedforward=2,
                   scalar_act="relu", vector_act="sigmoid", vector_gate=False,
                   drop_masks=(None, None), autoregressive_x=None, node_mask=None):
    """`GVPConvLayer.forward`, `models/gvp_layers.py:366-414`."""
    kw = dict(scalar_act=scalar_act, vector_act=vector_act, vector_gate=vector_gate)
    if autoregressive_x is not None:                                              # :382-398
        src, dst = edge_index
        fwd = src < dst
        ei_f, ei_b = edge_index[:, fwd], edge_index[:, ~fwd]
        ea_f = (edge_attr[0][fwd], edge_attr[1][fwd])
        ea_b = (edge_attr[0][~fwd], edge_attr[1][~fwd])
        a = gvp_conv(p, prefix + "conv.", x, ei_f, ea_f, "add", n_message, **kw)
        b = gvp_conv(p, prefix + "conv.", autoregressive_x, ei_b, ea_b, "add", n_message, **kw)
        cnt = torch.bincount(dst, minlength=x[0].shape[0]).clamp(min=1).to(x[0].dtype)
        dh = ((a[0] + b[0]) / cnt.unsqueeze(-1), (a[1] + b[1]) / cnt.view(-1, 1, 1))
    else:
        dh = gvp_conv(p, prefix + "conv.", x, edge_index, edge_attr, aggr, n_message, **kw)   # :401
    full = x
    if node_mask is not None:                                                     # :403-405
        x = (x[0][node_mask], x[1][node_mask])
        dh = (dh[0][node_mask], dh[1][node_mask])
    dh = dropout(dh, drop_masks[0])
    x = layer_norm(p, prefix + "norm.0.", (x[0] + dh[0], x[1] + dh[1]))           # :407
    dh = dropout(feed_forward(p, prefix, x, n_feedforward, **kw), drop_masks[1])  # :409
    x = layer_norm(p, prefix + "norm.1.", (x[0] + dh[0], x[1] + dh[1]))           # :410
    if node_mask is not None:                                                     # :412-414
        s_all, v_all = full[0].clone(), full[1].clone()
        s_all[node_mask], v_all[node_mask] = x[0], x[1]
        x = (s_all, v_all)
    return x


def lba_encoder(p, prefix, x, edge_index, ntypes, etypes, eattr, num_ntypes, num_etypes,
                num_convs=2, aggr="sum", drop_masks=None, return_hidden=False):
    """`VectorProteinGNN_LBAModel.forward`, `models/protein_gnn.py:361-388` (one-hot type embedding,
    `:139-152`; layer construction `:325-358`).  `drop_masks[k] = (masks0, masks1)` per conv layer."""
    xs, xv = x
    es, ev = eattr
    xs = torch.cat([F.one_hot(ntypes, num_ntypes).to(xs.dtype), xs], -1)          # one-hot FIRST :145-146
    es = torch.cat([F.one_hot(etypes, num_etypes).to(es.dtype), es], -1)          # :149-150
    h = layer_norm(p, prefix + "gvp_node.1.", gvp(p, prefix + "gvp_node.0.", (xs, xv), None, None, True))
    e = layer_norm(p, prefix + "gvp_edge.1.", gvp(p, prefix + "gvp_edge.0.", (es, ev), None, None, True))
    for k in range(num_convs):                                                    # :379-380
        h = gvp_conv_layer(p, f"{prefix}conv_list.{k}.", h, edge_index, e, aggr=aggr,
                           scalar_act="relu", vector_act=None, vector_gate=True,
                           drop_masks=(None, None) if drop_masks is None else drop_masks[k])
    hn = layer_norm(p, prefix + "gvp_norm_before_scalar.", h)                     # :385
    out = gvp(p, prefix + "gvp_to_scalar.", hn, "relu", None, True)               # :386 (vo=0 -> tensor)
    return (out, h, e) if return_hidden else out


# --------------------------------------------------------------------------------------------------
# helpers for tests / baselines
# --------------------------------------------------------------------------------------------------

def init_gvp_params(p, prefix, in_dims, out_dims, h_dim=None, vector_gate=False, gen=None, dtype=torch.float32):
    """Random parameters with `nn.Linear`'s default init bounds and the reference key names
    (`models/gvp_layers.py:129-140`)."""
    si, vi = in_dims
    so, vo = out_dims

    def lin(name, fan_out, fan_in, bias):
        bound = 1.0 / max(fan_in, 1) ** 0.5
        p[prefix + name + ".weight"] = (torch.rand(fan_out, fan_in, generator=gen, dtype=dtype) * 2 - 1) * bound
        if bias:
            p[prefix + name + ".bias"] = (torch.rand(fan_out, generator=gen, dtype=dtype) * 2 - 1) * bound

    if vi:
        h = h_dim or max(vi, vo)
        lin("wh", h, vi, False)
        lin("ws", so, h + si, True)
        if vo:
            lin("wv", vo, h, False)
            if vector_gate:
                lin("wsv", vo, so, True)
    else:
        lin("ws", so, si, True)
    p[prefix + "dummy_param"] = torch.empty(0, dtype=dtype)
    return p


def init_layer_norm_params(p, prefix, ns, gen=None, dtype=torch.float32):
    p[prefix + "scalar_norm.weight"] = 1 + 0.1 * torch.randn(ns, generator=gen, dtype=dtype)
    p[prefix + "scalar_norm.bias"] = 0.1 * torch.randn(ns, generator=gen, dtype=dtype)
    return p


def init_conv_layer_params(p, prefix, node_dims, edge_dims, n_message=3, n_feedforward=2,
                           vector_gate=True, gen=None, dtype=torch.float32):
    """Key layout of `GVPConvLayer` (`models/gvp_layers.py:347-364`)."""
    ns, nv = node_dims
    es, ev = edge_dims
    kw = dict(vector_gate=vector_gate, gen=gen, dtype=dtype)
    if n_message == 1:
        init_gvp_params(p, prefix + "conv.message_func.0.", (2 * ns + es, 2 * nv + ev), node_dims, **kw)
    else:
        init_gvp_params(p, prefix + "conv.message_func.0.", (2 * ns + es, 2 * nv + ev), node_dims, **kw)
        for l in range(1, n_message):
            init_gvp_params(p, f"{prefix}conv.message_func.{l}.", node_dims, node_dims, **kw)
    for k in range(2):
        init_layer_norm_params(p, f"{prefix}norm.{k}.", ns, gen, dtype)
        p[f"{prefix}dropout.{k}.vdropout.dummy_param"] = torch.empty(0, dtype=dtype)
    if n_feedforward == 1:
        init_gvp_params(p, prefix + "ff_func.0.", node_dims, node_dims, **kw)
    else:
        hid = (4 * ns, 2 * nv)
        init_gvp_params(p, prefix + "ff_func.0.", node_dims, hid, **kw)
        for l in range(1, n_feedforward - 1):
            init_gvp_params(p, f"{prefix}ff_func.{l}.", hid, hid, **kw)
        init_gvp_params(p, f"{prefix}ff_func.{n_feedforward - 1}.", hid, node_dims, **kw)
    return p


def init_lba_params(prefix, in_channels, edge_dim, num_ntypes, num_etypes, hidden, edge_hidden, out_channels,
                    num_convs, gen=None, dtype=torch.float32):
    """Key layout of `VectorProteinGNN_LBAModel` (`models/protein_gnn.py:321-358`)."""
    p = {}
    kw = dict(vector_gate=True, gen=gen, dtype=dtype)
    init_gvp_params(p, prefix + "gvp_node.0.", (in_channels[0] + num_ntypes, in_channels[1]), hidden, **kw)
    init_layer_norm_params(p, prefix + "gvp_node.1.", hidden[0], gen, dtype)
    init_gvp_params(p, prefix + "gvp_edge.0.", (edge_dim[0] + num_etypes, edge_dim[1]), edge_hidden, **kw)
    init_layer_norm_params(p, prefix + "gvp_edge.1.", edge_hidden[0], gen, dtype)
    for k in range(num_convs):
        init_conv_layer_params(p, f"{prefix}conv_list.{k}.", hidden, edge_hidden, gen=gen, dtype=dtype)
    init_layer_norm_params(p, prefix + "gvp_norm_before_scalar.", hidden[0], gen, dtype)
    init_gvp_params(p, prefix + "gvp_to_scalar.", hidden, (out_channels, 0), **kw)
    return p
