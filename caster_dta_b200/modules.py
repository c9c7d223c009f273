"""Drop-in GVP operator library backed by the sm_100a kernels.

Mirrors the public surface of the reference `models/gvp_layers.py` -- same constructors, same `forward`
signatures, same `(s, V)` tuple convention and IDENTICAL `state_dict` keys (including the zero-size
`dummy_param`s) -- so `model_kwargs.json` and the shipped checkpoint load unchanged.  All arithmetic runs in
`libcastergvp.so`; there is no eager-PyTorch path (CPU tensors raise).

    GVP            <- models/gvp_layers.py:111-175        fused row program, one launch
    LayerNorm      <- :221-242                            fused row program
    Dropout        <- :177-219                            (mask generation only; masks are consumed in-kernel)
    GVPConv        <- :244-308 (+ PyG propagate)          fused gather -> message GVPs -> segmented reduce
    GVPConvLayer   <- :311-414                            conv + ONE fused node-update launch
"""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, MAX_CHAIN


# ---- tuple helpers (models/gvp_layers.py:39-109) ----------------------------------------------------------------
def tuple_sum(*args):
    """Elementwise sum of (s, V) tuples."""
    s = sum(a[0] for a in args)
    v = sum(a[1] for a in args)
    return s, v


def tuple_cat(*args, dim=-1):
    """Concatenate (s, V) tuples; `dim` counts on the scalar tensor (the vector tensor has one more axis)."""
    nd = args[0][0].dim()
    d = dim % nd
    return torch.cat([a[0] for a in args], dim=d), torch.cat([a[1] for a in args], dim=d)


def tuple_index(x, idx):
    return x[0][idx], x[1][idx]


def randn(n, dims, device="cpu"):
    return torch.randn(n, dims[0], device=device), torch.randn(n, dims[1], 3, device=device)


def _norm_no_nan(x, axis=-1, keepdims=False, eps=1e-8, sqrt=True):
    out = torch.clamp(torch.sum(torch.square(x), axis, keepdims), min=eps)
    return torch.sqrt(out) if sqrt else out


def _split(x, nv):
    s = x[..., : x.shape[-1] - 3 * nv]
    v = x[..., x.shape[-1] - 3 * nv:].reshape(x.shape[:-1] + (nv, 3))
    return s, v


def _merge(s, v):
    return torch.cat([s, v.reshape(v.shape[:-2] + (3 * v.shape[-2],))], -1)


# ---- activation recognition ---------------------------------------------------------------------------------------
def _act_code(fn):
    """The reference passes callables (`F.relu`, `torch.sigmoid`, `nn.ReLU()`); the kernels know them by code.
    Anything else is rejected here, at construction time, rather than silently computed elsewhere."""
    if fn is None:
        return ACT_NONE
    if fn in (F.relu, torch.relu) or isinstance(fn, nn.ReLU):
        return ACT_RELU
    if fn in (torch.sigmoid, F.sigmoid) or isinstance(fn, nn.Sigmoid):
        return ACT_SIGMOID
    raise ValueError(f"unsupported GVP activation {fn!r}: the fused kernels implement None, ReLU and sigmoid")


def _unpack(x):
    """PyG collate turns tuples into lists (reference guards for it in explanation/explain_wrapper.py:123-127)."""
    if torch.is_tensor(x):
        return x, None
    return x[0], x[1]


class GVP(nn.Module):
    """Geometric Vector Perceptron (`models/gvp_layers.py:111-175`)."""

    def __init__(self, in_dims, out_dims, h_dim=None, activations=(F.relu, torch.sigmoid), vector_gate=False):
        super().__init__()
        self.si, self.vi = in_dims
        self.so, self.vo = out_dims
        self.vector_gate = vector_gate
        if self.vi:
            self.h_dim = h_dim or max(self.vi, self.vo)
            self.wh = nn.Linear(self.vi, self.h_dim, bias=False)
            self.ws = nn.Linear(self.h_dim + self.si, self.so)
            if self.vo:
                self.wv = nn.Linear(self.h_dim, self.vo, bias=False)
                if self.vector_gate:
                    self.wsv = nn.Linear(self.so, self.vo)
        else:
            self.h_dim = 0
            self.ws = nn.Linear(self.si, self.so)
        self.scalar_act, self.vector_act = activations
        self.dummy_param = nn.Parameter(torch.empty(0))
        self.spec = ops.GvpSpec(self.si, self.vi, self.so, self.vo, self.h_dim, _act_code(self.scalar_act),
                                _act_code(self.vector_act), vector_gate)

    def kernel_weights(self):
        """(wh, ws.weight, ws.bias, wv, wsv.weight, wsv.bias) with None where the GVP has no such layer."""
        sp = self.spec
        return (self.wh.weight if sp.has_wh else None, self.ws.weight, self.ws.bias,
                self.wv.weight if sp.has_wv else None,
                self.wsv.weight if sp.has_gate else None, self.wsv.bias if sp.has_gate else None)

    def forward(self, x):
        s, v = _unpack(x)
        if self.vi and v is None:
            raise ValueError("GVP with vector inputs expects a (s, V) tuple")
        prog = _row_program(self.si, self.vi, (self.spec,))
        out_s, out_v = ops.run_rows(prog, s, v if self.vi else None, weights=self.kernel_weights())
        return (out_s, out_v) if self.vo else out_s


@functools.lru_cache(maxsize=256)
def _row_program_cached(in_s, in_v, keys, onehot, residual_in, pre_norm, post_residual, post_norm):
    gvps = [ops.GvpSpec(*k) for k in keys]
    return ops.RowProgram(in_s, in_v, gvps, onehot, residual_in, pre_norm, post_residual, post_norm)


def _row_program(in_s, in_v, specs, onehot=0, residual_in=False, pre_norm=False, post_residual=False, post_norm=False):
    return _row_program_cached(int(in_s), int(in_v), tuple(sp.key() for sp in specs), int(onehot), bool(residual_in),
                               bool(pre_norm), bool(post_residual), bool(post_norm))


class _VDropout(nn.Module):
    """Vector-channel dropout: all three components of a channel are dropped together (`:177-198`)."""

    def __init__(self, drop_rate):
        super().__init__()
        self.drop_rate = drop_rate
        self.dummy_param = nn.Parameter(torch.empty(0))

    def mask(self, shape, device):
        keep = 1 - self.drop_rate
        return torch.bernoulli(keep * torch.ones(shape, device=device)) / keep

    def forward(self, x):
        if not self.training:
            return x
        return self.mask(x.shape[:-1], x.device).unsqueeze(-1) * x


class Dropout(nn.Module):
    """Combined dropout for (s, V) (`:200-219`).  `masks()` draws the two keep-masks in the reference's RNG order
    so that the fused node-update kernel can apply them."""

    def __init__(self, drop_rate):
        super().__init__()
        self.sdropout = nn.Dropout(drop_rate)
        self.vdropout = _VDropout(drop_rate)

    def masks(self, s_shape, v_shape, device):
        if not self.training or self.sdropout.p == 0:
            return None
        ms = self.sdropout(torch.ones(s_shape, device=device))
        mv = self.vdropout.mask(v_shape, device)
        if ops.MASK_LOG is not None:
            ops.MASK_LOG.setdefault("gvp", []).append((ms, mv))
        return ms, mv

    def forward(self, x):
        if torch.is_tensor(x):
            return self.sdropout(x)
        s, v = _unpack(x)
        return self.sdropout(s), self.vdropout(v)


class LayerNorm(nn.Module):
    """Combined LayerNorm for (s, V) (`:221-242`)."""

    def __init__(self, dims):
        super().__init__()
        self.s, self.v = dims
        self.scalar_norm = nn.LayerNorm(self.s)

    def forward(self, x):
        s, v = _unpack(x)
        prog = _row_program(self.s, self.v if v is not None else 0, (), pre_norm=True)
        out_s, out_v = ops.run_rows(prog, s, v if self.v else None, ln0=(self.scalar_norm.weight, self.scalar_norm.bias))
        return (out_s, out_v) if (self.v and v is not None) else out_s


class GVPConv(nn.Module):
    """Graph convolution / message passing with GVPs (`:244-308`); aggregation is a deterministic segmented
    reduction over dst-sorted edges instead of PyG's atomic `scatter_add_`."""

    def __init__(self, in_dims, out_dims, edge_dims, n_layers=3, module_list=None, aggr="mean",
                 activations=(F.relu, torch.sigmoid), vector_gate=False):
        super().__init__()
        if aggr not in ("mean", "add", "sum"):
            raise ValueError(f"aggr={aggr!r}: the fused kernel implements 'mean', 'add' and 'sum'")
        self.aggr = aggr
        self.explain = False          # attribute PyG's MessagePassing carries; explain_wrapper.py:66-68 toggles it
        self.si, self.vi = in_dims
        self.so, self.vo = out_dims
        self.se, self.ve = edge_dims
        GVP_ = functools.partial(GVP, activations=activations, vector_gate=vector_gate)
        module_list = module_list or []
        if not module_list:
            msg_in = (2 * self.si + self.se, 2 * self.vi + self.ve)
            if n_layers == 1:
                module_list.append(GVP_(msg_in, (self.so, self.vo), activations=(None, None)))
            else:
                module_list.append(GVP_(msg_in, out_dims))
                for _ in range(n_layers - 2):
                    module_list.append(GVP_(out_dims, out_dims))
                module_list.append(GVP_(out_dims, out_dims, activations=(None, None)))
        if len(module_list) > MAX_CHAIN or not all(isinstance(m, GVP) for m in module_list):
            raise ValueError(f"message_func must be 1..{MAX_CHAIN} GVP modules")
        self.message_func = nn.Sequential(*module_list)

    def _program(self, edge_sorted=False):
        return _conv_program(self.si, self.vi, self.se, self.ve, tuple(m.spec.key() for m in self.message_func),
                             "mean" if self.aggr == "mean" else "sum", bool(edge_sorted))

    def kernel_weights(self):
        w = []
        for m in self.message_func:
            w.extend(m.kernel_weights())
        return w

    def forward(self, x, edge_index, edge_attr, plan=None, edge_sorted=False):
        x_s, x_v = _unpack(x)
        e_s, e_v = _unpack(edge_attr)
        if plan is None:
            plan = ops.get_plan(edge_index, x_s.shape[0])
        return ops.run_conv(self._program(edge_sorted), plan, (x_s, x_v), (e_s, e_v), self.kernel_weights())


@functools.lru_cache(maxsize=256)
def _conv_program(ns, nv, es, ev, keys, aggr, edge_sorted):
    return ops.ConvProgram(ns, nv, es, ev, [ops.GvpSpec(*k) for k in keys], aggr, edge_sorted)


class GVPConvLayer(nn.Module):
    """Full GVP message-passing layer (`:311-414`): conv, residual + LayerNorm, feed-forward, residual + LayerNorm.
    Everything after the conv is ONE kernel launch."""

    def __init__(self, node_dims, edge_dims, n_message=3, n_feedforward=2, drop_rate=.1, autoregressive=False,
                 activations=(F.relu, torch.sigmoid), vector_gate=False, aggr=None):
        super().__init__()
        if autoregressive:
            if aggr is not None and aggr != "add":
                raise ValueError("Cannot use autoregressive and aggr together in GVPConvLayer unless aggr is set to 'add'")
            aggr = "add"
        elif aggr is None:
            aggr = "mean"
        self.node_dims = tuple(node_dims)
        self.conv = GVPConv(node_dims, node_dims, edge_dims, n_message, aggr=aggr, activations=activations,
                            vector_gate=vector_gate)
        GVP_ = functools.partial(GVP, activations=activations, vector_gate=vector_gate)
        self.norm = nn.ModuleList([LayerNorm(node_dims) for _ in range(2)])
        self.dropout = nn.ModuleList([Dropout(drop_rate) for _ in range(2)])
        ff = []
        if n_feedforward == 1:
            ff.append(GVP_(node_dims, node_dims, activations=(None, None)))
        else:
            hid = 4 * node_dims[0], 2 * node_dims[1]
            ff.append(GVP_(node_dims, hid))
            for _ in range(n_feedforward - 2):
                ff.append(GVP_(hid, hid))
            ff.append(GVP_(hid, node_dims, activations=(None, None)))
        if len(ff) > MAX_CHAIN:
            raise ValueError(f"n_feedforward > {MAX_CHAIN} is not supported by the fused node-update kernel")
        self.ff_func = nn.Sequential(*ff)

    def node_update(self, x, dh):
        """x <- LN1(x1 + D1(FF(x1))),  x1 = LN0(x + D0(dh))   (`:407-410`) in one launch."""
        ns, nv = self.node_dims
        n, dev = x[0].shape[0], x[0].device
        m0 = self.dropout[0].masks((n, ns), (n, nv), dev)
        m1 = self.dropout[1].masks((n, ns), (n, nv), dev)
        prog = _row_program(ns, nv, tuple(m.spec for m in self.ff_func), residual_in=True, pre_norm=True,
                            post_residual=True, post_norm=True)
        w = []
        for m in self.ff_func:
            w.extend(m.kernel_weights())
        ln0, ln1 = self.norm[0].scalar_norm, self.norm[1].scalar_norm
        return ops.run_rows(prog, x[0], x[1] if nv else None, h=dh, masks0=m0, masks1=m1, ln0=(ln0.weight, ln0.bias),
                            ln1=(ln1.weight, ln1.bias), weights=w)

    def forward(self, x, edge_index, edge_attr, autoregressive_x=None, node_mask=None, plan=None, edge_sorted=False):
        x = _unpack(x)
        edge_attr = _unpack(edge_attr)
        if autoregressive_x is not None:                                   # `:382-398`
            src, dst = edge_index
            fwd = src < dst
            n = x[0].shape[0]
            a = self.conv(x, edge_index[:, fwd].contiguous(), tuple_index(edge_attr, fwd))
            b = self.conv(_unpack(autoregressive_x), edge_index[:, ~fwd].contiguous(), tuple_index(edge_attr, ~fwd))
            count = torch.bincount(dst, minlength=n).clamp(min=1).to(a[0].dtype)
            dh = (a[0] + b[0]) / count.unsqueeze(-1), (a[1] + b[1]) / count.view(-1, 1, 1)
        else:
            dh = self.conv(x, edge_index, edge_attr, plan=plan, edge_sorted=edge_sorted)
        if node_mask is not None:                                          # `:403-414` (out of place)
            sub = self.node_update(tuple_index(x, node_mask), tuple_index(dh, node_mask))
            s_all, v_all = x[0].clone(), x[1].clone()
            s_all[node_mask], v_all[node_mask] = sub[0], sub[1]
            return s_all, v_all
        return self.node_update(x, dh)
