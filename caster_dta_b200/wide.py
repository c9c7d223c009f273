"""GEMM formulation of the fused GVPConv and of the node-level row programs for WIDE feature dims
(BASELINE config 5: nodes (100,16), edges (32,1)) -- the training path of those dims.

At the checkpoint dims (16,4)/(32,1) the message GVPs are too small for anything but register-resident FFMA kernels
(csrc/conv_reg.cu).  At (100,16) every projection of `GVP.forward` (`models/gvp_layers.py:142-175`) is a real GEMM over
the edges, and the generic shared-memory tile kernels that serve "any dims" (csrc/conv.cu, csrc/rows.cu) are ~100x off
the pace in the backward pass.  This module restates `GVPConv.forward/message` + PyG `propagate`
(`models/gvp_layers.py:291-308`) and the node update of `GVPConvLayer.forward` (`:407-410`) as a short sequence of dense
GEMMs over CHUNKS of dst-sorted edges, with three structural choices the eager reference does not make:

  * **node-level split of message GVP 0.**  The message input is `[s_j ; e_s ; s_i]`, `[V_j ; e_V ; V_i]` (`:306`), so
    `W_s [s_j ; e_s ; s_i ; vn] + b = P_j[src] + P_i[dst] + W_e e_s + W_vn vn` and `W_h [V_j ; e_V ; V_i] = Q_j[src] +
    Q_i[dst] + W_he e_V`: the node blocks of W_s / W_h (200 of 265 and 32 of 33 reduction rows at config-5 dims) are
    applied once per NODE; per edge only the projected rows are gathered.  Forward FLOPs per edge 124 846 -> ~78 000.
    The backward mirrors it: `ds'_0` and `dVh_0` are reduced per node over the target / source CSR views first
    (`R_i`, `R_j`), and `d_x = R_i W_i + R_j W_j`, `dW_i = R_i^T x`, ... are node-level GEMMs (30x fewer rows);
  * **plane-major vectors.**  Vector features are kept as three planes `[3, rows, C]`, so every W_h / W_mu projection is
    ONE `[3 rows, C] x [C, H]` GEMM without the two transposes of `:151,157`, norms and gates are plane-wise
    element-wise ops;
  * **nothing of size E x (2ns + es) is materialised**, the aggregation and the two per-node reductions are the
    deterministic CSR segmented sums of `cgvp_segment_reduce` (no atomics), and chunking bounds the intermediates
    (~5 KB per edge of a chunk) independently of E.

The GEMMs themselves are plain library GEMMs (cuBLAS through `torch.matmul`, fp32; TF32 tensor cores when
`set_tensor_cores(True)` selects the <= 1e-2 mode); gathers / reductions go through the C ABI.  There is no CPU path:
`_segsum` raises for CPU tensors (the unit tests substitute it to check the algebra against the fp64 oracle).

Backward formulas: SURVEY.md Appendix E, the same as `csrc/cgvp_reg.cuh::gvp_bwd_ds`.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID

EPS = 1e-8        # clamp of the squared vector norms, models/gvp_layers.py:79-86
LN_EPS = 1e-5     # nn.LayerNorm default, models/gvp_layers.py:229

# Serve wide descriptors through this module (forward in fp32 mode, backward always).  CGVP_WIDE_GEMM=0 keeps the generic
# tile kernels (parity tests compare the two).
ENABLED = os.environ.get("CGVP_WIDE_GEMM", "1") == "1"
MIN_DIM = 64                  # scalar node channels from which the GEMM formulation takes over
CHUNK_EDGES = 1 << 18         # edges per chunk (intermediates: ~5 KB per edge at config-5 dims)
# Training: the forward may leave its per-chunk intermediates for the backward of the same call (no recompute) when all of
# them fit this budget; larger graphs recompute per chunk, which bounds the memory independently of E.
STASH_BYTES = int(float(os.environ.get("CGVP_WIDE_STASH_GB", "8")) * 2 ** 30)


def set_enabled(on):
    global ENABLED
    ENABLED = bool(on)


# ---- small helpers ---------------------------------------------------------------------------------------------------
def _act(code, x):
    if code == ACT_RELU:
        return torch.relu(x)
    if code == ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def _act_bwd(code, y, g):
    """g * act'(.) expressed through the activation OUTPUT y (as csrc/cgvp_reg.cuh::actb), one kernel each."""
    if code == ACT_RELU:
        return torch.ops.aten.threshold_backward(g, y, 0.0)        # g where y > 0
    if code == ACT_SIGMOID:
        return torch.ops.aten.sigmoid_backward(g, y)               # g * y * (1 - y)
    return g


_SQRT_EPS = {}


def _clamped_norm(vp):
    """`_norm_no_nan` over the planes, `sqrt(max(sum_xyz v^2, 1e-8))` (`models/gvp_layers.py:79-86`), as ONE reduction plus a
    clamp: sqrt is monotonic, so the clamp moves behind it with the bound sqrt(1e-8) (rounded in the working dtype)."""
    lo = _SQRT_EPS.get(vp.dtype)
    if lo is None:
        lo = _SQRT_EPS[vp.dtype] = float(torch.tensor(EPS, dtype=vp.dtype).sqrt())
    return torch.linalg.vector_norm(vp, dim=0).clamp_(min=lo), lo


def _mm3(vp, w_t):
    """Plane-major vectors [3, R, C] times w_t [C, H] -> [3, R, H] as ONE GEMM over 3R rows."""
    r = vp.shape[1]
    return (vp.reshape(3 * r, vp.shape[2]) @ w_t).view(3, r, w_t.shape[1])


WGRAD_BLOCK = 2048            # rows per partial product of a weight-gradient GEMM (see _tdot)


def _tdot(a, b):
    """a^T b for a [R, A], b [R, B] with R >> A, B: every weight-gradient GEMM of this module.  The result is one small tile,
    so a plain GEMM call leaves the whole reduction over R to the one or two CTAs that own it; here the rows are cut into
    blocks of WGRAD_BLOCK, the blocks' partial products are ONE batched GEMM (a CTA per block) and the partials are summed in
    block order -- a deterministic split-K."""
    r = a.shape[0]
    nb = r // WGRAD_BLOCK
    if nb < 4:
        return a.t() @ b
    m = nb * WGRAD_BLOCK
    out = torch.bmm(a[:m].reshape(nb, WGRAD_BLOCK, a.shape[1]).transpose(1, 2), b[:m].reshape(nb, WGRAD_BLOCK, b.shape[1])).sum(0)
    if m < r:
        out.addmm_(a[m:].t(), b[m:])
    return out


def _mm3_t(ap, bp):
    """sum over planes and rows of a^T b:  [3, R, A], [3, R, B] -> [A, B]  (weight gradients of the vector GEMMs)."""
    r = ap.shape[1]
    return _tdot(ap.reshape(3 * r, ap.shape[2]), bp.reshape(3 * r, bp.shape[2]))


def _planes(v):
    """[R, C, 3] (xyz innermost, the reference layout) -> [3, R, C]."""
    return v.permute(2, 0, 1).contiguous()


def _rows(vp):
    """[3, R, C] -> [R, C, 3]."""
    return vp.permute(1, 2, 0).contiguous()


def _segsum(rows, rowptr, index, n, mean=False):
    """out[i] = sum_{p in [rowptr[i], rowptr[i+1])} rows[index[p]]  (index None = identity; mean: / max(count, 1)),
    the deterministic CSR segmented reduction of the C ABI (`cgvp_segment_reduce`, csrc/plan.cu)."""
    if not rows.is_cuda:
        raise RuntimeError("castergvp.wide needs CUDA tensors (there is no CPU fallback)")
    rows = rows.contiguous()
    out = torch.empty(n, rows.shape[1], dtype=torch.float32, device=rows.device)
    if n == 0 or rows.shape[1] == 0:
        return out
    if rows.shape[0] == 0:
        return out.zero_()
    _lib.check(_lib.lib().cgvp_segment_reduce(C.c_void_p(rows.data_ptr()), int(rows.shape[1]), C.c_void_p(rowptr.data_ptr()),
                                              None if index is None else C.c_void_p(index.data_ptr()), int(n),
                                              _lib.AGGR_MEAN if mean else _lib.AGGR_SUM, 0, C.c_void_p(out.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "cgvp_segment_reduce")
    return out


class _tf32:
    """Library GEMMs in TF32 while the <= 1e-2 tensor-core mode is selected; fp32 otherwise (the caller's setting)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        if _lib.TENSOR_CORES:
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


# ---- one GVP (models/gvp_layers.py:142-175) on rows [R, S] / planes [3, R, C] -----------------------------------------------
class _Gvp:
    """Spec + PyTorch-layout weights of one GVP: wh [h, vi], ws [so, si + h], bs [so], wv [vo, h], wsv [vo, so], bg [vo]."""

    def __init__(self, spec, w6):
        self.spec = spec
        self.wh, self.ws, self.bs, self.wv, self.wsv, self.bg = w6

    def zero_grads(self):
        return [None if w is None else torch.zeros_like(w) for w in (self.wh, self.ws, self.bs, self.wv, self.wsv, self.bg)]


WH, WS, BS, WV, WSV, BG = range(6)


def _gvp_tail(g, sp, vh):
    """From the pre-activation scalars s' and the hidden vectors Vh to the GVP outputs (`:156-173`)."""
    sp_ = g.spec
    sv = {"sp": sp, "vh": vh}
    s_out = _act(sp_.sact, sp)
    v_out = None
    if sp_.vo > 0:
        if sp_.vi > 0:
            vo = _mm3(vh, g.wv.t())                                           # :156
            sg = None
            if sp_.has_gate:                                                  # :158-163 (s' is PRE-activation)
                gi = _act(sp_.vact, sp)
                sg = torch.sigmoid(torch.addmm(g.bg, gi, g.wsv.t()))
                sv["gi"] = gi
            elif sp_.vact != ACT_NONE:                                        # :164-166
                n2, lo = _clamped_norm(vo)
                sg = _act(sp_.vact, n2)
                sv["n2"], sv["lo"] = n2, lo
            v_out = vo if sg is None else vo * sg
            sv["vo"], sv["sg"] = vo, sg
        else:                                                                 # :169-171
            v_out = sp.new_zeros(3, sp.shape[0], sp_.vo)
    return s_out, v_out, sv


def _gvp_forward(g, s, vp):
    sp_ = g.spec
    if sp_.vi > 0:
        vh = _mm3(vp, g.wh.t())                                               # :151-152
        vn, lo = _clamped_norm(vh)                                            # :153
        sp = torch.addmm(g.bs, s, g.ws[:, :sp_.si].t())                       # :154, [s ; vn] never concatenated
        sp.addmm_(vn, g.ws[:, sp_.si:].t())
    else:
        vh = vn = lo = None
        sp = torch.addmm(g.bs, s, g.ws.t())                                   # :168
    s_out, v_out, sv = _gvp_tail(g, sp, vh)
    sv["vn"], sv["vn_lo"] = vn, lo
    return s_out, v_out, sv


def _gvp_bwd_core(g, sv, gs, gv, grads):
    """Through the output stage: returns ds' and the part of dVh that arrives through W_mu (None without vector outputs);
    accumulates the gradients of wv, wsv, bg."""
    sp_ = g.spec
    sp = sv["sp"]
    # ReLU: the sign of s' is the sign of relu(s'), so the mask comes from s' itself (no recompute of the activation)
    ds = _act_bwd(sp_.sact, sp if sp_.sact == ACT_RELU else _act(sp_.sact, sp), gs)
    dvh = None
    if sp_.vi > 0 and sp_.vo > 0 and gv is not None:
        vo, sg = sv["vo"], sv["sg"]
        if sp_.has_gate:
            dg = torch.ops.aten.sigmoid_backward((gv * vo).sum(0), sg)        # (gv . Vo) sg (1 - sg)
            dvo = gv * sg
            gi = sv["gi"]
            ds = torch.addmm(ds, dg, g.wsv) if sp_.vact == ACT_NONE else ds + _act_bwd(sp_.vact, gi, dg @ g.wsv)
            grads[WSV].add_(_tdot(dg, gi))
            grads[BG].add_(dg.sum(0))
        elif sp_.vact != ACT_NONE:
            dot = (gv * vo).sum(0)
            n2 = sv["n2"]
            t = (_act_bwd(sp_.vact, sg, dot) / n2).masked_fill_(n2 <= sv["lo"], 0)      # the clamp passes no gradient
            dvo = gv * sg + vo * t
        else:
            dvo = gv
        dvh = _mm3(dvo, g.wv)
        grads[WV].add_(_mm3_t(dvo, sv["vh"]))
    return ds, dvh


def _norm_bwd(sv, dvn, dvh):
    """dVh += Vh * dvn / vn where the clamp of `_norm_no_nan` passes."""
    f = dvn.div_(sv["vn"]).masked_fill_(sv["vn"] <= sv["vn_lo"], 0)
    return sv["vh"] * f if dvh is None else dvh.addcmul_(sv["vh"], f)


def _gvp_bwd_finish(g, sv, s_in, v_in, ds, dvh, grads, need_dx=True):
    """From ds' / dVh to the input gradients; accumulates the gradients of ws, bs, wh."""
    sp_ = g.spec
    grads[BS].add_(ds.sum(0))
    if sp_.vi > 0:
        grads[WS][:, :sp_.si].add_(_tdot(ds, s_in))
        grads[WS][:, sp_.si:].add_(_tdot(ds, sv["vn"]))
        dvh = _norm_bwd(sv, ds @ g.ws[:, sp_.si:], dvh)
        grads[WH].add_(_mm3_t(dvh, v_in))
        if not need_dx:
            return None, None
        return ds @ g.ws[:, :sp_.si], _mm3(dvh, g.wh)
    grads[WS].add_(_tdot(ds, s_in))
    return (ds @ g.ws if need_dx else None), None


# ---- LayerNorm (models/gvp_layers.py:231-242) -------------------------------------------------------------------------------
def _ln_fwd(s, vp, w, b):
    mean = s.mean(1, keepdim=True)
    xc = s - mean
    rstd = torch.rsqrt((xc * xc).mean(1, keepdim=True) + LN_EPS)
    xhat = xc * rstd
    sv = {"xhat": xhat, "rstd": rstd}
    yv = None
    if vp is not None:
        q = (vp * vp).sum(0)                                                  # [R, C]
        rms = q.clamp(min=EPS).mean(1, keepdim=True).sqrt()                   # :240-241
        yv = vp / rms
        sv["xv"], sv["q"], sv["rms"] = vp, q, rms
    return xhat * w + b, yv, sv


def _ln_bwd(sv, dys, dyv, w):
    xhat, rstd = sv["xhat"], sv["rstd"]
    dyh = dys * w
    m1 = dyh.mean(1, keepdim=True)
    m2 = (dyh * xhat).mean(1, keepdim=True)
    dxs = rstd * (dyh - m1 - xhat * m2)
    dw, db = (dys * xhat).sum(0), dys.sum(0)
    dxv = None
    if dyv is not None:
        xv, q, rms = sv["xv"], sv["q"], sv["rms"]
        dot = (dyv * xv).sum((0, 2)).unsqueeze(1)                             # [R, 1]
        coef = dot / (xv.shape[2] * rms * rms * rms)
        dxv = dyv / rms - xv * torch.where(q >= EPS, coef.expand_as(q), torch.zeros_like(q))
    return dxs, dxv, dw, db


# ---- fused GVPConv (models/gvp_layers.py:291-308) ---------------------------------------------------------------------------
def conv_supported(prog):
    """Descriptors this module serves: wide node scalars, vector channels on nodes, every message GVP with vectors."""
    return (ENABLED and prog.ns >= MIN_DIM and prog.nv > 0 and all(g.vi > 0 and g.vo > 0 for g in prog.gvps)
            and prog.gvps[0].si == 2 * prog.ns + prog.es and prog.gvps[0].vi == 2 * prog.nv + prog.ev)


def _pad4(x):
    return (x + 3) // 4 * 4


class _ConvPass:
    """Everything the chunks of one conv call share: split weights of message GVP 0, per-node projections, index views."""

    def __init__(self, prog, plan, x_s, x_v, e_s, e_v, weights):
        self.prog, self.plan = prog, plan
        self.ns, self.nv, self.es, self.ev = prog.ns, prog.nv, prog.es, prog.ev
        self.g = [_Gvp(sp, weights[6 * i: 6 * i + 6]) for i, sp in enumerate(prog.gvps)]
        self.E, self.N = int(plan.E), int(plan.N)
        self.edge_sorted = bool(prog.desc.edge_sorted)
        self.src, self.dst = plan.src.long(), plan.dst.long()
        self.eid = None if self.edge_sorted else plan.perm.long()
        self.x_s, self.e_s, self.e_v = x_s, e_s, e_v
        self.xvp = _planes(x_v)
        g0, ns, nv, es, ev = self.g[0], self.ns, self.nv, self.es, self.ev
        self.ws_j, self.ws_e = g0.ws[:, :ns], g0.ws[:, ns:ns + es]
        self.ws_i, self.ws_vn = g0.ws[:, ns + es:2 * ns + es], g0.ws[:, 2 * ns + es:]
        self.wh_j, self.wh_e, self.wh_i = g0.wh[:, :nv], g0.wh[:, nv:nv + ev], g0.wh[:, nv + ev:]
        # per-node projections: the node blocks of W_s / W_h of message GVP 0, once per node instead of once per edge
        self.ps_j = x_s @ self.ws_j.t()
        self.ps_i = torch.addmm(g0.bs, x_s, self.ws_i.t())
        self.pv_j = _mm3(self.xvp, self.wh_j.t())
        self.pv_i = _mm3(self.xvp, self.wh_i.t())

    def chunks(self):
        step = max(int(CHUNK_EDGES), 1)
        for p0 in range(0, self.E, step):
            yield p0, min(p0 + step, self.E)

    def edge_rows(self, p0, p1):
        """Edge attributes of the chunk in processing (dst-sorted) order: e_s [Ec, es], e_V planes [3, Ec, ev]."""
        if self.edge_sorted:
            es_c, ev_c = self.e_s[p0:p1], self.e_v[p0:p1]
        else:
            ids = self.eid[p0:p1]
            es_c, ev_c = self.e_s.index_select(0, ids), self.e_v.index_select(0, ids)
        return es_c, (_planes(ev_c) if self.ev > 0 else None)

    def forward_chunk(self, p0, p1, es_c, evp_c):
        """The message chain on edges [p0, p1).  Returns the per-GVP (s_out, V_out) and saved intermediates."""
        g0 = self.g[0]
        s_, d_ = self.src[p0:p1], self.dst[p0:p1]
        vh = self.pv_j.index_select(1, s_)
        vh.add_(self.pv_i.index_select(1, d_))
        if self.ev > 0:
            vh.add_(_mm3(evp_c, self.wh_e.t()))
        vn, lo = _clamped_norm(vh)
        sp = self.ps_j.index_select(0, s_)
        sp.add_(self.ps_i.index_select(0, d_))
        if self.es > 0:
            sp.addmm_(es_c, self.ws_e.t())
        sp.addmm_(vn, self.ws_vn.t())
        s, v, sv = _gvp_tail(g0, sp, vh)
        sv["vn"], sv["vn_lo"] = vn, lo
        outs, saves = [(s, v)], [sv]
        for g in self.g[1:]:
            s, v, sv = _gvp_forward(g, s, v)
            outs.append((s, v))
            saves.append(sv)
        return outs, saves


def _tensor_bytes(obj, seen):
    """Bytes of the distinct tensors reachable from a nest of tuples / lists / dicts."""
    if torch.is_tensor(obj):
        key = (obj.data_ptr(), obj.numel())
        if key in seen:
            return 0
        seen.add(key)
        return obj.numel() * obj.element_size()
    if isinstance(obj, dict):
        return sum(_tensor_bytes(v, seen) for v in obj.values())
    if isinstance(obj, (list, tuple)):
        return sum(_tensor_bytes(v, seen) for v in obj)
    return 0


def conv_forward(prog, plan, x_s, x_v, e_s, e_v, weights, kept=None):
    """(out_s [N, ns], out_v [N, nv, 3]) of the fused GVPConv: chunked GEMM message chain, per-edge message rows, one
    deterministic segmented reduction over the sorted targets (sum, or mean = / max(in-degree, 1)).
    `kept`: an empty list (training).  If the intermediates of ALL chunks fit STASH_BYTES they are appended to it, one entry
    per chunk, and `conv_backward(..., kept=kept)` then skips its recompute; otherwise the list stays empty."""
    so, vo = prog.out_s, prog.out_v
    n = int(plan.N)
    if plan.E == 0:
        return x_s.new_zeros(n, so), x_s.new_zeros(n, vo, 3)
    with _tf32(), torch.no_grad():
        cp = _ConvPass(prog, plan, x_s, x_v, e_s, e_v, weights)
        width = _pad4(so + 3 * vo)
        msg = x_s.new_empty(cp.E, width)
        if width > so + 3 * vo:
            msg[:, so + 3 * vo:].zero_()
        keep = kept is not None
        for p0, p1 in cp.chunks():
            es_c, evp_c = cp.edge_rows(p0, p1)
            outs, saves = cp.forward_chunk(p0, p1, es_c, evp_c)
            ms, mv = outs[-1]
            msg[p0:p1, :so] = ms
            msg[p0:p1, so:so + 3 * vo].unflatten(1, (vo, 3)).copy_(mv.permute(1, 2, 0))       # _merge layout, :101-109
            if keep:
                entry = (es_c, evp_c, outs, saves)
                if p0 == 0 and _tensor_bytes(entry, set()) / (p1 - p0) * cp.E > STASH_BYTES:
                    keep = False
                else:
                    kept.append(entry)
            del outs, saves
        out = _segsum(msg, plan.rowptr, None, n, mean=prog.desc.aggr == _lib.AGGR_MEAN)
        return out[:, :so].contiguous(), out[:, so:so + 3 * vo].reshape(n, vo, 3).contiguous()


def conv_backward(prog, plan, x_s, x_v, e_s, e_v, weights, d_out_s, d_out_v, kept=None):
    """Gradients of `conv_forward`: (d_x_s, d_x_v, d_e_s, d_e_v, [6 weight gradients per message GVP]).  The message chain is
    recomputed per chunk unless the forward of the same call left its intermediates in `kept` (see `conv_forward`)."""
    n, e = int(plan.N), int(plan.E)
    gvps = [_Gvp(sp, weights[6 * i: 6 * i + 6]) for i, sp in enumerate(prog.gvps)]
    grads = [g.zero_grads() for g in gvps]
    d_e_s, d_e_v = torch.zeros_like(e_s), torch.zeros_like(e_v)
    if e == 0:
        return torch.zeros_like(x_s), torch.zeros_like(x_v), d_e_s, d_e_v, [t for g in grads for t in g]
    with _tf32(), torch.no_grad():
        cp = _ConvPass(prog, plan, x_s, x_v, e_s, e_v, weights)
        ns, nv, es, ev = cp.ns, cp.nv, cp.es, cp.ev
        g0, gr0 = cp.g[0], grads[0]
        so0, h0 = g0.spec.so, g0.spec.h
        if prog.desc.aggr == _lib.AGGR_MEAN:                  # d(message_e) = d_out[dst_e] / max(deg, 1)
            deg = (plan.rowptr[1:] - plan.rowptr[:-1]).clamp(min=1).to(d_out_s.dtype)
            d_out_s = d_out_s / deg.unsqueeze(1)
            d_out_v = d_out_v / deg.view(-1, 1, 1)
        dovp = _planes(d_out_v)
        # per-edge rows [ds'_0 | dVh_0 plane x | y | z] in sorted order, reduced per node over both CSR views afterwards
        width = _pad4(so0 + 3 * h0)
        drows = x_s.new_empty(e, width)
        if width > so0 + 3 * h0:
            drows[:, so0 + 3 * h0:].zero_()
        chunks = list(cp.chunks())
        if kept is not None and len(kept) != len(chunks):
            kept = None
        for ci, (p0, p1) in enumerate(chunks):
            d_ = cp.dst[p0:p1]
            if kept is not None:
                es_c, evp_c, outs, saves = kept[ci]
                kept[ci] = None                                    # release the chunk's intermediates as soon as it is done
            else:
                es_c, evp_c = cp.edge_rows(p0, p1)
                outs, saves = cp.forward_chunk(p0, p1, es_c, evp_c)
            gs, gv = d_out_s.index_select(0, d_), dovp.index_select(1, d_)
            for k in range(len(cp.g) - 1, 0, -1):
                ds, dvh = _gvp_bwd_core(cp.g[k], saves[k], gs, gv, grads[k])
                gs, gv = _gvp_bwd_finish(cp.g[k], saves[k], outs[k - 1][0], outs[k - 1][1], ds, dvh, grads[k])
            sv0 = saves[0]
            ds0, dvh0 = _gvp_bwd_core(g0, sv0, gs, gv, gr0)
            # message GVP 0: edge / norm blocks per edge, node blocks after the per-node reductions below
            if es > 0:
                gr0[WS][:, ns:ns + es].add_(_tdot(ds0, es_c))
            gr0[WS][:, 2 * ns + es:].add_(_tdot(ds0, sv0["vn"]))
            dvh0 = _norm_bwd(sv0, ds0 @ cp.ws_vn, dvh0)
            if ev > 0:
                gr0[WH][:, nv:nv + ev].add_(_mm3_t(dvh0, evp_c))
                dev_rows = _rows(_mm3(dvh0, cp.wh_e))
            if es > 0:
                des_rows = ds0 @ cp.ws_e
            if cp.edge_sorted:
                if es > 0:
                    d_e_s[p0:p1] = des_rows
                if ev > 0:
                    d_e_v[p0:p1] = dev_rows
            else:
                ids = cp.eid[p0:p1]
                if es > 0:
                    d_e_s.index_copy_(0, ids, des_rows)
                if ev > 0:
                    d_e_v.index_copy_(0, ids, dev_rows)
            drows[p0:p1, :so0] = ds0
            for k in range(3):
                drows[p0:p1, so0 + k * h0: so0 + (k + 1) * h0] = dvh0[k]
            del outs, saves
        r_i = _segsum(drows, plan.rowptr, None, n)                 # over the in-edges of each node
        r_j = _segsum(drows, plan.srowptr, plan.sperm, n)          # over its out-edges
        del drows
        ris, rjs = r_i[:, :so0], r_j[:, :so0]
        d_x_s = ris @ cp.ws_i
        d_x_s.addmm_(rjs, cp.ws_j)
        gr0[WS][:, :ns].copy_(_tdot(rjs, x_s))
        gr0[WS][:, ns + es:2 * ns + es].copy_(_tdot(ris, x_s))
        gr0[BS].copy_(ris.sum(0))                                  # every edge has exactly one target
        rvi = r_i[:, so0:so0 + 3 * h0].reshape(n, 3, h0).permute(1, 0, 2).contiguous()
        rvj = r_j[:, so0:so0 + 3 * h0].reshape(n, 3, h0).permute(1, 0, 2).contiguous()
        d_x_v = _rows(_mm3(rvi, cp.wh_i).add_(_mm3(rvj, cp.wh_j)))
        gr0[WH][:, :nv].copy_(_mm3_t(rvj, cp.xvp))
        gr0[WH][:, nv + ev:].copy_(_mm3_t(rvi, cp.xvp))
    return d_x_s, d_x_v, d_e_s, d_e_v, [t for g in grads for t in g]


# ---- row programs (include/castergvp.h: CgvpRowDesc) without gather / one-hot: node update, GVP, LayerNorm --------------------
def rows_supported(prog, t):
    return (ENABLED and prog.in_s >= MIN_DIM and prog.onehot == 0 and t.get("in_index") is None and t.get("types") is None
            and (not prog.residual_in or not prog.gvps or (prog.gvps[-1].so == prog.in_s and prog.gvps[-1].vo == prog.in_v)))


def _rows_recompute(prog, t, gvps):
    has_v = prog.in_v > 0
    x_s = t["in_s"]
    x_v = _planes(t["in_v"]) if has_v else None
    if prog.residual_in:                                                       # x + D0(dh), :407
        h_s = t["h_s"] if t.get("mask0_s") is None else t["h_s"] * t["mask0_s"]
        x_s = x_s + h_s
        if has_v:
            h_v = _planes(t["h_v"])
            x_v = x_v + (h_v if t.get("mask0_v") is None else h_v * t["mask0_v"])
    ln0 = None
    if prog.pre_norm:
        x_s, x_v, ln0 = _ln_fwd(x_s, x_v, t["ln0_w"], t["ln0_b"])
    s, v, ins, saves = x_s, x_v, [], []
    for g in gvps:
        ins.append((s, v))
        s, v, sv = _gvp_forward(g, s, v)
        saves.append(sv)
    if prog.post_residual:                                                     # x + D1(ff(x)), :410
        s = x_s + (s if t.get("mask1_s") is None else s * t["mask1_s"])
        if v is not None and x_v is not None:
            v = x_v + (v if t.get("mask1_v") is None else v * t["mask1_v"])
    ln1 = None
    if prog.post_norm:
        s, v, ln1 = _ln_fwd(s, v, t["ln1_w"], t["ln1_b"])
    return s, v, dict(ln0=ln0, ln1=ln1, ins=ins, saves=saves)


def rows_forward(prog, t, weights):
    """(out_s, out_v [R, C, 3] or None) of a row program in GEMM form (used by the unit tests; the forward of the wide node
    update has its own kernels, csrc/rows_wide.cu)."""
    gvps = [_Gvp(sp, weights[6 * i: 6 * i + 6]) for i, sp in enumerate(prog.gvps)]
    with _tf32(), torch.no_grad():
        s, v, _ = _rows_recompute(prog, t, gvps)
    return s, (None if v is None else _rows(v))


def rows_backward(prog, t, weights, d_out_s, d_out_v):
    """Gradients of a row program: dict(d_in_s, d_in_v, d_h_s, d_h_v, ln=[d_ln0_w, d_ln0_b, d_ln1_w, d_ln1_b], dw=[...])."""
    gvps = [_Gvp(sp, weights[6 * i: 6 * i + 6]) for i, sp in enumerate(prog.gvps)]
    grads = [g.zero_grads() for g in gvps]
    with _tf32(), torch.no_grad():
        _, v_out, fw = _rows_recompute(prog, t, gvps)
        ds_, dv_ = d_out_s, (_planes(d_out_v) if (d_out_v is not None and v_out is not None) else None)
        ln = [None] * 4
        if prog.post_norm:
            ds_, dv_, ln[2], ln[3] = _ln_bwd(fw["ln1"], ds_, dv_, t["ln1_w"])
        skip_s = skip_v = None
        if prog.post_residual:
            skip_s, skip_v = ds_, dv_
            if t.get("mask1_s") is not None:
                ds_ = ds_ * t["mask1_s"]
            if dv_ is not None and t.get("mask1_v") is not None:
                dv_ = dv_ * t["mask1_v"]
        for k in range(len(gvps) - 1, -1, -1):
            d1, dvh = _gvp_bwd_core(gvps[k], fw["saves"][k], ds_, dv_, grads[k])
            ds_, dv_ = _gvp_bwd_finish(gvps[k], fw["saves"][k], fw["ins"][k][0], fw["ins"][k][1], d1, dvh, grads[k])
        if skip_s is not None:
            ds_ = ds_ + skip_s
            if skip_v is not None:
                dv_ = skip_v if dv_ is None else dv_ + skip_v
        if prog.pre_norm:
            ds_, dv_, ln[0], ln[1] = _ln_bwd(fw["ln0"], ds_, dv_, t["ln0_w"])
        has_v = prog.in_v > 0
        if has_v and dv_ is None:
            dv_ = torch.zeros_like(_planes(t["in_v"]))
        out = dict(d_in_s=ds_, d_in_v=_rows(dv_) if has_v else None, d_h_s=None, d_h_v=None, ln=ln,
                   dw=[x for g in grads for x in g])
        if prog.residual_in:
            out["d_h_s"] = ds_ if t.get("mask0_s") is None else ds_ * t["mask0_s"]
            if has_v:
                out["d_h_v"] = _rows(dv_ if t.get("mask0_v") is None else dv_ * t["mask0_v"])
    return out
