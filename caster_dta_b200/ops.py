"""Torch-facing wrappers of the C ABI: graph plans, weight packing and the two autograd Functions.

PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic of the GVP hot path
happens in `libcastergvp.so`.  Every call is enqueued on the current CUDA stream.
"""
import ctypes as C
import functools
import os
import weakref

import torch

from . import _lib, wide
from ._lib import lib, check


DEBUG_CHECKS = os.environ.get("CGVP_DEBUG", "0") == "1"     # index-range checks with device syncs (debugging aid)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("castergvp ops need CUDA tensors (there is no CPU fallback)")
    return t.detach().contiguous().float()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256) + 256, dtype=torch.uint8, device=device)


def _aligned_ptr(ws, align=256):
    p = ws.data_ptr()
    return C.c_void_p((p + align - 1) // align * align), ws.numel() - align


# ---- dropout-mask recording (parity tests of train-mode steps) ---------------------------------------------------------
# None: every dropout site uses its stock path.  A dict: each site draws its keep-mask explicitly (same distribution,
# `mask / (1 - p)`), applies it by multiplication and leaves the mask under its site name, so that a test can feed the
# identical masks to the CPU oracle.  Works under CUDA-graph capture (the recorded tensors are the graph's own buffers
# and hold the masks of the latest replay).
MASK_LOG = None


def dropout(mod, x, site):
    """`mod(x)` for an `nn.Dropout` module, recording the mask under `site` while MASK_LOG is a dict."""
    if MASK_LOG is None or not mod.training or mod.p == 0:
        return mod(x)
    m = torch.nn.functional.dropout(torch.ones_like(x), mod.p, True)
    MASK_LOG[site] = m
    return x * m


# ----------------------------------------------------------------------------------------------------------------
class GvpSpec:
    """Static description of one GVP (`models/gvp_layers.py:123-140`)."""
    __slots__ = ("si", "vi", "so", "vo", "h", "sact", "vact", "gate")

    def __init__(self, si, vi, so, vo, h, sact, vact, gate):
        self.si, self.vi, self.so, self.vo = int(si), int(vi), int(so), int(vo)
        self.h = int(h) if vi else 0
        self.sact, self.vact, self.gate = int(sact), int(vact), int(bool(gate))

    def desc(self):
        return _lib.GvpDesc(self.si, self.vi, self.so, self.vo, max(self.h, 0), self.sact, self.vact, self.gate)

    def key(self):
        return (self.si, self.vi, self.so, self.vo, self.h, self.sact, self.vact, self.gate)

    @property
    def has_wh(self):
        return self.vi > 0

    @property
    def has_wv(self):
        return self.vi > 0 and self.vo > 0

    @property
    def has_gate(self):
        return self.has_wv and self.gate == 1

    def packed_floats(self):
        d = self.desc()
        return int(lib().cgvp_gvp_packed_floats(C.byref(d)))


WEIGHTS_PER_GVP = 6   # wh, ws.weight, ws.bias, wv, wsv.weight, wsv.bias (None where absent)


def pack_weights(specs, weights, device):
    """PyTorch-layout parameters -> one packed arena.  Returns (arena, [block offsets in floats])."""
    offs, total = [], 0
    for sp in specs:
        offs.append(total)
        total += (sp.packed_floats() + 3) // 4 * 4
    arena = torch.empty(max(total, 4), dtype=torch.float32, device=device)
    n = len(specs)
    if n == 0:
        return arena, offs
    descs = (_lib.GvpDesc * n)(*[sp.desc() for sp in specs])
    wts = (_lib.GvpWeights * n)()
    keep = []
    for i in range(n):
        w = [_f32(t) for t in weights[WEIGHTS_PER_GVP * i: WEIGHTS_PER_GVP * (i + 1)]]
        keep.append(w)
        for name, t in zip(("wh", "ws", "bs", "wv", "wsv", "bg"), w):
            setattr(wts[i], name, None if t is None else t.data_ptr())
    blocks = (C.c_void_p * n)(*[arena.data_ptr() + 4 * o for o in offs])
    _lib.timed_call("cgvp_pack_weights", lib().cgvp_pack_weights, n, descs, wts, blocks, _stream())
    return arena, offs


def unpack_grads(specs, packed_grads, offs, weights):
    """Packed gradient blocks -> list of PyTorch-layout gradient tensors (None where the weight is None)."""
    n = len(specs)
    out = []
    if n == 0:
        return out
    descs = (_lib.GvpDesc * n)(*[sp.desc() for sp in specs])
    gr = (_lib.GvpGrads * n)()
    for i in range(n):
        for j, name in enumerate(("wh", "ws", "bs", "wv", "wsv", "bg")):
            w = weights[WEIGHTS_PER_GVP * i + j]
            g = None if w is None else torch.empty(w.shape, dtype=torch.float32, device=packed_grads.device)
            out.append(g)
            setattr(gr[i], name, None if g is None else g.data_ptr())
    blocks = (C.c_void_p * n)(*[packed_grads.data_ptr() + 4 * o for o in offs])
    _lib.timed_call("cgvp_unpack_grads", lib().cgvp_unpack_grads, n, descs, blocks, gr, _stream())
    return out


# ----------------------------------------------------------------------------------------------------------------
class GraphPlan:
    """dst-sorted / src-sorted CSR views of an edge_index (see `CgvpPlan` in castergvp.h)."""

    def __init__(self, edge_index, num_nodes):
        if not edge_index.is_cuda:
            raise RuntimeError("GraphPlan needs a CUDA edge_index (there is no CPU fallback)")
        ei = edge_index.detach().contiguous().long()
        dev = ei.device
        self.E, self.N = int(ei.shape[1]), int(num_nodes)
        if DEBUG_CHECKS and self.E and not torch.cuda.is_current_stream_capturing():
            # the kernels trust the indices (as PyG's gather does); CGVP_DEBUG=1 checks them here, at the cost of a device sync
            lo, hi = int(ei.min()), int(ei.max())
            if lo < 0 or hi >= self.N:
                raise ValueError(f"edge_index holds node ids in [{lo}, {hi}] but the graph has {self.N} nodes")
        i32 = dict(dtype=torch.int32, device=dev)
        self.perm = torch.empty(self.E, **i32)
        self.src = torch.empty(self.E, **i32)
        self.dst = torch.empty(self.E, **i32)
        self.rowptr = torch.empty(self.N + 1, **i32)
        self.sperm = torch.empty(self.E, **i32)
        self.srowptr = torch.empty(self.N + 1, **i32)
        nbytes = lib().cgvp_plan_workspace_bytes(self.E, self.N)
        if nbytes < 0:
            raise RuntimeError("graph too large for int32 indexing")
        ws = _workspace(nbytes, dev)
        wp, wn = _aligned_ptr(ws)
        self.c = _lib.Plan(self.E, self.N, self.perm.data_ptr(), self.src.data_ptr(), self.dst.data_ptr(),
                           self.rowptr.data_ptr(), self.sperm.data_ptr(), self.srowptr.data_ptr())
        check(lib().cgvp_plan_build(_ptr(ei), C.byref(self.c), wp, wn, _stream()), "cgvp_plan_build")
        self._keep = ei


_plan_cache = {}


def get_plan(edge_index, num_nodes):
    """Plan cached per edge_index tensor (same storage, shape and version -> same plan), so the two conv layers of
    a model and repeated eager steps over the same batch sort the edges once.

    While a CUDA graph is being captured the cache is bypassed in both directions: the graph will be replayed after
    the caller has refilled `edge_index` with another batch, so `cgvp_plan_build` has to be part of the captured work
    (a cached plan would silently keep the old perm / rowptr), and a plan built on graph-private memory must not be
    handed to later eager calls."""
    if edge_index.is_cuda and torch.cuda.is_current_stream_capturing():
        return GraphPlan(edge_index, num_nodes)
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), edge_index.device.index)
    hit = _plan_cache.get(key)
    if hit is not None and hit[0]() is not None:
        return hit[1]
    if len(_plan_cache) > 64:
        _plan_cache.clear()
    plan = GraphPlan(edge_index, num_nodes)
    _plan_cache[key] = (weakref.ref(edge_index), plan)
    return plan


# ----------------------------------------------------------------------------------------------------------------
class RowProgram:
    """Static description of a fused row program (`CgvpRowDesc`)."""

    def __init__(self, in_s, in_v, gvps, onehot=0, residual_in=False, pre_norm=False, post_residual=False,
                 post_norm=False):
        assert len(gvps) <= _lib.MAX_CHAIN
        self.in_s, self.in_v, self.gvps = int(in_s), int(in_v), list(gvps)
        self.onehot, self.residual_in, self.pre_norm = int(onehot), bool(residual_in), bool(pre_norm)
        self.post_residual, self.post_norm = bool(post_residual), bool(post_norm)
        if gvps:
            self.out_s, self.out_v = gvps[-1].so, gvps[-1].vo
        else:
            self.out_s, self.out_v = self.onehot + self.in_s, self.in_v
        d = _lib.RowDesc()
        d.in_s, d.in_v, d.onehot, d.has_residual_in = self.in_s, self.in_v, self.onehot, int(self.residual_in)
        d.pre_norm, d.n_gvp, d.post_residual, d.post_norm = int(self.pre_norm), len(gvps), int(self.post_residual), int(self.post_norm)
        for i, g in enumerate(gvps):
            d.gvp[i] = g.desc()
        self.desc = d


def _row_args(prog, rows, t, arena, offs):
    a = _lib.RowArgs()
    a.rows = rows
    for name in ("in_s", "in_v", "types", "in_index", "h_s", "h_v", "mask0_s", "mask0_v", "mask1_s", "mask1_v",
                 "ln0_w", "ln0_b", "ln1_w", "ln1_b", "out_s", "out_v", "stash"):
        v = t.get(name)
        setattr(a, name, None if v is None else v.data_ptr())
    blocks = (C.c_void_p * max(len(offs), 1))(*[arena.data_ptr() + 4 * o for o in offs])
    a.h_packed = C.addressof(blocks)
    return a, blocks


class RowsFunction(torch.autograd.Function):
    """autograd wrapper of cgvp_rows_fwd / cgvp_rows_bwd."""

    @staticmethod
    def forward(ctx, prog, in_s, in_v, types, in_index, h_s, h_v, m0s, m0v, m1s, m1v, ln0_w, ln0_b, ln1_w, ln1_b,
                *weights):
        dev = in_s.device
        t = dict(in_s=_f32(in_s), in_v=_f32(in_v) if prog.in_v else None,
                 types=None if types is None else types.contiguous().long(),
                 in_index=None if in_index is None else in_index.contiguous().int(),
                 h_s=_f32(h_s), h_v=_f32(h_v) if prog.in_v else None,
                 mask0_s=_f32(m0s), mask0_v=_f32(m0v), mask1_s=_f32(m1s), mask1_v=_f32(m1v),
                 ln0_w=_f32(ln0_w), ln0_b=_f32(ln0_b), ln1_w=_f32(ln1_w), ln1_b=_f32(ln1_b))
        rows = int(in_index.shape[0]) if in_index is not None else int(in_s.shape[0])
        arena, offs = pack_weights(prog.gvps, weights, dev)
        t["out_s"] = torch.empty(rows, prog.out_s, dtype=torch.float32, device=dev)
        t["out_v"] = torch.empty(rows, prog.out_v, 3, dtype=torch.float32, device=dev)
        # training: the specialised kernels leave the GVPs' pre-activation scalars per row for the backward (USE_ROWS_STASH)
        t["stash"] = None
        if USE_ROWS_STASH and any(ctx.needs_input_grad) and rows > 0:
            sf = int(lib().cgvp_rows_stash_floats(C.byref(prog.desc)))
            if sf > 0:
                t["stash"] = torch.empty(rows, sf, dtype=torch.float32, device=dev)
        a, blocks = _row_args(prog, rows, t, arena, offs)
        nbytes = lib().cgvp_rows_workspace_bytes(C.byref(prog.desc), rows, 0)
        ws = _workspace(nbytes, dev)
        wp, wn = _aligned_ptr(ws)
        _lib.timed_call("cgvp_rows_fwd", lib().cgvp_rows_fwd, C.byref(prog.desc), C.byref(a), wp, wn, _stream())
        out_s, out_v = t.pop("out_s"), t.pop("out_v")     # never keep the outputs on ctx: node -> output -> node leaks
        ctx.prog, ctx.t, ctx.arena, ctx.offs, ctx.rows = prog, t, arena, offs, rows
        ctx.weights = weights
        ctx.in_rows = int(in_s.shape[0])
        ctx.mark_non_differentiable(*([] if prog.out_v else [out_v]))
        return out_s, out_v

    @staticmethod
    def backward(ctx, d_out_s, d_out_v):
        prog, t, rows = ctx.prog, dict(ctx.t), ctx.rows
        dev = t["in_s"].device
        need = ctx.needs_input_grad
        if rows > 0 and wide.rows_supported(prog, t):
            # wide dims (config 5): the backward as a short sequence of library GEMMs (wide.py) instead of the generic tile kernel
            r = wide.rows_backward(prog, t, [_f32(w) for w in ctx.weights], _f32(d_out_s), _f32(d_out_v) if prog.out_v else None)
            want_h = prog.residual_in and (need[5] or need[6])
            return (None, r["d_in_s"] if need[1] else None, r["d_in_v"] if (need[2] and prog.in_v) else None, None, None,
                    r["d_h_s"] if want_h else None, r["d_h_v"] if want_h else None, None, None, None, None, *r["ln"], *r["dw"])
        g = _lib.RowGradArgs()
        d_out_s = _f32(d_out_s)
        d_out_v = _f32(d_out_v) if prog.out_v else None
        keep = [d_out_s, d_out_v]
        g.d_out_s, g.d_out_v = d_out_s.data_ptr(), None if d_out_v is None else d_out_v.data_ptr()
        d_in_s = d_in_v = d_h_s = d_h_v = None
        gathered = t["in_index"] is not None
        alloc = torch.zeros if gathered else torch.empty
        if need[1]:
            d_in_s = alloc(ctx.in_rows, prog.in_s, dtype=torch.float32, device=dev)
            g.d_in_s = d_in_s.data_ptr()
        if need[2] and prog.in_v:
            d_in_v = alloc(ctx.in_rows, prog.in_v, 3, dtype=torch.float32, device=dev)
            g.d_in_v = d_in_v.data_ptr()
        if prog.residual_in and (need[5] or need[6]):
            d_h_s = torch.empty(rows, prog.in_s, dtype=torch.float32, device=dev)
            d_h_v = torch.empty(rows, prog.in_v, 3, dtype=torch.float32, device=dev)
            g.d_h_s, g.d_h_v = d_h_s.data_ptr(), d_h_v.data_ptr()
        lng = [None] * 4
        if prog.pre_norm:
            lng[0], lng[1] = torch.empty_like(t["ln0_w"]), torch.empty_like(t["ln0_b"])
            g.d_ln0_w, g.d_ln0_b = lng[0].data_ptr(), lng[1].data_ptr()
        if prog.post_norm:
            lng[2], lng[3] = torch.empty_like(t["ln1_w"]), torch.empty_like(t["ln1_b"])
            g.d_ln1_w, g.d_ln1_b = lng[2].data_ptr(), lng[3].data_ptr()
        total = sum((sp.packed_floats() + 3) // 4 * 4 for sp in prog.gvps)
        pg = torch.empty(max(total, 4), dtype=torch.float32, device=dev)
        gblocks = (C.c_void_p * max(len(ctx.offs), 1))(*[pg.data_ptr() + 4 * o for o in ctx.offs])
        g.h_packed_grads = C.addressof(gblocks)
        t["out_s"] = t["out_v"] = None
        a, blocks = _row_args(prog, rows, t, ctx.arena, ctx.offs)
        nbytes = lib().cgvp_rows_workspace_bytes(C.byref(prog.desc), rows, 1)
        ws = _workspace(nbytes, dev)
        wp, wn = _aligned_ptr(ws)
        _lib.timed_call("cgvp_rows_bwd", lib().cgvp_rows_bwd, C.byref(prog.desc), C.byref(a), C.byref(g), wp, wn, _stream())
        dw = unpack_grads(prog.gvps, pg, ctx.offs, ctx.weights)
        del keep
        return (None, d_in_s, d_in_v, None, None, d_h_s, d_h_v, None, None, None, None, *lng, *dw)


def run_rows(prog, in_s, in_v=None, types=None, in_index=None, h=None, masks0=None, masks1=None, ln0=None, ln1=None,
             weights=()):
    h_s, h_v = (None, None) if h is None else h
    m0s, m0v = (None, None) if masks0 is None else masks0
    m1s, m1v = (None, None) if masks1 is None else masks1
    l0w, l0b = (None, None) if ln0 is None else ln0
    l1w, l1b = (None, None) if ln1 is None else ln1
    return RowsFunction.apply(prog, in_s, in_v, types, in_index, h_s, h_v, m0s, m0v, m1s, m1v, l0w, l0b, l1w, l1b,
                              *weights)


# ----------------------------------------------------------------------------------------------------------------
class ConvProgram:
    """Static description of a fused GVPConv (`CgvpConvDesc`)."""

    def __init__(self, ns, nv, es, ev, gvps, aggr, edge_sorted=False):
        assert 1 <= len(gvps) <= _lib.MAX_CHAIN
        self.ns, self.nv, self.es, self.ev, self.gvps = int(ns), int(nv), int(es), int(ev), list(gvps)
        self.out_s, self.out_v = gvps[-1].so, gvps[-1].vo
        d = _lib.ConvDesc()
        d.ns, d.nv, d.es, d.ev, d.n_gvp = self.ns, self.nv, self.es, self.ev, len(gvps)
        for i, g in enumerate(gvps):
            d.gvp[i] = g.desc()
        d.aggr = _lib.AGGR_MEAN if aggr == "mean" else _lib.AGGR_SUM
        d.edge_sorted = int(bool(edge_sorted))
        self.desc = d


USE_CONV_STASH = True    # forward stash for the conv backward (cgvp_conv_fwd_stash / cgvp_conv_bwd_stash)
USE_ROWS_STASH = True    # forward stash of the row programs (CgvpRowArgs.stash): the backward skips the W_s recompute


class ConvFunction(torch.autograd.Function):
    """autograd wrapper of cgvp_conv_fwd / cgvp_conv_bwd."""

    @staticmethod
    def forward(ctx, prog, plan, x_s, x_v, e_s, e_v, *weights):
        dev = x_s.device
        x_s, x_v, e_s, e_v = _f32(x_s), _f32(x_v), _f32(e_s), _f32(e_v)
        ctx.prog, ctx.plan, ctx.saved, ctx.weights = prog, plan, (x_s, x_v, e_s, e_v), weights
        ctx.wide, ctx.kept = wide.conv_supported(prog), None
        if ctx.wide and not _lib.TENSOR_CORES:
            # wide dims (config 5), fp32 mode: chunked GEMM formulation (wide.py); the tensor-core mode keeps the fused
            # tcgen05 forward (csrc/conv_tc.cu) and shares the backward below
            ctx.kept = [] if any(ctx.needs_input_grad) else None          # training: intermediates for the backward, if they fit
            out_s, out_v = wide.conv_forward(prog, plan, x_s, x_v, e_s, e_v, [_f32(w) for w in weights], kept=ctx.kept)
            ctx.stash = ctx.arena = ctx.offs = None
            ctx.mark_non_differentiable(*([] if prog.out_v else [out_v]))
            return out_s, out_v
        arena, offs = pack_weights(prog.gvps, weights, dev)
        out_s = torch.empty(plan.N, prog.out_s, dtype=torch.float32, device=dev)
        out_v = torch.empty(plan.N, prog.out_v, 3, dtype=torch.float32, device=dev)
        blocks = (C.c_void_p * len(offs))(*[arena.data_ptr() + 4 * o for o in offs])
        nbytes = lib().cgvp_conv_workspace_bytes(C.byref(prog.desc), plan.E, plan.N, 0)
        ws = _workspace(nbytes, dev)
        wp, wn = _aligned_ptr(ws)
        # training: the specialised kernels leave the inputs of message GVPs 1 and 2 per edge for the backward (USE_CONV_STASH)
        stash = None
        if USE_CONV_STASH and any(ctx.needs_input_grad):
            sbytes = lib().cgvp_conv_stash_bytes(C.byref(prog.desc), plan.E)
            if sbytes > 0:
                stash = torch.empty(sbytes, dtype=torch.uint8, device=dev)
        _lib.timed_call("cgvp_conv_fwd", lib().cgvp_conv_fwd_stash, C.byref(prog.desc), C.byref(plan.c), _ptr(x_s), _ptr(x_v), _ptr(e_s),
                        _ptr(e_v), blocks, _ptr(out_s), _ptr(out_v), wp, wn, _ptr(stash), _stream())
        ctx.stash, ctx.arena, ctx.offs = stash, arena, offs
        ctx.mark_non_differentiable(*([] if prog.out_v else [out_v]))
        return out_s, out_v

    @staticmethod
    def backward(ctx, d_out_s, d_out_v):
        prog, plan = ctx.prog, ctx.plan
        x_s, x_v, e_s, e_v = ctx.saved
        dev = x_s.device
        d_out_s, d_out_v = _f32(d_out_s), _f32(d_out_v)
        if ctx.wide:
            kept, ctx.kept = ctx.kept, None
            d_x_s, d_x_v, d_e_s, d_e_v, dw = wide.conv_backward(prog, plan, x_s, x_v, e_s, e_v, [_f32(w) for w in ctx.weights],
                                                                d_out_s, d_out_v, kept=kept)
            return (None, None, d_x_s, d_x_v, d_e_s, d_e_v, *dw)
        d_x_s, d_x_v = torch.empty_like(x_s), torch.empty_like(x_v)
        d_e_s, d_e_v = torch.empty_like(e_s), torch.empty_like(e_v)
        total = sum((sp.packed_floats() + 3) // 4 * 4 for sp in prog.gvps)
        pg = torch.empty(max(total, 4), dtype=torch.float32, device=dev)
        blocks = (C.c_void_p * len(ctx.offs))(*[ctx.arena.data_ptr() + 4 * o for o in ctx.offs])
        gblocks = (C.c_void_p * len(ctx.offs))(*[pg.data_ptr() + 4 * o for o in ctx.offs])
        nbytes = lib().cgvp_conv_workspace_bytes(C.byref(prog.desc), plan.E, plan.N, 1)
        ws = _workspace(nbytes, dev)
        wp, wn = _aligned_ptr(ws)
        _lib.timed_call("cgvp_conv_bwd", lib().cgvp_conv_bwd_stash, C.byref(prog.desc), C.byref(plan.c), _ptr(x_s), _ptr(x_v), _ptr(e_s),
                        _ptr(e_v), blocks, _ptr(d_out_s), _ptr(d_out_v), _ptr(d_x_s), _ptr(d_x_v), _ptr(d_e_s), _ptr(d_e_v), 0,
                        gblocks, wp, wn, _ptr(ctx.stash), _stream())
        dw = unpack_grads(prog.gvps, pg, ctx.offs, ctx.weights)
        return (None, None, d_x_s, d_x_v, d_e_s, d_e_v, *dw)


def run_conv(prog, plan, x, edge_attr, weights):
    return ConvFunction.apply(prog, plan, x[0], x[1], edge_attr[0], edge_attr[1], *weights)


# ----------------------------------------------------------------------------------------------------------------
def gather_message_input(edge_index, x, edge_attr):
    """Stand-alone gather: (ms [E, 2ns+es], mv [E, 2nv+ev, 3]) as GVPConv.message builds them (gvp_layers.py:306)."""
    s, v = _f32(x[0]), _f32(x[1])
    es, ev = _f32(edge_attr[0]), _f32(edge_attr[1])
    ei = edge_index.contiguous().long()
    e = int(ei.shape[1])
    ns, nv, nes, nev = s.shape[1], v.shape[1], es.shape[1], ev.shape[1]
    ms = torch.empty(e, 2 * ns + nes, dtype=torch.float32, device=s.device)
    mv = torch.empty(e, 2 * nv + nev, 3, dtype=torch.float32, device=s.device)
    check(lib().cgvp_gather_message_input(_ptr(ei), e, ns, nv, nes, nev, _ptr(s), _ptr(v), _ptr(es), _ptr(ev), _ptr(ms),
                                          _ptr(mv), _stream()), "cgvp_gather_message_input")
    return ms, mv


def segment_reduce(rows, plan, aggr="sum", use_perm=True, out=None, beta=0):
    """Deterministic aggregation of per-edge rows at their target node (rows in ORIGINAL edge order if use_perm)."""
    rows = _f32(rows)
    width = rows.shape[1]
    if out is None:
        out = torch.empty(plan.N, width, dtype=torch.float32, device=rows.device)
    code = _lib.AGGR_MEAN if aggr == "mean" else _lib.AGGR_SUM
    check(lib().cgvp_segment_reduce(_ptr(rows), width, _ptr(plan.rowptr), _ptr(plan.perm) if use_perm else None, plan.N,
                                    code, int(beta), _ptr(out), _stream()), "cgvp_segment_reduce")
    return out


# ---- cross-attention core on packed rows (csrc/attention.cu) ---------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def attention_supported(num_heads, head_dim):
    return bool(lib().cgvp_attn_supported(int(num_heads), int(head_dim)))


class CrossAttnFunction(torch.autograd.Function):
    """out, weights = softmax(scale q k^T) v per graph on packed rows; `weights` (head-averaged, padded to
    [B, lq_max, lk_max]) is not differentiable -- the reference's training loop discards it (`train_model.py:564`)."""

    @staticmethod
    def forward(ctx, q, k, v, qptr, kptr, qbatch, kbatch, num_heads, lq_max, lk_max, q_fill, want_weights):
        q, k, v = _f32(q), _f32(k), _f32(v)
        nq, e = q.shape
        nk, b = k.shape[0], qptr.shape[0] - 1
        hd = e // num_heads
        scale = float(hd) ** -0.5
        dev = q.device
        out = torch.empty_like(q)
        stats = torch.empty(nq, num_heads, 2, dtype=torch.float32, device=dev)
        weights = w_fill = None
        if want_weights:
            weights = torch.zeros(b, lq_max, lk_max, dtype=torch.float32, device=dev)
            if q_fill is not None:
                q_fill = _f32(q_fill)
                w_fill = torch.zeros(b, lk_max, dtype=torch.float32, device=dev)
        _lib.timed_call("cgvp_attn_fwd", lib().cgvp_attn_fwd, _ptr(q), _ptr(k), _ptr(v), _ptr(qptr), _ptr(kptr), _ptr(qbatch), b,
                        nq, nk, num_heads, hd, scale, _ptr(q_fill) if w_fill is not None else None, int(lq_max), int(lk_max),
                        _ptr(out), _ptr(stats), _ptr(weights) if weights is not None else None,
                        _ptr(w_fill) if w_fill is not None else None, _stream())
        ctx.save_for_backward(q, k, v, out, stats, qptr, kptr, qbatch, kbatch)
        ctx.num_heads, ctx.scale = num_heads, scale
        if weights is None:
            weights = q.new_zeros(0)
        if w_fill is None:
            w_fill = q.new_zeros(0)
        ctx.mark_non_differentiable(weights, w_fill)
        return out, weights, w_fill

    @staticmethod
    def backward(ctx, d_out, _dw, _dwf):
        q, k, v, out, stats, qptr, kptr, qbatch, kbatch = ctx.saved_tensors
        d_out = _f32(d_out)
        nq, e = q.shape
        nk, b = k.shape[0], qptr.shape[0] - 1
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dsum = torch.empty(nq, ctx.num_heads, dtype=torch.float32, device=q.device)
        _lib.timed_call("cgvp_attn_bwd", lib().cgvp_attn_bwd, _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(stats), _ptr(d_out),
                        _ptr(qptr), _ptr(kptr), _ptr(qbatch), _ptr(kbatch), b, nq, nk, ctx.num_heads, e // ctx.num_heads,
                        ctx.scale, _ptr(dsum), _ptr(dq), _ptr(dk), _ptr(dv), _stream())
        return dq, dk, dv, None, None, None, None, None, None, None, None, None


# ---- dense layers around the encoder: weight / bias gradient on the tensor cores (csrc/linear_tc.cu) ---------------------
def linear_wgrad(dy, x, want_bias=True):
    """(dW [N,K], db [N] or None) = (dy^T x, column sums of dy) for dy [M,N], x [M,K]; fp32-accurate 3xTF32."""
    dy, x = _f32(dy), _f32(x)
    m, n = dy.shape
    k = x.shape[1]
    dw = torch.empty(n, k, dtype=torch.float32, device=dy.device)
    db = torch.empty(n, dtype=torch.float32, device=dy.device) if want_bias else None
    nbytes = lib().cgvp_linear_wgrad_workspace_bytes(m, n, k)
    ws = _workspace(nbytes, dy.device)
    wp, wn = _aligned_ptr(ws)
    _lib.timed_call("cgvp_linear_wgrad", lib().cgvp_linear_wgrad, _ptr(dy), _ptr(x), m, n, k, _ptr(dw), _ptr(db), wp, wn, _stream())
    return dw, db


@functools.lru_cache(maxsize=4096)
def linear_wgrad_supported(m, n, k):
    return bool(lib().cgvp_linear_wgrad_supported(int(m), int(n), int(k)))


# The forward / input-gradient GEMMs of csrc/linear_tc.cu are correct (fp32-accurate) but at the ~1 GFLOP sizes of this model
# they take ~30 us against ~16 us for the stock SIMT GEMM (the per-CTA weight-slice load is amortised over 1-2 row tiles), so
# they are opt-in; the weight-gradient kernel (3x faster than the stock path) is always on.
USE_TC_LINEAR_GEMM = False


def linear_gemm_supported(m, out_cols, red_len):
    return USE_TC_LINEAR_GEMM and bool(lib().cgvp_linear_gemm_supported(int(m), int(out_cols), int(red_len)))


def _linear_fwd(x, w, b):
    m, k = x.shape
    n = w.shape[0]
    if not linear_gemm_supported(m, n, k):
        return torch.nn.functional.linear(x, w, b)
    x, w, b = _f32(x), _f32(w), _f32(b)
    y = torch.empty(m, n, dtype=torch.float32, device=x.device)
    _lib.timed_call("cgvp_linear_fwd", lib().cgvp_linear_fwd, _ptr(x), _ptr(w), _ptr(b), m, n, k, _ptr(y), _stream())
    return y


def _linear_dgrad(dy, w):
    m, n = dy.shape
    k = w.shape[1]
    if not linear_gemm_supported(m, k, n):
        return dy @ w
    dy, w = _f32(dy), _f32(w)
    dx = torch.empty(m, k, dtype=torch.float32, device=dy.device)
    _lib.timed_call("cgvp_linear_dgrad", lib().cgvp_linear_dgrad, _ptr(dy), _ptr(w), m, n, k, _ptr(dx), _stream())
    return dx


WGRAD_STREAM = None      # optional side stream for the dense-layer weight gradients (set_wgrad_stream / join_wgrad_stream)


def set_wgrad_stream(stream):
    """Run `cgvp_linear_wgrad` on `stream` (None = on the backward's own stream).  The caller must call
    `join_wgrad_stream()` after `backward()` and before anything reads the parameter gradients."""
    global WGRAD_STREAM
    WGRAD_STREAM = stream


def join_wgrad_stream():
    if WGRAD_STREAM is not None:
        torch.cuda.current_stream().wait_stream(WGRAD_STREAM)


class LinearFunction(torch.autograd.Function):
    """y = x w^T + b on the tensor cores with fp32 accuracy (3xTF32, csrc/linear_tc.cu): forward, input gradient and
    weight / bias gradient; shapes outside the kernels' range fall back to the stock GEMMs."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        # the side stream is only safe when autograd merely STORES the returned gradients (leaf parameters): a sliced weight
        # (e.g. a third of nn.MultiheadAttention.in_proj_weight) sends them through more backward kernels on the main stream
        ctx.side_ok = w.is_leaf and (b is None or b.is_leaf)
        return _linear_fwd(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dw = db = None
        want_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        side = WGRAD_STREAM if ctx.side_ok else None
        if want_w and side is not None:
            # parameter gradients are leaves of the backward pass: compute them off the critical path (the driver joins
            # with join_wgrad_stream() after backward, before the optimizer / the end of a graph capture)
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dw, db = linear_wgrad(dy, x, ctx.has_bias)
            dy.record_stream(side)
            x.record_stream(side)
        dx = _linear_dgrad(dy, w) if ctx.needs_input_grad[0] else None
        if want_w and side is None:
            dw, db = linear_wgrad(dy, x, ctx.has_bias)
        return dx, dw, db


def linear(x, w, b=None):
    """`F.linear` for 2-D CUDA activations; large row counts take the tensor-core weight-gradient path."""
    if x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and w.dtype == torch.float32 and not torch.is_autocast_enabled():
        m, n, k = x.shape[0], w.shape[0], w.shape[1]
        if torch.is_grad_enabled() and (w.requires_grad or x.requires_grad):
            if linear_wgrad_supported(m, n, k):
                return LinearFunction.apply(x.contiguous(), w, b)
        elif linear_gemm_supported(m, n, k):
            return _linear_fwd(x.contiguous(), w, b)
    return torch.nn.functional.linear(x, w, b)


# ---- row LayerNorm on packed activations (csrc/layernorm.cu) ---------------------------------------------------------------
class LayerNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x, g, b = _f32(x), _f32(gamma), _f32(beta)
        rows, d = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device)
        _lib.timed_call("cgvp_layernorm_fwd", lib().cgvp_layernorm_fwd, _ptr(x), _ptr(g), _ptr(b), rows, d, float(eps), _ptr(y),
                        _ptr(stats), _stream())
        ctx.save_for_backward(x, g, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, stats = ctx.saved_tensors
        dy = _f32(dy)
        rows, d = x.shape
        dx, dg, db = torch.empty_like(x), torch.empty_like(g), torch.empty_like(g)
        ws = _workspace(lib().cgvp_layernorm_workspace_bytes(rows, d), x.device)
        wp, wn = _aligned_ptr(ws)
        _lib.timed_call("cgvp_layernorm_bwd", lib().cgvp_layernorm_bwd, _ptr(dy), _ptr(x), _ptr(stats), _ptr(g), rows, d, _ptr(dx),
                        _ptr(dg), _ptr(db), wp, wn, _stream())
        return dx, dg, db, None


@functools.lru_cache(maxsize=None)
def _layernorm_supported(d):
    return bool(lib().cgvp_layernorm_supported(d))


def layer_norm(x, mod):
    """`mod(x)` for an affine nn.LayerNorm over the last dimension of packed 2-D CUDA activations."""
    if (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and mod.elementwise_affine and mod.bias is not None
            and len(mod.normalized_shape) == 1 and x.shape[0] > 0 and _layernorm_supported(int(x.shape[1]))):
        return LayerNormFunction.apply(x, mod.weight, mod.bias, mod.eps)
    return mod(x)
