"""The GEMM formulation of the wide-dims GVPConv / node update (`caster_dta_b200/wide.py`) against the fp64 oracle, on the CPU.

`wide.py` is device-agnostic torch algebra around two C-ABI primitives (the CSR segmented sums); here the primitive is
substituted by a plain torch loop (test infrastructure) so that the algebra -- the node-level split of message GVP 0, the
plane-major vector layout, the hand-derived backward, chunking -- is checked against autograd through `oracle/gvp_oracle.py`
without a GPU.  The GPU tests (`tests/test_gpu_parity.py`) run the same entry points with the real primitives.
"""
import types

import pytest
import torch
import torch.nn.functional as F

from caster_dta_b200 import modules, ops, wide
from helpers import case, golden
from oracle import gvp_oracle

TOL = 1e-10


def cpu_plan(ei, n):
    """What `ops.GraphPlan` builds on the device (`cgvp_plan_build`): stable dst sort, then a stable source sort of it."""
    src, dst = ei[0], ei[1]
    perm = torch.argsort(dst, stable=True)
    s_sorted, d_sorted = src[perm], dst[perm]
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(d_sorted, minlength=n), 0)
    sperm = torch.argsort(s_sorted, stable=True)
    srowptr = torch.zeros(n + 1, dtype=torch.int64)
    srowptr[1:] = torch.cumsum(torch.bincount(s_sorted, minlength=n), 0)
    i32 = lambda t: t.to(torch.int32)
    return types.SimpleNamespace(E=int(ei.shape[1]), N=int(n), perm=i32(perm), src=i32(s_sorted), dst=i32(d_sorted),
                                 rowptr=i32(rowptr), sperm=i32(sperm), srowptr=i32(srowptr))


def segsum_reference(rows, rowptr, index, n, mean=False):
    out = rows.new_zeros(n, rows.shape[1])
    for i in range(n):
        a, b = int(rowptr[i]), int(rowptr[i + 1])
        if b > a:
            ids = torch.arange(a, b) if index is None else index[a:b].long()
            out[i] = rows[ids].sum(0)
            if mean:
                out[i] /= max(b - a, 1)
    return out


@pytest.fixture(autouse=True)
def _substitute_primitive(monkeypatch):
    monkeypatch.setattr(wide, "_segsum", segsum_reference)
    monkeypatch.setattr(wide, "ENABLED", True)
    monkeypatch.setattr(wide, "MIN_DIM", 1)
    monkeypatch.setattr(wide, "WGRAD_BLOCK", 16)      # edge-level weight gradients take the blocked path, node-level ones the plain one


def layer_case(n, e, nd, ed, seed, hub=False, isolated=False):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - (2 if isolated else 0), (e,), generator=g)      # isolated: the last two nodes get no in-edges
    if hub and e:
        dst[: e // 3] = 3
    ei = torch.stack([src, dst])
    p = gvp_oracle.init_conv_layer_params({}, "", nd, ed, gen=g, dtype=torch.float64)
    x = (torch.randn(n, nd[0], generator=g, dtype=torch.float64), torch.randn(n, nd[1], 3, generator=g, dtype=torch.float64))
    ea = (torch.randn(e, ed[0], generator=g, dtype=torch.float64), torch.randn(e, ed[1], 3, generator=g, dtype=torch.float64))
    if e:
        x[1][0] = 0                                                            # zero vectors: the clamp branches of the norms
        ea[1][0] = 0
    return p, ei, x, ea


def conv_weights(p, prefix, n_gvp=3):
    w = []
    for l in range(n_gvp):
        k = f"{prefix}{l}."
        w += [p.get(k + "wh.weight"), p[k + "ws.weight"], p[k + "ws.bias"], p.get(k + "wv.weight"), p.get(k + "wsv.weight"),
              p.get(k + "wsv.bias")]
    return w


def conv_program(nd, ed, aggr, edge_sorted=False, acts=(F.relu, None), gate=True):
    conv = modules.GVPConv(nd, nd, ed, aggr=aggr, activations=acts, vector_gate=gate)
    return conv._program(edge_sorted)


def close(a, b, what, tol=TOL):
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = float((a - b).abs().max()) if b.numel() else 0.0
    assert a.shape == b.shape, f"{what}: {tuple(a.shape)} vs {tuple(b.shape)}"
    assert err <= tol * max(scale, 1e-3), f"{what}: abs err {err:.3e} at scale {scale:.3e}"


@pytest.mark.parametrize("n,e,nd,ed,aggr,hub,chunk,acts,gate", [
    (40, 300, (100, 16), (32, 1), "mean", False, 1 << 18, ("relu", None), True),      # BASELINE config-5 dims
    (40, 300, (100, 16), (32, 1), "sum", True, 64, ("relu", None), True),             # several chunks, a hub node
    (25, 90, (10, 3), (7, 2), "sum", False, 17, ("relu", None), True),                # odd dims, ragged chunks
    (25, 90, (12, 2), (5, 0), "mean", True, 1000, ("relu", None), True),              # no edge vectors
    (25, 90, (12, 2), (0, 1), "sum", False, 31, ("relu", None), True),                # no edge scalars
    (30, 120, (9, 4), (6, 1), "mean", False, 50, ("relu", "sigmoid"), False),         # reference defaults: no gate, sigmoid on the norms
    (30, 120, (9, 4), (6, 1), "sum", False, 50, ("sigmoid", "relu"), True),           # gate fed through an activation
    (12, 0, (100, 16), (32, 1), "mean", False, 64, ("relu", None), True),             # empty graph
])
def test_conv_forward_backward_match_oracle(n, e, nd, ed, aggr, hub, chunk, acts, gate, monkeypatch):
    monkeypatch.setattr(wide, "CHUNK_EDGES", chunk)
    p, ei, x, ea = layer_case(n, e, nd, ed, seed=n + e + nd[0], hub=hub, isolated=True)
    act_fn = {"relu": F.relu, "sigmoid": torch.sigmoid, None: None}
    prog = conv_program(nd, ed, aggr, acts=(act_fn[acts[0]], act_fn[acts[1]]), gate=gate)
    if not gate:
        p = {k: v for k, v in p.items() if ".wsv." not in k}
    assert wide.conv_supported(prog)
    plan = cpu_plan(ei, n)
    w = conv_weights(p, "conv.message_func.")
    out_s, out_v = wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w)
    leaves = [t.clone().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
    ref = gvp_oracle.gvp_conv(pl, "conv.", (leaves[0], leaves[1]), ei, (leaves[2], leaves[3]), aggr=aggr,
                              scalar_act=acts[0], vector_act=acts[1], vector_gate=gate)
    close(out_s, ref[0].detach(), "out_s")
    close(out_v, ref[1].detach(), "out_v")
    g = torch.Generator().manual_seed(1)
    cs, cv = torch.randn(ref[0].shape, generator=g, dtype=torch.float64), torch.randn(ref[1].shape, generator=g, dtype=torch.float64)
    ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
    d_x_s, d_x_v, d_e_s, d_e_v, dw = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv)
    zero = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
    for got, leaf, name in zip((d_x_s, d_x_v, d_e_s, d_e_v), leaves, ("d_x_s", "d_x_v", "d_e_s", "d_e_v")):
        close(got, zero(leaf), name)
    names = ("wh.weight", "ws.weight", "ws.bias", "wv.weight", "wsv.weight", "wsv.bias")
    for l in range(3):
        for j, nm in enumerate(names):
            key = f"conv.message_func.{l}.{nm}"
            if key in pl:
                close(dw[6 * l + j], zero(pl[key]), "grad " + key)
            else:
                assert dw[6 * l + j] is None


def test_conv_edge_sorted_input_order():
    """`edge_sorted`: edge attributes (and their gradient) already in dst-sorted order, as the LBA encoder feeds them."""
    n, e, nd, ed = 30, 200, (16, 4), (8, 1)
    p, ei, x, ea = layer_case(n, e, nd, ed, seed=5)
    plan = cpu_plan(ei, n)
    perm = plan.perm.long()
    w = conv_weights(p, "conv.message_func.")
    a = wide.conv_forward(conv_program(nd, ed, "sum", False), plan, x[0], x[1], ea[0], ea[1], w)
    b = wide.conv_forward(conv_program(nd, ed, "sum", True), plan, x[0], x[1], ea[0][perm], ea[1][perm], w)
    close(b[0], a[0], "out_s")
    close(b[1], a[1], "out_v")
    cs, cv = torch.randn_like(a[0]), torch.randn_like(a[1])
    ga = wide.conv_backward(conv_program(nd, ed, "sum", False), plan, x[0], x[1], ea[0], ea[1], w, cs, cv)
    gb = wide.conv_backward(conv_program(nd, ed, "sum", True), plan, x[0], x[1], ea[0][perm], ea[1][perm], w, cs, cv)
    close(gb[0], ga[0], "d_x_s")
    close(gb[1], ga[1], "d_x_v")
    close(gb[2], ga[2][perm], "d_e_s")
    close(gb[3], ga[3][perm], "d_e_v")
    for u, v in zip(ga[4], gb[4]):
        if u is not None:
            close(v, u, "weight grad")


@pytest.mark.parametrize("nd,drop", [((100, 16), True), ((10, 3), False), ((12, 2), True)])
def test_node_update_backward_matches_oracle(nd, drop):
    """GVPConvLayer node update x <- LN1(x1 + D1(FF(x1))), x1 = LN0(x + D0(dh)) (`gvp_layers.py:407-410`): forward and every
    gradient (inputs, residual addend, LayerNorm parameters, the two feed-forward GVPs) with fixed dropout masks."""
    n = 37
    g = torch.Generator().manual_seed(nd[0])
    p = gvp_oracle.init_conv_layer_params({}, "", nd, (4, 1), gen=g, dtype=torch.float64)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x, dh = (r(n, nd[0]), r(n, nd[1], 3)), (r(n, nd[0]), r(n, nd[1], 3))
    x[1][0] = 0
    dh[1][0] = 0
    masks = (None, None)
    if drop:
        mk = lambda *s: (torch.rand(*s, generator=g) < 0.8).double() / 0.8
        masks = ((mk(n, nd[0]), mk(n, nd[1])), (mk(n, nd[0]), mk(n, nd[1])))
    layer = modules.GVPConvLayer(nd, (4, 1), drop_rate=0.2, activations=(F.relu, None), vector_gate=True)
    prog = modules._row_program(nd[0], nd[1], tuple(m.spec for m in layer.ff_func), residual_in=True, pre_norm=True,
                                post_residual=True, post_norm=True)
    w = conv_weights(p, "ff_func.", 2)
    t = dict(in_s=x[0], in_v=x[1], h_s=dh[0], h_v=dh[1], in_index=None, types=None,
             mask0_s=None if masks[0] is None else masks[0][0], mask0_v=None if masks[0] is None else masks[0][1],
             mask1_s=None if masks[1] is None else masks[1][0], mask1_v=None if masks[1] is None else masks[1][1],
             ln0_w=p["norm.0.scalar_norm.weight"], ln0_b=p["norm.0.scalar_norm.bias"],
             ln1_w=p["norm.1.scalar_norm.weight"], ln1_b=p["norm.1.scalar_norm.bias"])
    assert wide.rows_supported(prog, t)
    # oracle: the same node update written with its functions
    leaves = [q.clone().requires_grad_() for q in (x[0], x[1], dh[0], dh[1])]
    pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
    d0 = gvp_oracle.dropout((leaves[2], leaves[3]), masks[0])
    x1 = gvp_oracle.layer_norm(pl, "norm.0.", (leaves[0] + d0[0], leaves[1] + d0[1]))
    d1 = gvp_oracle.dropout(gvp_oracle.feed_forward(pl, "", x1, 2, "relu", None, True), masks[1])
    ref = gvp_oracle.layer_norm(pl, "norm.1.", (x1[0] + d1[0], x1[1] + d1[1]))
    out_s, out_v = wide.rows_forward(prog, t, w)
    close(out_s, ref[0].detach(), "out_s")
    close(out_v, ref[1].detach(), "out_v")
    cs, cv = r(n, nd[0]), r(n, nd[1], 3)
    ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
    got = wide.rows_backward(prog, t, w, cs, cv)
    for key, leaf in zip(("d_in_s", "d_in_v", "d_h_s", "d_h_v"), leaves):
        close(got[key], leaf.grad, key)
    for i, key in enumerate(("norm.0.scalar_norm.weight", "norm.0.scalar_norm.bias", "norm.1.scalar_norm.weight",
                             "norm.1.scalar_norm.bias")):
        close(got["ln"][i], pl[key].grad, "grad " + key)
    names = ("wh.weight", "ws.weight", "ws.bias", "wv.weight", "wsv.weight", "wsv.bias")
    for l in range(2):
        for j, nm in enumerate(names):
            close(got["dw"][6 * l + j], pl[f"ff_func.{l}.{nm}"].grad, f"grad ff_func.{l}.{nm}")


def test_scalar_output_row_program_matches_oracle():
    """LayerNorm + GVP (ns, nv) -> (out, 0): the read-out stage (`protein_gnn.py:385-386`) at wide dims."""
    n, nd, out = 29, (20, 5), 12
    g = torch.Generator().manual_seed(2)
    p = {}
    gvp_oracle.init_layer_norm_params(p, "ln.", nd[0], g, torch.float64)
    gvp_oracle.init_gvp_params(p, "g.", nd, (out, 0), vector_gate=True, gen=g, dtype=torch.float64)
    x = (torch.randn(n, nd[0], generator=g, dtype=torch.float64), torch.randn(n, nd[1], 3, generator=g, dtype=torch.float64))
    gv = modules.GVP(nd, (out, 0), activations=(F.relu, None), vector_gate=True)
    prog = modules._row_program(nd[0], nd[1], (gv.spec,), pre_norm=True)
    w = [p["g.wh.weight"], p["g.ws.weight"], p["g.ws.bias"], None, None, None]
    t = dict(in_s=x[0], in_v=x[1], in_index=None, types=None, ln0_w=p["ln.scalar_norm.weight"], ln0_b=p["ln.scalar_norm.bias"])
    leaves = [q.clone().requires_grad_() for q in x]
    pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
    ref = gvp_oracle.gvp(pl, "g.", gvp_oracle.layer_norm(pl, "ln.", (leaves[0], leaves[1])), "relu", None, True)
    out_s, out_v = wide.rows_forward(prog, t, w)
    assert out_v is None
    close(out_s, ref.detach(), "out")
    cs = torch.randn(n, out, generator=g, dtype=torch.float64)
    (ref * cs).sum().backward()
    got = wide.rows_backward(prog, t, w, cs, None)
    close(got["d_in_s"], leaves[0].grad, "d_in_s")
    close(got["d_in_v"], leaves[1].grad, "d_in_v")
    close(got["ln"][0], pl["ln.scalar_norm.weight"].grad, "d_ln_w")
    close(got["ln"][1], pl["ln.scalar_norm.bias"].grad, "d_ln_b")
    close(got["dw"][0], pl["g.wh.weight"].grad, "d_wh")
    close(got["dw"][1], pl["g.ws.weight"].grad, "d_ws")
    close(got["dw"][2], pl["g.ws.bias"].grad, "d_bs")


def test_there_is_no_cpu_path(monkeypatch):
    """The product primitive refuses CPU tensors: the substitution above is test infrastructure only."""
    monkeypatch.undo()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wide._segsum(torch.zeros(4, 4), torch.zeros(3, dtype=torch.int32), None, 2)


# ---- autograd wiring (ops.ConvFunction / ops.RowsFunction route wide descriptors here) -----------------------------------------
def _permissive_f32(t):
    return None if t is None else t.detach().contiguous()


def test_gvpconv_module_routes_wide_dims_through_the_gemm_formulation(monkeypatch):
    """`GVPConv.forward` + autograd at config-5 dims: outputs and every gradient through `ops.ConvFunction` equal the oracle
    (host pointers never reach the C ABI: the primitive and the CUDA-only input check are substituted)."""
    monkeypatch.setattr(ops, "_f32", _permissive_f32)
    monkeypatch.setattr(wide, "MIN_DIM", 64)
    n, e, nd, ed = 33, 260, (100, 16), (32, 1)
    p, ei, x, ea = layer_case(n, e, nd, ed, seed=11, hub=True)
    conv = modules.GVPConv(nd, nd, ed, aggr="mean", activations=(F.relu, None), vector_gate=True).double()
    conv.load_state_dict({k[len("conv."):]: v for k, v in p.items() if k.startswith("conv.")}, strict=True)
    leaves = [t.clone().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    out = conv((leaves[0], leaves[1]), ei, (leaves[2], leaves[3]), plan=cpu_plan(ei, n))
    ref_leaves = [t.clone().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
    ref = gvp_oracle.gvp_conv(pl, "conv.", (ref_leaves[0], ref_leaves[1]), ei, (ref_leaves[2], ref_leaves[3]), aggr="mean",
                              scalar_act="relu", vector_act=None, vector_gate=True)
    close(out[0].detach(), ref[0].detach(), "out_s")
    close(out[1].detach(), ref[1].detach(), "out_v")
    cs, cv = torch.randn_like(ref[0]), torch.randn_like(ref[1])
    ((out[0] * cs).sum() + (out[1] * cv).sum()).backward()
    ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
    for a, b, name in zip(leaves, ref_leaves, ("x_s", "x_v", "e_s", "e_v")):
        close(a.grad, b.grad, "grad " + name)
    for name, prm in conv.named_parameters():
        if prm.numel():
            close(prm.grad, pl["conv." + name].grad, "grad " + name)


def test_rows_function_backward_routes_wide_node_update(monkeypatch):
    """`ops.RowsFunction.backward` hands a wide node-update program to `wide.rows_backward` and returns the gradients in the
    positions of `RowsFunction.forward`'s arguments."""
    monkeypatch.setattr(ops, "_f32", _permissive_f32)
    monkeypatch.setattr(wide, "MIN_DIM", 64)
    n, nd = 21, (100, 16)
    g = torch.Generator().manual_seed(4)
    p = gvp_oracle.init_conv_layer_params({}, "", nd, (4, 1), gen=g, dtype=torch.float64)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    layer = modules.GVPConvLayer(nd, (4, 1), drop_rate=0.0, activations=(F.relu, None), vector_gate=True)
    prog = modules._row_program(nd[0], nd[1], tuple(m.spec for m in layer.ff_func), residual_in=True, pre_norm=True,
                                post_residual=True, post_norm=True)
    w = conv_weights(p, "ff_func.", 2)
    t = dict(in_s=r(n, nd[0]), in_v=r(n, nd[1], 3), h_s=r(n, nd[0]), h_v=r(n, nd[1], 3), in_index=None, types=None,
             mask0_s=None, mask0_v=None, mask1_s=None, mask1_v=None, stash=None,
             ln0_w=p["norm.0.scalar_norm.weight"], ln0_b=p["norm.0.scalar_norm.bias"],
             ln1_w=p["norm.1.scalar_norm.weight"], ln1_b=p["norm.1.scalar_norm.bias"])
    n_args = 15 + len(w)           # prog, in_s, in_v, types, in_index, h_s, h_v, 4 masks, 4 LayerNorm tensors, weights
    ctx = types.SimpleNamespace(prog=prog, t=t, rows=n, weights=w, in_rows=n, needs_input_grad=[True] * n_args)
    cs, cv = r(n, nd[0]), r(n, nd[1], 3)
    got = ops.RowsFunction.backward(ctx, cs, cv)
    want = wide.rows_backward(prog, t, w, cs, cv)
    assert len(got) == n_args
    assert got[0] is None and got[3] is None and got[4] is None and all(x is None for x in got[7:11])
    for pos, key in ((1, "d_in_s"), (2, "d_in_v"), (5, "d_h_s"), (6, "d_h_v")):
        assert torch.equal(got[pos], want[key]), key
    for i in range(4):
        assert torch.equal(got[11 + i], want["ln"][i])
    for i in range(len(w)):
        assert torch.equal(got[15 + i], want["dw"][i])


def test_blocked_weight_gradient_product():
    """`_tdot`: the blocked (batched-GEMM) a^T b equals the plain product, ragged tail included."""
    g = torch.Generator().manual_seed(0)
    for r in (3, 64, 70, 127):
        a, b = torch.randn(r, 5, generator=g, dtype=torch.float64), torch.randn(r, 7, generator=g, dtype=torch.float64)
        close(wide._tdot(a, b), a.t() @ b, f"rows {r}", tol=1e-13)
    a = torch.randn(100, 12, generator=g, dtype=torch.float64)
    close(wide._tdot(a[:, 2:9], a[:, :5]), a[:, 2:9].t() @ a[:, :5], "strided operands", tol=1e-13)


def test_layer_against_the_reference_fixture_at_config5_dims():
    """`tests/golden/layer_wide.npz`: ONE GVPConvLayer at nodes (100,16) / edges (32,1) evaluated by the UNMODIFIED reference
    (make_golden.py).  The GEMM formulation, composed the way autograd composes it (conv -> node update; node update backward
    -> conv backward), reproduces the reference's outputs and every gradient in fp64."""
    c = case(golden("layer_wide"), "layer_wide")
    nd, ed = (100, 16), (32, 1)
    p = {k: v.double() for k, v in c["param"].items()}
    x_s, x_v, e_s, e_v = (c[k].double() for k in ("s", "v", "es", "ev"))
    ei = c["edge_index"]
    n = x_s.shape[0]
    plan = cpu_plan(ei, n)
    layer = modules.GVPConvLayer(nd, ed, drop_rate=0.0, activations=(F.relu, None), vector_gate=True, aggr="mean")
    cprog = layer.conv._program(False)
    rprog = modules._row_program(nd[0], nd[1], tuple(m.spec for m in layer.ff_func), residual_in=True, pre_norm=True,
                                 post_residual=True, post_norm=True)
    cw, rw = conv_weights(p, "conv.message_func."), conv_weights(p, "ff_func.", 2)
    dh = wide.conv_forward(cprog, plan, x_s, x_v, e_s, e_v, cw)
    t = dict(in_s=x_s, in_v=x_v, h_s=dh[0], h_v=dh[1], in_index=None, types=None, mask0_s=None, mask0_v=None, mask1_s=None,
             mask1_v=None, ln0_w=p["norm.0.scalar_norm.weight"], ln0_b=p["norm.0.scalar_norm.bias"],
             ln1_w=p["norm.1.scalar_norm.weight"], ln1_b=p["norm.1.scalar_norm.bias"])
    out_s, out_v = wide.rows_forward(rprog, t, rw)
    close(out_s, c["out_s"], "out_s")
    close(out_v, c["out_v"], "out_v")
    r = wide.rows_backward(rprog, t, rw, c["cot_s"].double(), c["cot_v"].double())
    g = wide.conv_backward(cprog, plan, x_s, x_v, e_s, e_v, cw, r["d_h_s"], r["d_h_v"])
    close(r["d_in_s"] + g[0], c["grad_s"], "grad_s")
    close(r["d_in_v"] + g[1], c["grad_v"], "grad_v")
    close(g[2], c["grad_es"], "grad_es")
    close(g[3], c["grad_ev"], "grad_ev")
    names = ("wh.weight", "ws.weight", "ws.bias", "wv.weight", "wsv.weight", "wsv.bias")
    for prefix, grads, count in (("conv.message_func.", g[4], 3), ("ff_func.", r["dw"], 2)):
        for l in range(count):
            for j, nm in enumerate(names):
                close(grads[6 * l + j], c["grad_param"][f"{prefix}{l}.{nm}"], f"grad {prefix}{l}.{nm}")
    for i, key in enumerate(("norm.0.scalar_norm.weight", "norm.0.scalar_norm.bias", "norm.1.scalar_norm.weight",
                             "norm.1.scalar_norm.bias")):
        close(r["ln"][i], c["grad_param"][key], "grad " + key)


def test_plain_gvp_and_layernorm_programs_match_oracle():
    """The stand-alone `GVP` and `LayerNorm` modules at wide dims are row programs too (one GVP, no norms / only the norm)."""
    n, nd = 23, (20, 5)
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(n, nd[0], generator=g, dtype=torch.float64), torch.randn(n, nd[1], 3, generator=g, dtype=torch.float64))
    x[1][2] = 0
    cs, cv = torch.randn(n, 9, generator=g, dtype=torch.float64), torch.randn(n, 4, 3, generator=g, dtype=torch.float64)
    # GVP (20,5) -> (9,4), reference defaults: (relu, sigmoid), no gate
    p = gvp_oracle.init_gvp_params({}, "g.", nd, (9, 4), vector_gate=False, gen=g, dtype=torch.float64)
    gv = modules.GVP(nd, (9, 4))
    prog = modules._row_program(nd[0], nd[1], (gv.spec,))
    w = [p["g.wh.weight"], p["g.ws.weight"], p["g.ws.bias"], p["g.wv.weight"], None, None]
    t = dict(in_s=x[0], in_v=x[1], in_index=None, types=None)
    assert wide.rows_supported(prog, t)
    leaves = [q.clone().requires_grad_() for q in x]
    pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
    ref = gvp_oracle.gvp(pl, "g.", (leaves[0], leaves[1]), "relu", "sigmoid", False)
    out = wide.rows_forward(prog, t, w)
    close(out[0], ref[0].detach(), "gvp s")
    close(out[1], ref[1].detach(), "gvp V")
    ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
    got = wide.rows_backward(prog, t, w, cs, cv)
    close(got["d_in_s"], leaves[0].grad, "gvp d_in_s")
    close(got["d_in_v"], leaves[1].grad, "gvp d_in_v")
    for j, nm in enumerate(("wh.weight", "ws.weight", "ws.bias", "wv.weight")):
        close(got["dw"][j], pl["g." + nm].grad, "gvp grad " + nm)
    assert got["dw"][4] is None and got["dw"][5] is None and got["d_h_s"] is None
    # LayerNorm only
    p = gvp_oracle.init_layer_norm_params({}, "ln.", nd[0], g, torch.float64)
    prog = modules._row_program(nd[0], nd[1], (), pre_norm=True)
    t = dict(in_s=x[0], in_v=x[1], in_index=None, types=None, ln0_w=p["ln.scalar_norm.weight"], ln0_b=p["ln.scalar_norm.bias"])
    assert wide.rows_supported(prog, t)
    leaves = [q.clone().requires_grad_() for q in x]
    pl = {k: v.clone().requires_grad_() for k, v in p.items()}
    ref = gvp_oracle.layer_norm(pl, "ln.", (leaves[0], leaves[1]))
    out = wide.rows_forward(prog, t, [])
    close(out[0], ref[0].detach(), "ln s")
    close(out[1], ref[1].detach(), "ln V")
    cs, cv = torch.randn_like(x[0]), torch.randn_like(x[1])
    ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
    got = wide.rows_backward(prog, t, [], cs, cv)
    close(got["d_in_s"], leaves[0].grad, "ln d_in_s")
    close(got["d_in_v"], leaves[1].grad, "ln d_in_v")
    close(got["ln"][0], pl["ln.scalar_norm.weight"].grad, "ln d_w")
    close(got["ln"][1], pl["ln.scalar_norm.bias"].grad, "ln d_b")


def test_conv_random_dims_sweep(monkeypatch):
    """Seeded sweep over random dims / graphs / chunk sizes / aggregations: forward and every gradient against the oracle."""
    rng = torch.Generator().manual_seed(2024)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    for trial in range(12):
        nd, ed = (ri(1, 24), ri(1, 6)), (ri(0, 9), ri(0, 3))
        n, e = ri(5, 40), ri(1, 200)
        aggr = ("mean", "sum")[trial % 2]
        monkeypatch.setattr(wide, "CHUNK_EDGES", ri(1, 64))
        monkeypatch.setattr(wide, "WGRAD_BLOCK", ri(2, 32))
        p, ei, x, ea = layer_case(n, e, nd, ed, seed=1000 + trial, hub=trial % 3 == 0, isolated=n > 6)
        prog = conv_program(nd, ed, aggr)
        assert wide.conv_supported(prog)
        plan = cpu_plan(ei, n)
        w = conv_weights(p, "conv.message_func.")
        out = wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w)
        leaves = [t.clone().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
        pl = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
        ref = gvp_oracle.gvp_conv(pl, "conv.", (leaves[0], leaves[1]), ei, (leaves[2], leaves[3]), aggr=aggr, scalar_act="relu",
                                  vector_act=None, vector_gate=True)
        tag = f"trial {trial} nd={nd} ed={ed} n={n} e={e} {aggr}"
        close(out[0], ref[0].detach(), tag + " out_s")
        close(out[1], ref[1].detach(), tag + " out_v")
        cs, cv = torch.randn(ref[0].shape, generator=rng, dtype=torch.float64), torch.randn(ref[1].shape, generator=rng, dtype=torch.float64)
        ((ref[0] * cs).sum() + (ref[1] * cv).sum()).backward()
        got = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv)
        zero = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
        for g_, leaf, name in zip(got[:4], leaves, ("d_x_s", "d_x_v", "d_e_s", "d_e_v")):
            close(g_, zero(leaf), tag + " " + name)
        names = ("wh.weight", "ws.weight", "ws.bias", "wv.weight", "wsv.weight", "wsv.bias")
        for l in range(3):
            for j, nm in enumerate(names):
                close(got[4][6 * l + j], zero(pl[f"conv.message_func.{l}.{nm}"]), f"{tag} grad {l}.{nm}")


def test_forward_stash_for_the_backward(monkeypatch):
    """Training: the forward leaves its per-chunk intermediates when they fit `STASH_BYTES`; the backward then gives the same
    gradients as with its own recompute (same arithmetic: same bits).  Over budget the list stays empty."""
    monkeypatch.setattr(wide, "CHUNK_EDGES", 64)
    n, e, nd, ed = 40, 300, (100, 16), (32, 1)
    p, ei, x, ea = layer_case(n, e, nd, ed, seed=77, hub=True)
    prog, plan, w = conv_program(nd, ed, "mean"), cpu_plan(ei, n), conv_weights(p, "conv.message_func.")
    cs, cv = torch.randn(n, nd[0], dtype=torch.float64), torch.randn(n, nd[1], 3, dtype=torch.float64)
    ref = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv)
    kept = []
    out = wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w, kept=kept)
    assert len(kept) == 5 and torch.equal(out[0], wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w)[0])
    got = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv, kept=kept)
    assert all(k is None for k in kept), "chunks are released as the backward consumes them"
    for a, b in zip(got[:4], ref[:4]):
        assert torch.equal(a, b)
    for a, b in zip(got[4], ref[4]):
        assert (a is None and b is None) or torch.equal(a, b)
    monkeypatch.setattr(wide, "STASH_BYTES", 1000)
    kept = []
    wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w, kept=kept)
    assert kept == []
    got = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv, kept=kept)     # empty list: recompute
    assert torch.equal(got[0], ref[0])


def test_properties_rotation_permutation_mean():
    """SURVEY.md section 4 (iii) for the GEMM formulation: scalar outputs invariant and vector outputs equivariant under a random
    orthogonal transform (`protein_gnn.py:362`), invariance under a permutation of the edge order, mean = sum / in-degree."""
    n, e, nd, ed = 30, 240, (100, 16), (32, 1)
    p, ei, x, ea = layer_case(n, e, nd, ed, seed=31, hub=True, isolated=True)
    w = conv_weights(p, "conv.message_func.")
    prog_sum, prog_mean = conv_program(nd, ed, "sum"), conv_program(nd, ed, "mean")
    plan = cpu_plan(ei, n)
    base = wide.conv_forward(prog_sum, plan, x[0], x[1], ea[0], ea[1], w)
    # rotation (+ reflection): V -> V Q^T
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)))
    rot = wide.conv_forward(prog_sum, plan, x[0], x[1] @ q.t(), ea[0], ea[1] @ q.t(), w)
    close(rot[0], base[0], "scalars are invariant")
    close(rot[1], base[1] @ q.t(), "vectors are equivariant")
    # edge permutation
    perm = torch.randperm(e, generator=torch.Generator().manual_seed(2))
    ei2 = ei[:, perm]
    per = wide.conv_forward(prog_sum, cpu_plan(ei2, n), x[0], x[1], ea[0][perm], ea[1][perm], w)
    close(per[0], base[0], "edge order (s)")
    close(per[1], base[1], "edge order (V)")
    # mean vs sum
    deg = torch.bincount(ei[1], minlength=n).clamp(min=1).double()
    mean = wide.conv_forward(prog_mean, plan, x[0], x[1], ea[0], ea[1], w)
    close(mean[0], base[0] / deg.unsqueeze(1), "mean = sum / deg (s)")
    close(mean[1], base[1] / deg.view(-1, 1, 1), "mean = sum / deg (V)")
    assert float(base[0][-2:].abs().max()) == 0.0 and float(base[1][-2:].abs().max()) == 0.0, "nodes without in-edges get zeros"
