"""GPU parity of the wide-dims (BASELINE config 5: nodes (100,16), edges (32,1)) training path: the GEMM formulation
(`caster_dta_b200/wide.py`: chunked library GEMMs + C-ABI segmented reductions) against the fp64 oracle AND against the
generic shared-memory tile kernels of the same library, forward and every gradient.  `tests/test_wide_gemm_cpu.py` checks the
same algebra on the CPU in fp64 at 1e-10."""
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close, assert_param_grads_close, assert_rows_close, case, golden

pytestmark = pytest.mark.gpu

TOL = 1e-4
DEV = "cuda"


@pytest.fixture(params=["gemm", "tile"])
def wide_mode(request):
    from caster_dta_b200 import wide
    prev = wide.ENABLED
    wide.set_enabled(request.param == "gemm")
    yield request.param
    wide.set_enabled(prev)


def _case(n, e, nd, ed, seed, hub):
    from oracle import gvp_oracle
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - 3, (e,), generator=g)          # the last three nodes have no in-edges
    if hub:
        dst[: e // 3] = 7
    order = torch.argsort(src * n + dst, stable=True)
    ei = torch.stack([src[order], dst[order]])
    p = gvp_oracle.init_conv_layer_params({}, "", nd, ed, gen=g)
    x = (torch.randn(n, nd[0], generator=g), torch.randn(n, nd[1], 3, generator=g))
    ea = (torch.randn(e, ed[0], generator=g), torch.randn(e, ed[1], 3, generator=g))
    x[1][5] = 0                                               # a zero vector row: the clamp branches of the norms
    return p, ei, x, ea


def _run_layer(p, ei, x, ea, nd, ed, aggr, drop=0.0, seed=None):
    import caster_dta_b200 as cg
    m = cg.GVPConvLayer(nd, ed, drop_rate=drop, activations=(F.relu, None), vector_gate=True, aggr=aggr)
    m.load_state_dict(p, strict=True)
    m.to(DEV).train()
    leaves = [t.float().to(DEV).requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    if seed is not None:
        torch.manual_seed(seed)
    out = m((leaves[0], leaves[1]), ei.to(DEV), (leaves[2], leaves[3]))
    return m, leaves, out


@pytest.mark.parametrize("n,e,aggr,hub,chunk", [(700, 9000, "mean", False, 1 << 18), (611, 7001, "sum", True, 2048)])
def test_wide_layer_forward_backward_vs_oracle(n, e, aggr, hub, chunk, wide_mode):
    from caster_dta_b200 import wide
    from oracle import gvp_oracle
    nd, ed = (100, 16), (32, 1)
    p, ei, x, ea = _case(n, e, nd, ed, seed=n + e, hub=hub)
    prev, wide.CHUNK_EDGES = wide.CHUNK_EDGES, chunk
    try:
        m, leaves, out = _run_layer(p, ei, x, ea, nd, ed, aggr)
        g = torch.Generator().manual_seed(1)
        cs, cv = torch.randn(out[0].shape, generator=g), torch.randn(out[1].shape, generator=g)
        ((out[0] * cs.to(DEV)).sum() + (out[1] * cv.to(DEV)).sum()).backward()
        torch.cuda.synchronize()
    finally:
        wide.CHUNK_EDGES = prev
    p64 = {k: v.double().requires_grad_(v.numel() > 0) for k, v in p.items()}
    l64 = [t.double().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    ref = gvp_oracle.gvp_conv_layer(p64, "", (l64[0], l64[1]), ei, (l64[2], l64[3]), aggr=aggr, scalar_act="relu",
                                    vector_act=None, vector_gate=True)
    ((ref[0] * cs.double()).sum() + (ref[1] * cv.double()).sum()).backward()
    assert_close(out[0], ref[0], TOL, f"s [{wide_mode}]")
    assert_close(out[1], ref[1], TOL, f"V [{wide_mode}]")
    for t, r, k in zip(leaves, l64, ("grad_s", "grad_v", "grad_es", "grad_ev")):
        assert_rows_close(t.grad, r.grad, TOL, f"{k} [{wide_mode}]", atol=1e-6)
    assert_param_grads_close({k: q.grad for k, q in m.named_parameters() if q.numel()},
                             {k: p64[k].grad for k, q in m.named_parameters() if q.numel()}, f"[{wide_mode}]")


def test_wide_gemm_and_tile_paths_agree_with_dropout():
    """Train mode with dropout 0.2 (same seeded torch RNG stream for both): the GEMM formulation and the generic tile
    kernels give the same outputs and gradients to fp32 round-off, and the GEMM formulation is bit-reproducible."""
    from caster_dta_b200 import wide
    nd, ed = (100, 16), (32, 1)
    p, ei, x, ea = _case(500, 6000, nd, ed, seed=3, hub=True)
    res = {}
    prev = wide.ENABLED
    try:
        for mode in ("gemm", "gemm2", "tile"):
            wide.set_enabled(mode != "tile")
            m, leaves, out = _run_layer(p, ei, x, ea, nd, ed, "mean", drop=0.2, seed=5)
            (out[0].square().sum() + out[1].square().sum()).backward()
            res[mode] = [out[0].detach(), out[1].detach()] + [t.grad for t in leaves] + [q.grad for q in m.parameters() if q.numel()]
    finally:
        wide.set_enabled(prev)
    for a, b in zip(res["gemm"], res["gemm2"]):
        assert torch.equal(a, b), "the GEMM formulation must be bit-reproducible"
    for i in range(2):
        assert_close(res["gemm"][i], res["tile"][i].cpu(), TOL, ("s", "V")[i], atol=1e-6)
    for i, k in enumerate(("grad_s", "grad_v", "grad_es", "grad_ev")):
        assert_rows_close(res["gemm"][2 + i], res["tile"][2 + i], TOL, k, atol=1e-6)
    assert_param_grads_close({i: t for i, t in enumerate(res["gemm"][6:])}, {i: t for i, t in enumerate(res["tile"][6:])}, "gemm vs tile")


def test_wide_conv_tensor_core_mode_backward():
    """<= 1e-2 mode (`set_tensor_cores(True)`): tcgen05 bf16 forward + TF32 library GEMMs in the backward, against the oracle."""
    import caster_dta_b200 as cg
    from caster_dta_b200 import _lib, wide
    from oracle import gvp_oracle
    nd, ed = (100, 16), (32, 1)
    p, ei, x, ea = _case(900, 27000, nd, ed, seed=8, hub=False)
    conv = cg.GVPConv(nd, nd, ed, aggr="mean", activations=(F.relu, None), vector_gate=True)
    conv.load_state_dict({k[len("conv."):]: v for k, v in p.items() if k.startswith("conv.")}, strict=True)
    conv.to(DEV)
    leaves = [t.float().to(DEV).requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    prev = wide.ENABLED
    wide.set_enabled(True)
    _lib.set_tensor_cores(True)
    try:
        out = conv((leaves[0], leaves[1]), ei.to(DEV), (leaves[2], leaves[3]))
        (out[0].sum() + out[1].sum()).backward()
        torch.cuda.synchronize()
    finally:
        _lib.set_tensor_cores(False)
        wide.set_enabled(prev)
    p64 = {k: v.double().requires_grad_(v.numel() > 0) for k, v in p.items()}
    l64 = [t.double().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
    ref = gvp_oracle.gvp_conv(p64, "conv.", (l64[0], l64[1]), ei, (l64[2], l64[3]), aggr="mean", scalar_act="relu",
                              vector_act=None, vector_gate=True)
    (ref[0].sum() + ref[1].sum()).backward()
    assert_close(out[0], ref[0], 1e-2, "s")
    assert_close(out[1], ref[1], 1e-2, "V")
    # Gradients: the backward differentiates its own TF32 recompute of the message chain.  At 2^-11 operand precision a ReLU
    # pre-activation near zero takes the other branch on ~1e-3 of the elements, i.e. in a noticeable share of the 200-element
    # edge rows (first GPU run: 1.0 % of the d(edge scalars) rows beyond 3e-2, profiles/r2_wide_check_first_gpu_run.log; a CPU
    # emulation that rounds every GEMM operand to TF32 gives 1.04 % on this very case, with L2-relative errors of 0.6-2.5 % for
    # the per-row gradients and <= 0.11 % for the parameter gradients).  Aggregate bounds, ~3x / ~18x above those: per-row
    # gradients within 8 % in L2, parameter gradients within 2 %.  The fp32 mode carries the tight per-row bounds (tests above).
    def l2_close(a, b, what, bound):
        a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
        assert torch.isfinite(a64).all(), f"{what}: non-finite values"
        rel = float((a64 - b64).norm() / b64.norm().clamp(min=1e-30))
        assert rel <= bound, f"{what}: L2-relative error {rel:.3e} > {bound}"

    for t, r, k in zip(leaves, l64, ("grad_s", "grad_v", "grad_es", "grad_ev")):
        l2_close(t.grad, r.grad, k, 0.08)
    for name, prm in conv.named_parameters():
        if prm.numel():
            l2_close(prm.grad, p64["conv." + name].grad, "grad " + name, 0.02)


def test_wide_layer_against_the_reference_fixture(wide_mode):
    """`tests/golden/layer_wide.npz`: one GVPConvLayer at config-5 dims evaluated by the unmodified reference in fp64."""
    import caster_dta_b200 as cg
    c = case(golden("layer_wide"), "layer_wide")
    nd, ed = (100, 16), (32, 1)
    m = cg.GVPConvLayer(nd, ed, drop_rate=0.0, activations=(F.relu, None), vector_gate=True, aggr="mean")
    m.load_state_dict({k: v.float() for k, v in c["param"].items()}, strict=True)
    m.to(DEV).eval()
    leaves = [c[k].float().to(DEV).requires_grad_() for k in ("s", "v", "es", "ev")]
    out = m((leaves[0], leaves[1]), c["edge_index"].to(DEV), (leaves[2], leaves[3]))
    assert_close(out[0], c["out_s"], TOL, f"s [{wide_mode}]")
    assert_close(out[1], c["out_v"], TOL, f"V [{wide_mode}]")
    ((out[0] * c["cot_s"].float().to(DEV)).sum() + (out[1] * c["cot_v"].float().to(DEV)).sum()).backward()
    for t, k in zip(leaves, ("grad_s", "grad_v", "grad_es", "grad_ev")):
        assert_rows_close(t.grad, c[k], TOL, f"{k} [{wide_mode}]", atol=1e-6)
    named = {k: q.grad for k, q in m.named_parameters() if q.numel()}
    assert_param_grads_close(named, {k: c["grad_param"][k] for k in named}, f"[{wide_mode}]")
