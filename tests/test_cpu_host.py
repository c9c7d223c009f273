"""CPU-side checks: the C-ABI library loads and exports every symbol of include/castergvp.h, the drop-in modules
keep the reference's state_dict contract, and the host-side helpers behave.  No kernel is launched here."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    from caster_dta_b200 import _lib
    header = open(os.path.join(ROOT, "include", "castergvp.h")).read()
    declared = set(re.findall(r"\b(cgvp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.cgvp_version() >= 100
    assert isinstance(handle.cgvp_last_error(), bytes)


def test_packed_size_query_needs_no_gpu():
    from caster_dta_b200 import ops
    sp = ops.GvpSpec(64, 9, 16, 4, 9, 1, 0, 1)          # checkpoint message GVP 0
    # fwd: wh_t 12x12 + ws_t 76x16 + wv_t 12x4 + wsv_t 20x4; bwd: 12x12 + 16x76 + 4x12 + 4x16
    assert sp.packed_floats() == (144 + 1216 + 48 + 80) + (144 + 1216 + 48 + 64)


def test_argument_errors_are_reported_not_crashed():
    from caster_dta_b200 import _lib
    import ctypes as C
    L = _lib.lib()
    d = _lib.ConvDesc()
    d.ns, d.n_gvp = 16, 0
    assert L.cgvp_conv_workspace_bytes(C.byref(d), 10, 10, 0) == -1
    assert b"n_gvp" in L.cgvp_last_error()


def test_state_dict_contract():
    """Key names / shapes of the drop-in modules equal the reference's (Appendix B of SURVEY.md)."""
    import caster_dta_b200 as cg
    import torch.nn.functional as F
    layer = cg.GVPConvLayer((16, 4), (32, 1), activations=(F.relu, None), vector_gate=True, aggr="sum")
    sd = layer.state_dict()
    assert sd["conv.message_func.0.ws.weight"].shape == (16, 73)
    assert sd["conv.message_func.0.wh.weight"].shape == (9, 9)
    assert sd["conv.message_func.2.wsv.bias"].shape == (4,)
    assert sd["ff_func.0.ws.weight"].shape == (64, 24) and sd["ff_func.1.ws.weight"].shape == (16, 72)
    assert sd["dropout.1.vdropout.dummy_param"].shape == (0,) and sd["ff_func.0.dummy_param"].shape == (0,)
    assert sd["norm.1.scalar_norm.bias"].shape == (16,)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_shipped_checkpoint_loads_strict():
    import caster_dta_b200 as cg
    root = os.path.join(REF, "pretrained_model_downstream")
    kw = json.load(open(os.path.join(root, "model_kwargs.json")))
    model = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])
    ck = [f for f in sorted(os.listdir(root)) if f.startswith("bestvalmodel")][0]
    sd = torch.load(os.path.join(root, ck), weights_only=True, map_location="cpu")
    res = cg.load_state_dict_from_checkpoint(model, sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sum(p.numel() for p in model.parameters()) == 764396
    assert sum(p.numel() for p in model.protein_gnn.parameters()) == 15117


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_module_keys_equal_live_reference():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    ref_shim.install()
    from models import gvp_layers as ref
    import caster_dta_b200 as cg
    import torch.nn.functional as F
    for kw in (dict(n_message=3, n_feedforward=2), dict(n_message=1, n_feedforward=1), dict(n_message=4, n_feedforward=3)):
        a = ref.GVPConvLayer((10, 3), (7, 2), activations=(F.relu, torch.sigmoid), vector_gate=False, **kw)
        b = cg.GVPConvLayer((10, 3), (7, 2), activations=(F.relu, torch.sigmoid), vector_gate=False, **kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(sa[k].shape == sb[k].shape for k in sa)


def test_unsupported_configurations_are_rejected_at_construction():
    import caster_dta_b200 as cg
    with pytest.raises(ValueError):
        cg.GVP((4, 2), (4, 2), activations=(torch.tanh, None))
    with pytest.raises(ValueError):
        cg.GVPConv((4, 2), (4, 2), (3, 1), aggr="max")
    with pytest.raises(ValueError):
        cg.GVPConvLayer((4, 2), (3, 1), autoregressive=True, aggr="mean")
    with pytest.raises(NotImplementedError):
        cg.SelectableProteinModelWrapper(in_channels=(17, 3), edge_dim=(32, 1), base_conv="gatv2")


def test_no_cpu_fallback():
    import caster_dta_b200 as cg
    with pytest.raises(RuntimeError):
        cg.GVP((4, 2), (4, 2))((torch.randn(3, 4), torch.randn(3, 2, 3)))
    with pytest.raises(RuntimeError):
        cg.GraphPlan(torch.zeros(2, 5, dtype=torch.long), 4)


def test_tuple_helpers():
    import caster_dta_b200.modules as m
    a = (torch.ones(3, 2), torch.ones(3, 4, 3))
    s, v = m.tuple_sum(a, a, a)
    assert float(s.sum()) == 18 and float(v.sum()) == 108
    s, v = m.tuple_cat(a, a)
    assert s.shape == (3, 4) and v.shape == (3, 8, 3)
    merged = m._merge(*a)
    s2, v2 = m._split(merged, 4)
    assert torch.equal(s2, a[0]) and torch.equal(v2, a[1])


def test_synthetic_shapes():
    from caster_dta_b200 import synth
    pb = synth.protein_batch_coords("tiny", 3, seed=1)
    n = pb["coords"].shape[0]
    assert pb["x_s"].shape == (n, 17) and pb["x_v"].shape == (n, 3, 3) and pb["ptr"][-1] == n
    assert pb["ntypes"].max() < 20
    mol = synth.molecule_batch(3, seed=1)
    assert mol["x"].shape[1] == 41 and mol["eattr"].shape[1] == 9 and mol["edge_index"].max() < mol["x"].shape[0]
    ei, nn_ = synth.conv_microbench_graph(3000, k=30)
    assert ei.shape == (2, 3000) and nn_ == 100 and bool((np.diff(ei[0]) >= 0).all())


def test_packed_cross_attention_matches_padded_formulation():
    """The packed-row cross-attention block against the padded nn.MultiheadAttention formulation of the reference
    (`models/joint_gnn.py:321-408`): outputs on real rows, attention maps everywhere (padded query rows included)."""
    import torch
    from caster_dta_b200 import joint
    torch.manual_seed(4)
    for d2 in (32, 24):                                   # same / different embedding dims (packed vs split in-proj weights)
        blk = joint.CrossAttentionModule(32, d2, 4, 0.0, True, 2, 0.0).eval()
        with torch.no_grad():
            blk.preattn_norm1.bias.normal_()
            blk.preattn_norm2.bias.normal_()
        n1, n2 = [5, 9, 2], [3, 1, 4]
        b1 = torch.repeat_interleave(torch.arange(3), torch.tensor(n1))
        b2 = torch.repeat_interleave(torch.arange(3), torch.tensor(n2))
        x1, x2 = torch.randn(sum(n1), 32), torch.randn(sum(n2), d2)
        e1, m1 = joint.to_dense_batch(x1, b1)
        e2, m2 = joint.to_dense_batch(x2, b2)
        r1, r2, (w1, w2) = blk(e1, e2, m1, m2, True)
        i1, i2 = joint.DenseIndex(b1, sum(n1)), joint.DenseIndex(b2, sum(n2))
        p1, p2, (v1, v2) = blk.forward_packed(x1, x2, i1, i2, True)
        assert torch.allclose(i1.unpad(r1), p1, atol=1e-5) and torch.allclose(i2.unpad(r2), p2, atol=1e-5)
        assert torch.allclose(w1, v1, atol=1e-6) and torch.allclose(w2, v2, atol=1e-6)
        q1, q2, _ = blk.forward_packed(x1, x2, i1, i2, False)           # fused SDPA core
        assert torch.allclose(p1, q1, atol=1e-5) and torch.allclose(p2, q2, atol=1e-5)
        # gradients through pad / unpad
        x1g = x1.clone().requires_grad_()
        blk.forward_packed(x1g, x2, i1, i2, True)[0].square().sum().backward()
        e1g = x1.clone().requires_grad_()
        d, _ = joint.to_dense_batch(e1g, b1)
        (i1.unpad(blk(d, e2, m1, m2, True)[0])).square().sum().backward()
        assert torch.allclose(x1g.grad, e1g.grad, atol=1e-4)
