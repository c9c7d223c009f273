"""Pin the CPU oracle against golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

CPU only.  fp64 throughout, so the bar is round-off: 1e-12 relative.
"""
import numpy as np
import pytest
import torch

from helpers import assert_close, case, golden, json_blob, none_str
from oracle import featurizer_oracle, gvp_oracle, joint_oracle

TOL = 1e-12

GVP_CASES = ["gvp_relu_gate", "gvp_none_gate", "gvp_none_nogate", "gvp_relu_sigmoid_nogate",
             "gvp_relu_sigmoid_gate", "gvp_scalar_out", "gvp_scalar_in", "gvp_hdim", "gvp_ckpt_msg0"]


def _leaf(t):
    return t.clone().double().requires_grad_()


def _check_grads(c, outs, cots, inputs, params):
    names = list(params)
    loss = sum((o * ct).sum() for o, ct in zip(outs, cots))
    gr = torch.autograd.grad(loss, list(inputs.values()) + [params[k] for k in names], allow_unused=True)
    for (k, _), g in zip(inputs.items(), gr):
        assert_close(g, c["grad_" + k], TOL, "grad_" + k)
    for k, g in zip(names, gr[len(inputs):]):
        if k in c["grad_param"]:
            assert_close(g, c["grad_param"][k], TOL, "grad_param " + k)


@pytest.mark.parametrize("name", GVP_CASES)
def test_gvp_forward_backward(name):
    c = case(golden("gvp_units"), name)
    p = {k: _leaf(v) if v.numel() else v for k, v in c["param"].items()}
    vi, vo = int(c["in_dims"][1]), int(c["out_dims"][1])
    s, v = _leaf(c["s"]), _leaf(c["v"])
    x = (s, v) if vi else s
    out = gvp_oracle.gvp(p, "", x, none_str(c["scalar_act"]), none_str(c["vector_act"]),
                         bool(c["vector_gate"]), vo_if_scalar_in=vo if not vi else 0)
    outs = list(out) if isinstance(out, tuple) else [out]
    assert_close(outs[0], c["out_s"], TOL, "s")
    if vo:
        assert_close(outs[1], c["out_v"], TOL, "V")
    if vi:  # gradient cases
        cots = [c["cot_s"]] + ([c["cot_v"]] if vo else [])
        _check_grads(c, outs, cots, {"s": s, "v": v}, {k: t for k, t in p.items() if t.numel()})


@pytest.mark.parametrize("name", ["ln_sv", "ln_s"])
def test_layer_norm(name):
    c = case(golden("gvp_units"), name)
    p = {k: _leaf(v) for k, v in c["param"].items()}
    s, v = _leaf(c["s"]), _leaf(c["v"])
    if int(c["dims"][1]):
        os_, ov = gvp_oracle.layer_norm(p, "", (s, v))
        assert_close(os_, c["out_s"], TOL)
        assert_close(ov, c["out_v"], TOL)
        _check_grads(c, [os_, ov], [c["cot_s"], c["cot_v"]], {"s": s, "v": v}, p)
    else:
        os_ = gvp_oracle.layer_norm(p, "", s)
        assert_close(os_, c["out_s"], TOL)
        _check_grads(c, [os_], [c["cot_s"]], {"s": s}, p)


@pytest.mark.parametrize("name", ["conv_mean", "conv_sum", "conv_single"])
def test_gvp_conv(name):
    c = case(golden("gvp_units"), name)
    p = {k: _leaf(v) if v.numel() else v for k, v in c["param"].items()}
    s, v, es, ev = _leaf(c["s"]), _leaf(c["v"]), _leaf(c["es"]), _leaf(c["ev"])
    out = gvp_oracle.gvp_conv(p, "", (s, v), c["edge_index"], (es, ev), aggr=str(c["aggr"]),
                              n_layers=int(c["n_layers"]), scalar_act="relu", vector_act=None, vector_gate=True)
    assert_close(out[0], c["out_s"], TOL)
    assert_close(out[1], c["out_v"], TOL)
    _check_grads(c, list(out), [c["cot_s"], c["cot_v"]], {"s": s, "v": v, "es": es, "ev": ev},
                 {k: t for k, t in p.items() if t.numel()})


@pytest.mark.parametrize("name", ["layer_mean", "layer_sum", "layer_ff1", "layer_mask", "layer_autoreg", "layer_wide"])
def test_gvp_conv_layer(name):
    # layer_wide: BASELINE config-5 dims, nodes (100,16) / edges (32,1), its own fixture file
    c = case(golden("layer_wide" if name == "layer_wide" else "gvp_units"), name)
    p = {k: _leaf(v) if v.numel() else v for k, v in c["param"].items()}
    s, v, es, ev = _leaf(c["s"]), _leaf(c["v"]), _leaf(c["es"]), _leaf(c["ev"])
    aggr = none_str(c["aggr"]) or "mean"
    kw = {}
    if "node_mask" in c:
        kw["node_mask"] = c["node_mask"].bool()
    if "ar_s" in c:
        kw["autoregressive_x"] = (c["ar_s"].double(), c["ar_v"].double())
    out = gvp_oracle.gvp_conv_layer(p, "", (s, v), c["edge_index"], (es, ev), aggr=aggr,
                                    n_feedforward=int(c["n_feedforward"]), scalar_act="relu", vector_act=None,
                                    vector_gate=True, **kw)
    assert_close(out[0], c["out_s"], TOL)
    assert_close(out[1], c["out_v"], TOL)
    if "cot_s" in c:
        _check_grads(c, list(out), [c["cot_s"], c["cot_v"]], {"s": s, "v": v, "es": es, "ev": ev},
                     {k: t for k, t in p.items() if t.numel()})


@pytest.mark.parametrize("name", ["radius4", "knn10"])
def test_lba_encoder_checkpoint(name):
    g = golden("lba_checkpoint")
    kw = json_blob(g)
    c = case(g, name)
    p = {k[len("param/"):]: torch.from_numpy(g[k]).double().requires_grad_(g[k].size > 0)
         for k in g.files if k.startswith("param/")}
    xs, xv = _leaf(c["x_s"]), _leaf(c["x_v"])
    out = gvp_oracle.lba_encoder(p, "gnn_model.", (xs, xv), c["edge_index"], c["ntypes"], c["etypes"],
                                 (c["e_s"].double(), c["e_v"].double()), kw["num_ntypes"], kw["num_etypes"],
                                 kw["num_convs"], kw["aggr"])
    assert_close(out, c["out"], TOL)
    names = [k for k, t in p.items() if t.numel()]
    gr = torch.autograd.grad((out * c["cot"]).sum(), [xs, xv] + [p[k] for k in names])
    assert_close(gr[0], c["grad_x_s"], TOL)
    assert_close(gr[1], c["grad_x_v"], TOL)
    for k, gk in zip(names, gr[2:]):
        assert_close(gk, c["grad_param"][k], 1e-11, k, atol=1e-13)
    # the fp32 reference run stays within the stated fp32 tolerance of the fp64 one
    assert_close(c["out_fp32"], c["out"], 1e-4)


def test_joint_small():
    g = golden("joint_small")
    kw = json_blob(g)
    p = {k: v.double() if v.dtype.is_floating_point else v for k, v in case(g, "model")["param"].items()}
    pr, mo, out = case(g, "prot"), case(g, "mol"), case(g, "out")
    prot = dict(x=(pr["x_s"].double(), pr["x_v"].double()), edge_index=pr["edge_index"], ntypes=pr["ntypes"],
                etypes=pr["etypes"], eattr=(pr["e_s"].double(), pr["e_v"].double()), batch=pr["batch"])
    mol = dict(x=mo["x"].double(), edge_index=mo["edge_index"], ntypes=mo["ntypes"], etypes=mo["etypes"],
               eattr=mo["eattr"].double(), batch=mo["batch"])
    pred, weights = joint_oracle.joint_forward(p, kw, prot, mol)
    assert_close(pred, out["pred"], 1e-11)
    assert_close(weights[0][0], out["attn_p2m"], 1e-11)
    assert_close(weights[0][1], out["attn_m2p"], 1e-11)


@pytest.mark.parametrize("name", ["radius4", "knn30"])
def test_joint_full_checkpoint(name):
    """Oracle fed with the WHOLE shipped checkpoint reproduces the reference's affinities (config 1)."""
    g = golden("joint_checkpoint")
    kw = json_blob(g)
    p = {k: v.double() if v.dtype.is_floating_point else v for k, v in case(g, "model")["param"].items()}
    assert sum(v.numel() for v in p.values()) == 764396
    pr, mo, out = case(g, name + "/prot"), case(g, name + "/mol"), case(g, name + "/out")
    prot = dict(x=(pr["x_s"].double(), pr["x_v"].double()), edge_index=pr["edge_index"], ntypes=pr["ntypes"],
                etypes=pr["etypes"], eattr=(pr["e_s"].double(), pr["e_v"].double()), batch=pr["batch"])
    mol = dict(x=mo["x"].double(), edge_index=mo["edge_index"], ntypes=mo["ntypes"], etypes=mo["etypes"],
               eattr=mo["eattr"].double(), batch=mo["batch"])
    pred, weights = joint_oracle.joint_forward(p, kw, prot, mol)
    assert_close(pred, out["pred"], 1e-11)
    assert_close(weights[0][0], out["attn_p2m"], 1e-10)
    rs = case(g, "rescale")
    assert_close(pred * float(rs["std"]) + float(rs["mean"]), out["affinity"], 1e-11)
    assert_close(out["pred_fp32"], out["pred"], 1e-4)           # the reference's own fp32 run sits inside the tolerance


FEAT_SETTINGS = ["dist4_self", "dist8_noself", "num10_self", "num8_noself", "prop_self", "num_gt_n"]


@pytest.mark.parametrize("prot", ["p41", "p97"])
@pytest.mark.parametrize("setting", FEAT_SETTINGS)
def test_featurizer(prot, setting):
    g = golden("featurizer")
    coords = g[f"{prot}/coords"]
    c = case(g, f"{prot}/{setting}")
    thr = float(c["thresh"])
    ei, s, v = featurizer_oracle.residue_graph(coords, thr, str(c["thresh_type"]), bool(c["keep_self"]))
    assert np.array_equal(ei, c["edge_index"].numpy()), "edge_index must be bit-exact"
    assert np.array_equal(v, c["edge_v"].numpy()), "directions must be bit-exact"
    assert np.array_equal(s, c["edge_s"].numpy()), "RBF / pos-enc must be bit-exact (same libm)"


@pytest.mark.parametrize("prot", ["p41", "p97"])
def test_node_geometry(prot):
    g = golden("featurizer")
    s, v = featurizer_oracle.node_geometry_features(g[f"{prot}/coords"])
    ref_s, ref_v = g[f"{prot}/node_s"], g[f"{prot}/node_v"]
    assert np.allclose(s, ref_s[:, :6], atol=1e-6)
    assert np.allclose(v, ref_v, atol=1e-6)
