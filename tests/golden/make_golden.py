"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs `/root/reference`, read-only):

    python tests/golden/make_golden.py

The reference modules (`models/gvp_layers.py`, `models/protein_gnn.py`, `models/joint_gnn.py`,
`utils/create_protein_features.py`, `utils/create_graphs.py`) are imported unchanged through the dependency
shim `ref_shim.py`, evaluated in fp64 on seeded inputs whose values are exactly representable in fp32, and
the inputs / parameters / outputs / gradients are written as `*.npz`.  These files are what pins the oracle
(`oracle/`) and, through it, the CUDA path: the reference ships no tests or golden vectors of its own.

Fixtures:
  gvp_units.npz      GVP (8 activation/gate/shape variants), LayerNorm, GVPConv, GVPConvLayer (+node_mask,
                     +autoregressive) -- outputs and autograd gradients in fp64.
  layer_wide.npz     ONE GVPConvLayer at the BASELINE config-5 dims, nodes (100,16) / edges (32,1), vector_gate, (ReLU, None),
                     mean aggregation (171 909 parameters, stored as fp32): outputs and every gradient in fp64.
  featurizer.npz     compute_residue_edge_features + construct_graph on synthetic backbones, 6 settings.
  lba_checkpoint.npz protein-GNN slice of the shipped checkpoint (15 117 parameters) + a small synthetic batch
                     (radius 4 A and kNN-10 graphs) + the reference embeddings [N,64].
  joint_small.npz    a reduced-width JointGNN (random init, seed 9): state_dict, batch, predicted affinity.
  joint_checkpoint.npz  the WHOLE shipped checkpoint (764 396 parameters) + two small batches (radius 4 A, kNN-30) + the
                     reference's predicted affinities, residue embeddings and attention maps (fp64, and fp32 predictions).
  sampler.json       mini-batches produced by the reference PMD_BatchSampler (dataset/dual_dataset.py:424-522) on a
                     seeded list of pair sizes, several settings, shuffle off.
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402

ref_shim.install()

import torch.nn.functional as F  # noqa: E402
from models import gvp_layers as ref  # noqa: E402
from models.joint_gnn import JointGNN  # noqa: E402
import utils.create_protein_features as ref_feat  # noqa: E402
import utils.create_graphs as ref_graphs  # noqa: E402

from caster_dta_b200 import synth  # noqa: E402

ACTS = {None: None, "relu": F.relu, "sigmoid": torch.sigmoid}


def f32(t):
    """Round to fp32-representable values, keep as fp64."""
    return t.float().double()


def fix_module(m):
    m.double()
    with torch.no_grad():
        for prm in m.parameters():
            prm.copy_(f32(prm))
    return m


def rand_graph(n, e, gen, isolated=3):
    src = torch.randint(0, n, (e,), generator=gen)
    dst = torch.randint(0, n - isolated, (e,), generator=gen)      # last `isolated` nodes get no in-edges
    key = src * n + dst
    order = torch.argsort(key, stable=True)
    return torch.stack([src[order], dst[order]])


def put(store, case, **arrays):
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        store[f"{case}/{k}"] = np.asarray(v)


def put_state(store, case, module):
    for k, v in module.state_dict().items():
        store[f"{case}/param/{k}"] = v.detach().numpy()


def grads_of(outs, cots, wrt):
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    return torch.autograd.grad(loss, wrt, allow_unused=True)


def make_gvp_units():
    store = {}
    g = torch.Generator().manual_seed(9)
    n = 50
    variants = {
        "gvp_relu_gate": dict(i=(8, 3), o=(6, 4), acts=("relu", None), gate=True, h=None),
        "gvp_none_gate": dict(i=(8, 3), o=(6, 4), acts=(None, None), gate=True, h=None),
        "gvp_none_nogate": dict(i=(8, 3), o=(6, 4), acts=(None, None), gate=False, h=None),
        "gvp_relu_sigmoid_nogate": dict(i=(8, 3), o=(6, 4), acts=("relu", "sigmoid"), gate=False, h=None),
        "gvp_relu_sigmoid_gate": dict(i=(8, 3), o=(6, 4), acts=("relu", "sigmoid"), gate=True, h=None),
        "gvp_scalar_out": dict(i=(8, 3), o=(6, 0), acts=("relu", None), gate=True, h=None),
        "gvp_scalar_in": dict(i=(8, 0), o=(6, 2), acts=("relu", None), gate=True, h=None),
        "gvp_hdim": dict(i=(5, 2), o=(7, 3), acts=("relu", None), gate=True, h=5),
        "gvp_ckpt_msg0": dict(i=(64, 9), o=(16, 4), acts=("relu", None), gate=True, h=None),
    }
    for name, c in variants.items():
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        m = fix_module(ref.GVP(c["i"], c["o"], h_dim=c["h"], activations=(ACTS[c["acts"][0]], ACTS[c["acts"][1]]),
                               vector_gate=c["gate"]))
        s = f32(torch.randn(n, c["i"][0], generator=g)).requires_grad_()
        v = f32(torch.randn(n, c["i"][1], 3, generator=g)).requires_grad_()
        v.data[3] = 0.0                                                # exercises the 1e-8 clamp
        x = (s, v) if c["i"][1] else s
        out = m(x)
        outs = list(out) if isinstance(out, tuple) else [out]
        cots = [f32(torch.randn(o.shape, generator=g)) for o in outs]
        params = [p for p in m.parameters() if p.numel()]
        wrt = [s] + ([v] if c["i"][1] else []) + params
        gr = grads_of(outs, cots, wrt)
        put(store, name, in_dims=c["i"], out_dims=c["o"], h_dim=-1 if c["h"] is None else c["h"],
            scalar_act=str(c["acts"][0]), vector_act=str(c["acts"][1]), vector_gate=int(c["gate"]),
            s=s, v=v, out_s=outs[0], cot_s=cots[0])
        if len(outs) > 1:
            put(store, name, out_v=outs[1], cot_v=cots[1])
        put_state(store, name, m)
        put(store, name, grad_s=gr[0])
        k = 1
        if c["i"][1]:
            put(store, name, grad_v=gr[1])
            k = 2
        for (pn, _), gp in zip([(a, b) for a, b in m.named_parameters() if b.numel()], gr[k:]):
            store[f"{name}/grad_param/{pn}"] = (torch.zeros(1) if gp is None else gp).numpy()

    # LayerNorm
    for name, dims in {"ln_sv": (6, 4), "ln_s": (6, 0)}.items():
        torch.manual_seed(3)
        m = ref.LayerNorm(dims)
        with torch.no_grad():
            m.scalar_norm.weight.copy_(1 + 0.2 * torch.randn(dims[0]))
            m.scalar_norm.bias.copy_(0.2 * torch.randn(dims[0]))
        fix_module(m)
        s = f32(torch.randn(n, dims[0], generator=g)).requires_grad_()
        v = f32(torch.randn(n, max(dims[1], 1), 3, generator=g)).requires_grad_()
        out = m((s, v)) if dims[1] else m(s)
        outs = list(out) if isinstance(out, tuple) else [out]
        cots = [f32(torch.randn(o.shape, generator=g)) for o in outs]
        wrt = [s] + ([v] if dims[1] else []) + list(m.parameters())
        gr = grads_of(outs, cots, wrt)
        put(store, name, dims=dims, s=s, v=v, out_s=outs[0], cot_s=cots[0], grad_s=gr[0])
        if dims[1]:
            put(store, name, out_v=outs[1], cot_v=cots[1], grad_v=gr[1])
        put_state(store, name, m)
        for (pn, _), gp in zip(m.named_parameters(), gr[-2:]):
            store[f"{name}/grad_param/{pn}"] = gp.numpy()

    # GVPConv / GVPConvLayer on a small random multigraph with isolated nodes
    nn_, ee = 40, 260
    ei = rand_graph(nn_, ee, g)
    nd, ed = (16, 4), (32, 1)
    for name, c in {
        "conv_mean": dict(aggr="mean", n_layers=3),
        "conv_sum": dict(aggr="sum", n_layers=3),
        "conv_single": dict(aggr="add", n_layers=1),
    }.items():
        torch.manual_seed(11)
        m = fix_module(ref.GVPConv(nd, nd, ed, n_layers=c["n_layers"], aggr=c["aggr"],
                                   activations=(F.relu, None), vector_gate=True))
        s = f32(torch.randn(nn_, nd[0], generator=g)).requires_grad_()
        v = f32(torch.randn(nn_, nd[1], 3, generator=g)).requires_grad_()
        es = f32(torch.randn(ee, ed[0], generator=g)).requires_grad_()
        ev = f32(torch.randn(ee, ed[1], 3, generator=g)).requires_grad_()
        outs = list(m((s, v), ei, (es, ev)))
        cots = [f32(torch.randn(o.shape, generator=g)) for o in outs]
        named = [(a, b) for a, b in m.named_parameters() if b.numel()]
        gr = grads_of(outs, cots, [s, v, es, ev] + [b for _, b in named])
        put(store, name, node_dims=nd, edge_dims=ed, aggr=c["aggr"], n_layers=c["n_layers"], edge_index=ei,
            s=s, v=v, es=es, ev=ev, out_s=outs[0], out_v=outs[1], cot_s=cots[0], cot_v=cots[1],
            grad_s=gr[0], grad_v=gr[1], grad_es=gr[2], grad_ev=gr[3])
        put_state(store, name, m)
        for (pn, _), gp in zip(named, gr[4:]):
            store[f"{name}/grad_param/{pn}"] = gp.numpy()

    for name, c in {
        "layer_mean": dict(aggr=None, n_ff=2, mask=False, ar=False),
        "layer_sum": dict(aggr="sum", n_ff=2, mask=False, ar=False),
        "layer_ff1": dict(aggr="sum", n_ff=1, mask=False, ar=False),
        "layer_mask": dict(aggr="sum", n_ff=2, mask=True, ar=False),
        "layer_autoreg": dict(aggr=None, n_ff=2, mask=False, ar=True),
    }.items():
        torch.manual_seed(13)
        m = fix_module(ref.GVPConvLayer(nd, ed, n_feedforward=c["n_ff"], drop_rate=0.1, autoregressive=c["ar"],
                                        activations=(F.relu, None), vector_gate=True, aggr=c["aggr"])).eval()
        with torch.no_grad():
            for k in range(2):
                m.norm[k].scalar_norm.weight.copy_(f32(1 + 0.2 * torch.randn(nd[0])))
                m.norm[k].scalar_norm.bias.copy_(f32(0.2 * torch.randn(nd[0])))
        s = f32(torch.randn(nn_, nd[0], generator=g)).requires_grad_()
        v = f32(torch.randn(nn_, nd[1], 3, generator=g)).requires_grad_()
        es = f32(torch.randn(ee, ed[0], generator=g)).requires_grad_()
        ev = f32(torch.randn(ee, ed[1], 3, generator=g)).requires_grad_()
        kw = {}
        if c["mask"]:
            mask = torch.rand(nn_, generator=g) < 0.6
            kw["node_mask"] = mask
            put(store, name, node_mask=mask)
        if c["ar"]:
            ars = f32(torch.randn(nn_, nd[0], generator=g))
            arv = f32(torch.randn(nn_, nd[1], 3, generator=g))
            kw["autoregressive_x"] = (ars, arv)
            put(store, name, ar_s=ars, ar_v=arv)
        # node_mask writes in place into the inputs (gvp_layers.py:413): feed clones, keep grads off for it
        if c["mask"]:
            outs = list(m((s.detach().clone(), v.detach().clone()), ei, (es.detach(), ev.detach()), **kw))
            put(store, name, out_s=outs[0], out_v=outs[1])
        else:
            outs = list(m((s, v), ei, (es, ev), **kw))
            cots = [f32(torch.randn(o.shape, generator=g)) for o in outs]
            named = [(a, b) for a, b in m.named_parameters() if b.numel()]
            gr = grads_of(outs, cots, [s, v, es, ev] + [b for _, b in named])
            put(store, name, out_s=outs[0], out_v=outs[1], cot_s=cots[0], cot_v=cots[1],
                grad_s=gr[0], grad_v=gr[1], grad_es=gr[2], grad_ev=gr[3])
            for (pn, _), gp in zip(named, gr[4:]):
                store[f"{name}/grad_param/{pn}"] = gp.numpy()
        put(store, name, node_dims=nd, edge_dims=ed, aggr=str(c["aggr"]), n_feedforward=c["n_ff"],
            autoregressive=int(c["ar"]), edge_index=ei, s=s, v=v, es=es, ev=ev)
        put_state(store, name, m)
    np.savez_compressed(os.path.join(HERE, "gvp_units.npz"), **store)
    print("gvp_units.npz", len(store), "arrays")


def make_layer_wide():
    """BASELINE config 5: `GVPConvLayer(vector_gate=True, activations=(ReLU, None), aggr='mean', drop_rate=0)`
    at nodes (100,16), edges (32,1) on a small random multigraph (isolated nodes, a zero vector row)."""
    store = {}
    g = torch.Generator().manual_seed(95)
    nn_, ee = 48, 360
    nd, ed = (100, 16), (32, 1)
    ei = rand_graph(nn_, ee, g)
    torch.manual_seed(17)
    m = fix_module(ref.GVPConvLayer(nd, ed, drop_rate=0.0, activations=(F.relu, None), vector_gate=True, aggr="mean")).eval()
    with torch.no_grad():
        for k in range(2):
            m.norm[k].scalar_norm.weight.copy_(f32(1 + 0.2 * torch.randn(nd[0])))
            m.norm[k].scalar_norm.bias.copy_(f32(0.2 * torch.randn(nd[0])))
    s = f32(torch.randn(nn_, nd[0], generator=g)).requires_grad_()
    v = f32(torch.randn(nn_, nd[1], 3, generator=g)).requires_grad_()
    v.data[3] = 0.0
    es = f32(torch.randn(ee, ed[0], generator=g)).requires_grad_()
    ev = f32(torch.randn(ee, ed[1], 3, generator=g)).requires_grad_()
    outs = list(m((s, v), ei, (es, ev)))
    cots = [f32(torch.randn(o.shape, generator=g)) for o in outs]
    named = [(a, b) for a, b in m.named_parameters() if b.numel()]
    gr = grads_of(outs, cots, [s, v, es, ev] + [b for _, b in named])
    name = "layer_wide"
    put(store, name, node_dims=nd, edge_dims=ed, aggr="mean", n_feedforward=2, autoregressive=0, edge_index=ei,
        s=s, v=v, es=es, ev=ev, out_s=outs[0], out_v=outs[1], cot_s=cots[0], cot_v=cots[1],
        grad_s=gr[0], grad_v=gr[1], grad_es=gr[2], grad_ev=gr[3])
    for k, t in m.state_dict().items():                                 # fp32-representable by construction: store as fp32
        store[f"{name}/param/{k}"] = t.detach().float().numpy()
    for (pn, _), gp in zip(named, gr[4:]):
        store[f"{name}/grad_param/{pn}"] = gp.numpy()
    np.savez_compressed(os.path.join(HERE, "layer_wide.npz"), **store)
    print("layer_wide.npz", len(store), "arrays,", sum(p.numel() for p in m.parameters()), "parameters")


def reference_graph(coords, idents, thresh, ttype, keep_self):
    node = ref_feat.compute_residue_node_features(coords, idents, True, False, False, True)
    edge = ref_feat.compute_residue_edge_features(coords, idents, thresh, ttype, keep_self, True)
    n = coords.shape[0]
    data = ref_graphs.construct_graph(node, edge, idents, np.zeros((n, n), dtype=np.int64))
    return data


def make_featurizer():
    store = {}
    rng = np.random.default_rng(9)
    settings = {
        "dist4_self": (4.0, "dist", True),
        "dist8_noself": (8.0, "dist", False),
        "num10_self": (10, "num", True),
        "num8_noself": (8, "num", False),
        "prop_self": (0.15, "prop", True),
        "num_gt_n": (64, "num", False),
    }
    for pname, n, sa in (("p41", 41, False), ("p97", 97, True)):
        coords = synth.random_backbone(n, rng, self_avoiding=sa)
        idents = rng.integers(0, 20, size=n)
        store[f"{pname}/coords"] = coords
        store[f"{pname}/idents"] = idents
        for sname, (thr, tt, ks) in settings.items():
            d = reference_graph(coords, idents, thr, tt, ks)
            case = f"{pname}/{sname}"
            put(store, case, thresh=thr, thresh_type=tt, keep_self=int(ks), edge_index=d.edge_index,
                edge_s=d.edge_attr[0], edge_v=d.edge_attr[1])
        put(store, pname, node_s=d.x[0], node_v=d.x[1])
    np.savez_compressed(os.path.join(HERE, "featurizer.npz"), **store)
    print("featurizer.npz", len(store), "arrays")


def build_protein_inputs(seed, pairs, thresh, ttype):
    """Small batch through the REFERENCE featurizer (node + edge), collated like PyG Batch."""
    rng = np.random.default_rng(seed)
    xs, xv, nts, eis, ess, evs, batch = [], [], [], [], [], [], []
    off = 0
    for k in range(pairs):
        n = int(rng.integers(25, 60))
        coords = synth.random_backbone(n, rng, self_avoiding=True)
        idents = rng.integers(0, 20, size=n)
        d = reference_graph(coords, idents, thresh, ttype, True)
        xs.append(d.x[0]); xv.append(d.x[1]); nts.append(d.node_type)
        eis.append(d.edge_index + off); ess.append(d.edge_attr[0]); evs.append(d.edge_attr[1])
        batch.append(torch.full((n,), k, dtype=torch.long))
        off += n
    ei = torch.cat(eis, 1)
    return dict(x=(torch.cat(xs), torch.cat(xv)), edge_index=ei, ntypes=torch.cat(nts),
                etypes=torch.zeros(ei.shape[1], dtype=torch.long), eattr=(torch.cat(ess), torch.cat(evs)),
                batch=torch.cat(batch))


def make_lba_checkpoint():
    root = os.path.join(ref_shim.REFERENCE_ROOT, "pretrained_model_downstream")
    kw = json.load(open(os.path.join(root, "model_kwargs.json")))
    ck = [f for f in sorted(os.listdir(root)) if f.startswith("bestvalmodel")][0]
    sd = torch.load(os.path.join(root, ck), weights_only=True, map_location="cpu")
    sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}
    model = JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])
    model.load_state_dict(sd)
    enc = model.protein_gnn.double().eval()
    store = {"kwargs_json": np.frombuffer(json.dumps(kw["protein_gnn_kwargs"]).encode(), dtype=np.uint8)}
    for k, v in sd.items():
        if k.startswith("protein_gnn."):
            store["param/" + k[len("protein_gnn."):]] = v.numpy()
    for case, (thr, tt) in {"radius4": (4.0, "dist"), "knn10": (10, "num")}.items():
        inp = build_protein_inputs(21, 3, thr, tt)
        xs = inp["x"][0].double().requires_grad_()
        xv = inp["x"][1].double().requires_grad_()
        es, ev = inp["eattr"][0].double(), inp["eattr"][1].double()
        out = enc((xs, xv), inp["edge_index"], inp["ntypes"], inp["etypes"], eattr=(es, ev), batch=inp["batch"])
        cot = f32(torch.randn(out.shape, generator=torch.Generator().manual_seed(5)))
        named = [(a, b) for a, b in enc.named_parameters() if b.numel()]
        gr = torch.autograd.grad((out * cot).sum(), [xs, xv] + [b for _, b in named])
        put(store, case, x_s=inp["x"][0], x_v=inp["x"][1], edge_index=inp["edge_index"], ntypes=inp["ntypes"],
            etypes=inp["etypes"], e_s=inp["eattr"][0], e_v=inp["eattr"][1], batch=inp["batch"], out=out,
            cot=cot, grad_x_s=gr[0], grad_x_v=gr[1])
        for (pn, _), gp in zip(named, gr[2:]):
            store[f"{case}/grad_param/{pn}"] = gp.numpy()
        with torch.no_grad():
            out32 = model.protein_gnn.float()(inp["x"], inp["edge_index"], inp["ntypes"], inp["etypes"],
                                               eattr=inp["eattr"], batch=inp["batch"])
        model.protein_gnn.double()
        put(store, case, out_fp32=out32)
    np.savez_compressed(os.path.join(HERE, "lba_checkpoint.npz"), **store)
    print("lba_checkpoint.npz", len(store), "arrays")


def make_joint_small():
    kw = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "pretrained_model_downstream", "model_kwargs.json")))
    kw["protein_gnn_kwargs"]["out_channels"] = 16
    kw["molecule_gnn_kwargs"]["out_channels"] = 16
    kw["joint_gnn_kwargs"]["pairwise_embedding_dim"] = 32
    kw["joint_gnn_kwargs"]["n_attention_heads"] = 4
    torch.manual_seed(9)
    model = fix_module(JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])).eval()
    prot = build_protein_inputs(33, 4, 4.0, "dist")
    mol = synth.molecule_batch(4, seed=33, lo=8, hi=20)
    mol = {k: torch.from_numpy(v) for k, v in mol.items()}
    pd = dict(x=(prot["x"][0].double(), prot["x"][1].double()), edge_index=prot["edge_index"], ntypes=prot["ntypes"],
              etypes=prot["etypes"], eattr=(prot["eattr"][0].double(), prot["eattr"][1].double()), batch=prot["batch"])
    md = dict(x=mol["x"].double(), edge_index=mol["edge_index"], ntypes=mol["ntypes"], etypes=mol["etypes"],
              eattr=mol["eattr"].double(), batch=mol["batch"])
    with torch.no_grad():
        pred, attn = model(pd, md)
        emb = model.protein_gnn(**pd)
    store = {"kwargs_json": np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)}
    put_state(store, "model", model)
    put(store, "prot", x_s=prot["x"][0], x_v=prot["x"][1], edge_index=prot["edge_index"], ntypes=prot["ntypes"],
        etypes=prot["etypes"], e_s=prot["eattr"][0], e_v=prot["eattr"][1], batch=prot["batch"])
    put(store, "mol", **mol)
    put(store, "out", pred=pred, residue_embed=emb, attn_p2m=attn[0][0], attn_m2p=attn[0][1])
    np.savez_compressed(os.path.join(HERE, "joint_small.npz"), **store)
    print("joint_small.npz", len(store), "arrays")


def make_joint_checkpoint():
    """The WHOLE shipped checkpoint (764 396 parameters, `pretrained_model_downstream/bestvalmodel_*.pt`) loaded the way
    `inference/inference_utils.py:40-68` does, evaluated the way `inference/evaluation.py:43-46` does (eval mode,
    `forward_with_graphs`-equivalent call, predictions un-standardised with `dataset_rescale_params.json`) on two small
    synthetic batches: radius 4 A + self loops (the checkpoint's own graph type, `dataset_kwargs.json`) and kNN-30
    (BASELINE config 1).  The weights travel inside the fixture (fp32), so the GPU box needs no reference checkout."""
    root = os.path.join(ref_shim.REFERENCE_ROOT, "pretrained_model_downstream")
    kw = json.load(open(os.path.join(root, "model_kwargs.json")))
    rescale = json.load(open(os.path.join(root, "dataset_rescale_params.json")))["standardize"]
    ck = [f for f in sorted(os.listdir(root)) if f.startswith("bestvalmodel")][0]
    sd = torch.load(os.path.join(root, ck), weights_only=True, map_location="cpu")
    sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}
    model = JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])
    model.load_state_dict(sd)
    model.eval()
    store = {"kwargs_json": np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)}
    for k, v in sd.items():
        store["model/param/" + k] = v.numpy()
    put(store, "rescale", mean=rescale["scale_mean_factor"], std=rescale["scale_std_factor"])
    for case, (thr, tt, pairs, lo, hi, seed) in {"radius4": (4.0, "dist", 4, 90, 260, 41), "knn30": (30, "num", 3, 60, 110, 43)}.items():
        rng = np.random.default_rng(seed)
        xs, xv, nts, eis, ess, evs, batch = [], [], [], [], [], [], []
        off = 0
        for k in range(pairs):
            n = int(rng.integers(lo, hi))
            coords = synth.random_backbone(n, rng, self_avoiding=(tt == "dist"))
            idents = rng.integers(0, 20, size=n)
            d = reference_graph(coords, idents, thr, tt, True)
            xs.append(d.x[0]); xv.append(d.x[1]); nts.append(d.node_type)
            eis.append(d.edge_index + off); ess.append(d.edge_attr[0]); evs.append(d.edge_attr[1])
            batch.append(torch.full((n,), k, dtype=torch.long))
            off += n
        ei = torch.cat(eis, 1)
        prot = dict(x=(torch.cat(xs), torch.cat(xv)), edge_index=ei, ntypes=torch.cat(nts),
                    etypes=torch.zeros(ei.shape[1], dtype=torch.long), eattr=(torch.cat(ess), torch.cat(evs)), batch=torch.cat(batch))
        mol = {k: torch.from_numpy(v) for k, v in synth.molecule_batch(pairs, seed=seed).items()}
        outs = {}
        for name, dt in (("fp64", torch.float64), ("fp32", torch.float32)):
            model.to(dt)
            pd = dict(x=(prot["x"][0].to(dt), prot["x"][1].to(dt)), edge_index=prot["edge_index"], ntypes=prot["ntypes"],
                      etypes=prot["etypes"], eattr=(prot["eattr"][0].to(dt), prot["eattr"][1].to(dt)), batch=prot["batch"])
            md = dict(x=mol["x"].to(dt), edge_index=mol["edge_index"], ntypes=mol["ntypes"], etypes=mol["etypes"],
                      eattr=mol["eattr"].to(dt), batch=mol["batch"])
            with torch.no_grad():
                pred, attn = model(pd, md)
                emb = model.protein_gnn(**pd)
            outs[name] = (pred, attn, emb)
        model.float()
        pred, attn, emb = outs["fp64"]
        put(store, case + "/prot", x_s=prot["x"][0], x_v=prot["x"][1], edge_index=prot["edge_index"], ntypes=prot["ntypes"],
            etypes=prot["etypes"], e_s=prot["eattr"][0], e_v=prot["eattr"][1], batch=prot["batch"])
        put(store, case + "/mol", **mol)
        put(store, case + "/out", pred=pred, affinity=pred * rescale["scale_std_factor"] + rescale["scale_mean_factor"],
            residue_embed=emb, attn_p2m=attn[0][0], attn_m2p=attn[0][1], pred_fp32=outs["fp32"][0])
        print(case, "pred", pred.flatten().tolist(), "fp32-fp64", float((outs["fp32"][0].double() - pred).abs().max()))
    np.savez_compressed(os.path.join(HERE, "joint_checkpoint.npz"), **store)
    print("joint_checkpoint.npz", len(store), "arrays")


def make_sampler():
    """`dataset/dual_dataset.py` imports mdtraj / rdkit at module level, so the sampler class alone is compiled from
    its own source text (located with `ast`, executed unchanged) and driven with a stand-in dataset."""
    import ast
    path = os.path.join(ref_shim.REFERENCE_ROOT, "dataset", "dual_dataset.py")
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "PMD_BatchSampler")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    sampler_cls = ns["PMD_BatchSampler"]

    class G:
        def __init__(self, n, e):
            self.num_nodes, self.num_edges = n, e

    rng = np.random.default_rng(9)
    pairs = 200
    pn = rng.integers(25, 2000, size=pairs)
    pe = pn * 30 - rng.integers(0, 200, size=pairs)
    mn = rng.integers(8, 60, size=pairs)
    me = mn * 3 + rng.integers(0, 6, size=pairs)
    data = [(G(int(a), int(b)), G(int(c), int(d)), None) for a, b, c, d in zip(pn, pe, mn, me)]
    settings = [
        dict(max_num=400000, count_elem="edge", graph_type="both", include_nodepair=True, max_bsize=None),
        dict(max_num=400000, count_elem="edge", graph_type="both", include_nodepair=True, max_bsize=8),
        dict(max_num=150000, count_elem="edge", graph_type="protein", include_nodepair=False, max_bsize=None),
        dict(max_num=6000, count_elem="node", graph_type="both", include_nodepair=False, max_bsize=32),
        dict(max_num=900, count_elem="edge", graph_type="molecule", include_nodepair=False, max_bsize=None),
        dict(max_num=30000, count_elem="node", graph_type="both", include_nodepair=True, max_bsize=None),
    ]
    cases = []
    for kw in settings:
        batches = [list(b) for b in sampler_cls(data, shuffle=False, skip_too_big=False, **kw)]
        cases.append({"kwargs": kw, "batches": batches})
    out = {"protein_nodes": pn.tolist(), "protein_edges": pe.tolist(), "molecule_nodes": mn.tolist(),
           "molecule_edges": me.tolist(), "cases": cases}
    json.dump(out, open(os.path.join(HERE, "sampler.json"), "w"))
    print("sampler.json", [len(c["batches"]) for c in cases], "batches")


if __name__ == "__main__":
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "sampler":
        make_sampler()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "joint_checkpoint":
        make_joint_checkpoint()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "layer_wide":
        make_layer_wide()
        sys.exit(0)
    make_gvp_units()
    make_layer_wide()
    make_featurizer()
    make_lba_checkpoint()
    make_joint_small()
    make_joint_checkpoint()
    make_sampler()
