"""Per-tensor errors of the fused GVPConv (both kernel families) against the fp64 oracle: outputs, input / edge-attribute
gradients and every parameter gradient, on hub / ragged random graphs.  `python scripts/conv_vs_oracle.py` on a B200."""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch.nn.functional as F
import caster_dta_b200 as cg
from caster_dta_b200 import _lib
from oracle import gvp_oracle
from test_gpu_parity import _random_layer_case
DEV='cuda'
for (n,e,hub,aggr) in [(1237,30011,True,"mean"),(1237,30011,True,"sum"),(3000,45000,False,"sum")]:
    nd, ed = (16,4),(32,1)
    p, ei, x, ea = _random_layer_case(n,e,nd,ed,seed=4,hub=hub,aggr=aggr)
    conv = cg.GVPConv(nd, nd, ed, aggr=aggr, activations=(F.relu, None), vector_gate=True)
    conv.load_state_dict({k[len("conv."):]: v for k, v in p.items() if k.startswith("conv.")}, strict=True)
    conv.to(DEV)
    g = torch.Generator().manual_seed(1)
    cs, cv = torch.randn(n,16,generator=g), torch.randn(n,4,3,generator=g)
    res = {}
    for mode in (True, False):
        _lib.set_fast_paths(mode)
        t = [a.clone().to(DEV).requires_grad_() for a in (x[0],x[1],ea[0],ea[1])]
        conv.zero_grad(set_to_none=True)
        out = conv((t[0],t[1]), ei.to(DEV), (t[2],t[3]))
        ((out[0]*cs.to(DEV)).sum()+(out[1]*cv.to(DEV)).sum()).backward()
        res[mode] = [out[0].detach().cpu().double(), out[1].detach().cpu().double()] + [a.grad.cpu().double() for a in t] + [q.grad.cpu().double() for q in conv.parameters() if q.numel()]
    _lib.set_fast_paths(True)
    p64 = {k: v.double().requires_grad_(v.numel()>0) for k,v in p.items()}
    l64 = [a.double().requires_grad_() for a in (x[0],x[1],ea[0],ea[1])]
    ref = gvp_oracle.gvp_conv(p64, "conv.", (l64[0],l64[1]), ei, (l64[2],l64[3]), aggr=aggr, scalar_act="relu", vector_act=None, vector_gate=True)
    ((ref[0]*cs.double()).sum()+(ref[1]*cv.double()).sum()).backward()
    names = [k for k,v in conv.named_parameters() if v.numel()]
    refs = [ref[0].detach(), ref[1].detach()] + [a.grad for a in l64] + [p64["conv."+k].grad for k in names]
    labels = ["out_s","out_v","d_x_s","d_x_v","d_e_s","d_e_v"] + names
    print(n,e,hub,aggr)
    for lab, f, gnr, r in zip(labels, res[True], res[False], refs):
        sc = float(r.abs().max())
        print(f"  {lab:28s} fast {float((f-r).abs().max())/sc:.2e} generic {float((gnr-r).abs().max())/sc:.2e}")
