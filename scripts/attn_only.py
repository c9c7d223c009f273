"""Run the fused cross-attention kernels alone at Davis-shape sizes (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caster_dta_b200 import joint, ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
nres = torch.randint(300, 1000, (32,), generator=g)
natm = torch.randint(20, 46, (32,), generator=g)
br = torch.repeat_interleave(torch.arange(32), nres).to(dev)
ba = torch.repeat_interleave(torch.arange(32), natm).to(dev)
dr, da = joint.DenseIndex(br, int(nres.sum())), joint.DenseIndex(ba, int(natm.sum()))
R = torch.randn(int(nres.sum()), 128, device=dev, requires_grad=True)
A = torch.randn(int(natm.sum()), 128, device=dev, requires_grad=True)
for _ in range(2):
    o1, w1, _ = ops.CrossAttnFunction.apply(R, A, A, dr.ptr, da.ptr, dr.batch, da.batch, 8, dr.m, da.m, None, True)
    o2, w2, _ = ops.CrossAttnFunction.apply(A, R, R, da.ptr, dr.ptr, da.batch, dr.batch, 8, da.m, dr.m, None, True)
    (o1.sum() + o2.sum()).backward()
torch.cuda.synchronize()
print("ok")
