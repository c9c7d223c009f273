"""CPU emulation of the TF32 mode of the wide-dims backward (every GEMM operand rounded to a 10-bit mantissa) on the case of
tests/test_gpu_wide.py::test_wide_conv_tensor_core_mode_backward, against the fp64 oracle: per-tensor L2-relative errors and the
share of per-edge gradient rows beyond 3e-2.  Test infrastructure (imports oracle/ and the test helpers); prints a table."""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch.nn.functional as F
from caster_dta_b200 import modules, wide
from oracle import gvp_oracle
import test_wide_gemm_cpu as T
import test_gpu_wide as G

def tf32(x):
    if x.dtype != torch.float32: return x
    xi = x.contiguous().view(torch.int32)
    # round to nearest even on 13 dropped bits
    r = (xi + 0x0FFF + ((xi >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)

_mm, _addmm, _bmm = torch.Tensor.__matmul__, torch.addmm, torch.bmm
_addmm_ = torch.Tensor.addmm_
torch.Tensor.__matmul__ = lambda a, b: _mm(tf32(a), tf32(b))
torch.addmm = lambda c, a, b, **k: _addmm(c, tf32(a), tf32(b), **k)
torch.bmm = lambda a, b: _bmm(tf32(a), tf32(b))
torch.Tensor.addmm_ = lambda c, a, b: _addmm_(c, tf32(a), tf32(b))

wide._segsum = lambda rows, rowptr, index, n, mean=False: T.segsum_reference(rows, rowptr, index, n, mean)
wide.ENABLED = True
nd, ed = (100, 16), (32, 1)
p, ei, x, ea = G._case(900, 27000, nd, ed, seed=8, hub=False)
conv = modules.GVPConv(nd, nd, ed, aggr="mean", activations=(F.relu, None), vector_gate=True)
prog = conv._program(False)
plan = T.cpu_plan(ei, 900)
w = T.conv_weights(p, "conv.message_func.")
p64 = {k: v.double().requires_grad_(v.numel() > 0) for k, v in p.items()}
l64 = [t.double().requires_grad_() for t in (x[0], x[1], ea[0], ea[1])]
torch.Tensor.__matmul__, torch.addmm = _mm, _addmm
ref = gvp_oracle.gvp_conv(p64, "conv.", (l64[0], l64[1]), ei, (l64[2], l64[3]), aggr="mean", scalar_act="relu", vector_act=None, vector_gate=True)
(ref[0].sum() + ref[1].sum()).backward()
torch.Tensor.__matmul__ = lambda a, b: _mm(tf32(a), tf32(b))
torch.addmm = lambda c, a, b, **k: _addmm(c, tf32(a), tf32(b), **k)
cs, cv = torch.ones_like(ref[0]).float(), torch.ones_like(ref[1]).float()
out = wide.conv_forward(prog, plan, x[0], x[1], ea[0], ea[1], w)
print("fwd tf32 vs oracle", float((out[0].double()-ref[0]).abs().max()/ref[0].abs().max()), float((out[1].double()-ref[1]).abs().max()/ref[1].abs().max()))
got = wide.conv_backward(prog, plan, x[0], x[1], ea[0], ea[1], w, cs, cv)
l2 = lambda a, b: float((a.double()-b).norm()/b.norm())
for t, r, k in zip(got[:4], l64, ("grad_s","grad_v","grad_es","grad_ev")):
    d = (t.double()-r.grad).abs().reshape(t.shape[0], -1).amax(1)
    print(k, "L2", l2(t, r.grad), "rows>3e-2", float((d > 3e-2*r.grad.abs().max()).double().mean()))
names = ("wh.weight","ws.weight","ws.bias","wv.weight","wsv.weight","wsv.bias")
worst = 0
for l in range(3):
    for j, nm in enumerate(names):
        key = f"conv.message_func.{l}.{nm}"
        e = l2(got[4][6*l+j], p64[key].grad); worst = max(worst, e)
        print(key, e)
print("worst param", worst)
