import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caster_dta_b200 import ops
torch.manual_seed(0)
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dy = torch.randn(m, n, device="cuda"); x = torch.randn(m, k, device="cuda")
dw, db = ops.linear_wgrad(dy, x)
ref = (dy.double().t() @ x.double()).float()
print("db err", float((db - dy.sum(0)).abs().max()))
print("nonzero frac", float((dw != 0).float().mean()), "max ref", float(ref.abs().max()), "max dw", float(dw.abs().max()))
err = (dw - ref).abs()
print("err by 32-row block x 8-col block:")
for rb in range(0, n, 32):
    print(rb, [round(float(err[rb:rb+32, cb:cb+8].max()), 2) for cb in range(0, k, 8)])
# single-element probes: which (row m, col n) of dy and (row m, col k) of x end up where
for (mm, nn, kk) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0), (0, 5, 3), (9, 40, 17), (3, 100, 31)]:
    dy2 = torch.zeros_like(dy); x2 = torch.zeros_like(x)
    dy2[mm, nn] = 1.0; x2[mm, kk] = 1.0
    d2, _ = ops.linear_wgrad(dy2, x2)
    nz = d2.nonzero().tolist()
    print("probe dy[%d,%d] x[%d,%d] -> nonzeros" % (mm, nn, mm, kk), nz[:6], [float(d2[a, b]) for a, b in nz[:6]])
