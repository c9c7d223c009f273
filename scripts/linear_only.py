"""Dense-layer kernels alone at Davis-shape sizes (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caster_dta_b200 import ops
m = 22806
for n, k in ((128, 128), (256, 128), (128, 256)):
    x = torch.randn(m, k, device="cuda", requires_grad=True)
    w = torch.randn(n, k, device="cuda", requires_grad=True)
    b = torch.randn(n, device="cuda", requires_grad=True)
    for _ in range(3):
        y = ops.linear(x, w, b)
        y.sum().backward()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    xd = x.detach(); dy = torch.randn(m, n, device="cuda")
    ev[0].record()
    for _ in range(20): ops._linear_fwd(xd, w.detach(), b.detach())
    ev[1].record()
    for _ in range(20): ops._linear_dgrad(dy, w.detach())
    ev[2].record()
    for _ in range(20): ops.linear_wgrad(dy, xd)
    ev[3].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) / 20 * 1e3 for i in range(3)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.nn.functional.linear(xd, w.detach(), b.detach())
    e1.record(); torch.cuda.synchronize()
    print(f"N={n} K={k}: fwd {t[0]:.1f} us  dgrad {t[1]:.1f} us  wgrad {t[2]:.1f} us   cublas fwd {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
