import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caster_dta_b200 import ops, synth
dev = torch.device("cuda")
ei_np, n = synth.conv_microbench_graph(10_000_000, 30)
ei = torch.from_numpy(ei_np).to(dev); e = ei.shape[1]
x = (torch.randn(n, 16, device=dev), torch.randn(n, 4, 3, device=dev))
ea = (torch.randn(e, 32, device=dev), torch.randn(e, 1, 3, device=dev))
plan = ops.GraphPlan(ei, n)
rows = torch.randn(e, 28, device=dev)
out = torch.empty(n, 28, device=dev)
for _ in range(3):
    ops.gather_message_input(ei, x, ea)
    ops.segment_reduce(rows, plan, "sum", use_perm=False, out=out)
    ops.segment_reduce(rows, plan, "mean", use_perm=True, out=out)
torch.cuda.synchronize()
print("ok")
