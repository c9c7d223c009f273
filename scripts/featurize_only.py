import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import caster_dta_b200 as cg
from caster_dta_b200 import synth
dev = torch.device("cuda")
pb = synth.protein_batch_coords("davis", 256, 9)
coords, ptr = torch.from_numpy(pb["coords"]).to(dev), torch.from_numpy(pb["ptr"]).to(dev)
for tt, th in (("num", 30), ("dist", 4.0)):
    for _ in range(2):
        cg.residue_graph_batch(coords, ptr, th, tt, True)
torch.cuda.synchronize()
print("ok")
