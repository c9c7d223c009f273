"""CPU-side profile of the eager inference step with a cached protein embedding (launch-bound regime)."""
import os, sys, json, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import caster_dta_b200 as cg
from caster_dta_b200 import synth
dev = torch.device("cuda")
from caster_dta_b200.configs import caster_dta_2_2
kw = caster_dta_2_2()
model = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"]).to(dev).eval()
nres = torch.randint(300, 1000, (32,))
batch = torch.repeat_interleave(torch.arange(32), nres).to(dev)
embed = torch.randn(int(nres.sum()), 64, device=dev)
mol = {k: torch.from_numpy(v).to(dev) for k, v in synth.molecule_batch(32, seed=1).items()}
molg = dict(x=mol["x"], edge_index=mol["edge_index"], ntypes=mol["ntypes"], etypes=mol["etypes"], eattr=mol["eattr"], batch=mol["batch"],
            num_graphs=32, max_nodes=int(torch.bincount(mol["batch"]).max()))
prot = dict(batch=batch, num_graphs=32, max_nodes=int(nres.max()), protein_embed=embed)
with torch.no_grad():
    for _ in range(5):
        model(prot, molg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        model(prot, molg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"cpu issue {1e3 * (t1 - t0) / 100:.3f} ms/batch, incl. drain {1e3 * (t2 - t0) / 100:.3f} ms/batch")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        model(prot, molg)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
