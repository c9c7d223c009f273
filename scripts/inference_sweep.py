#!/usr/bin/env python
"""BASELINE config 4: BindingDB-scale inference sweep, pairs sharded over the GPUs with no communication.

    python scripts/inference_sweep.py --pairs 20000 [--graph dist4|knn30] [--ligands-per-protein 1]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/inference_sweep.py ...

Each rank takes pairs r, r+W, ... (caster_dta_b200.parallel.shard_pairs semantics), builds the residue graphs ON THE GPU
from synthetic BindingDB-shape backbones (featurizer kernel, inside the timed region), runs CASTER-DTA(2,2) in eval mode
through the public module API and keeps the predictions on the device.  A pool of distinct batches is cycled to stand in
for the full pair list.  `--ligands-per-protein L > 1` enables the unique-protein embedding cache: the
protein encoder runs once per protein batch and its residue embeddings are reused for the next L-1 ligand batches.
Prints one JSON line (rank 0): pairs/s over all ranks (max-over-ranks device time).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import caster_dta_b200 as cg
from caster_dta_b200 import synth
from caster_dta_b200 import joint
from caster_dta_b200.configs import caster_dta_2_2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20000, help="pairs in the whole sweep (all ranks)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--pool", type=int, default=16, help="distinct synthetic batches cycled per rank")
    ap.add_argument("--graph", default="dist4", choices=["dist4", "knn30"])
    ap.add_argument("--ligands-per-protein", type=int, default=1)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    kw = caster_dta_2_2()
    torch.manual_seed(9)
    model = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"]).to(dev).eval()
    thresh, ttype = (4.0, "dist") if args.graph == "dist4" else (30, "num")
    pool = []
    for b in range(args.pool):
        pb = synth.protein_batch_coords("bindingdb", args.batch, 1000 * rank + b, self_avoiding=False)
        mol = synth.molecule_batch(args.batch, 1000 * rank + b)
        t = lambda a: torch.from_numpy(a).to(dev)
        pool.append(dict(coords=t(pb["coords"]), ptr=t(pb["ptr"]), x_s=t(pb["x_s"]), x_v=t(pb["x_v"]), nt=t(pb["ntypes"]),
                         batch=t(pb["batch"]), max_res=int(np.diff(pb["ptr"]).max()),
                         mol={k: t(v) for k, v in mol.items()}, max_atoms=int(np.bincount(mol["batch"]).max())))
    aa_table = torch.rand(20, 11, generator=torch.Generator().manual_seed(5)).to(dev)      # synthetic residue property table
    my_batches = len(range(rank, args.pairs // args.batch, world))
    preds, edges, residues = [], 0, 0

    def run(nb, count):
        nonlocal edges, residues
        embed, dp, dense = None, None, None
        for i in range(nb):
            d = pool[i % len(pool)]
            m = d["mol"]
            molg = dict(x=m["x"], edge_index=m["edge_index"], ntypes=m["ntypes"], etypes=m["etypes"], eattr=m["eattr"],
                        batch=m["batch"], num_graphs=args.batch, max_nodes=d["max_atoms"])
            if embed is None or i % args.ligands_per_protein == 0:
                dp = d                                   # the protein batch the next L ligand batches are paired with
                # backbone coordinates -> node features -> residue graph -> encoder, all on the device
                pgb = cg.protein_graph_batch(d["coords"], d["ptr"], d["nt"], aa_table, thresh, ttype, True)
                ei = pgb["edge_index"]
                embed = model.protein_gnn(**pgb)
                dense = None
                if count:
                    edges += int(ei.shape[1]); residues += int(d["x_s"].shape[0])
            if dense is None:
                dense = joint.DenseIndex(dp["batch"], int(dp["batch"].shape[0]), num_graphs=args.batch, max_nodes=dp["max_res"])
            prot = dict(batch=dp["batch"], num_graphs=args.batch, max_nodes=dp["max_res"], protein_embed=embed, dense_index=dense)
            pred, _ = model(prot, molg)
            if count:
                preds.append(pred)

    with torch.no_grad():
        run(min(4, my_batches), False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(my_batches, True)
        b.record()
        torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([my_batches * args.batch, edges, residues], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"workload": f"CASTER-DTA(2,2) inference sweep, bindingdb-shape, graph {args.graph}, batch {args.batch}",
                          "n_gpus": world, "pairs": int(cnt[0]), "ms": ms, "pairs_per_s": float(cnt[0]) / (ms * 1e-3),
                          "protein_graphs_built": int(cnt[2]), "protein_edges": int(cnt[1]),
                          "ligands_per_protein": args.ligands_per_protein,
                          "finite": bool(torch.isfinite(torch.cat(preds)).all()),
                          "note": "node + edge featurizer from backbone coordinates, encoder, cross-attention, head per batch; coordinates resident, no collective"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
