"""Two-GPU data-parallel correctness over NCCL (needs >= 2 CUDA devices; skipped otherwise -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

SURVEY.md 8(e) / 4(v): (1) the all-reduced gradient of a global batch sharded over two ranks equals the single-process
gradient of the whole batch (all-reduce launched after the graph replay, the default for N > 1); (2) after real training
steps (dropout on, every rank its own masks, the all-reduce AND Adam captured inside each rank's CUDA graph) the replicas hold
bit-identical weights.
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import caster_dta_b200 as cg
    from caster_dta_b200 import loader, parallel, training
    from caster_dta_b200.configs import caster_dta_2_2
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        kw = caster_dta_2_2()

        def build():
            torch.manual_seed(5)
            return cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"]).to(dev)

        ds = loader.SyntheticPairDataset("tiny", 64, seed=12, edge_thresh=10)
        table = torch.from_numpy(ds.aa_table)
        mk = dict(edge_thresh=10, thresh_type="num", keep_self_loops=True, max_len=1100, max_atoms=140)
        spec = loader.BucketSpec(9, node_gran=256, atom_gran=64, mol_edge_gran=256)
        mine = list(loader.PairBatchLoader(ds, 10_000_000, 8, spec=spec, shuffle=True, seed=3, rank=rank, world_size=world, pin=False))
        # ---- (1) gradient of the sharded global batch == single-process gradient (eval mode: no dropout) ---------------------
        model = build().eval()
        parallel.broadcast_parameters(model, 0)
        opt = parallel.FlatAdam(model, lr=1e-3)
        step = training.BucketedTrainStep(model, opt, table, launch_mode="graph", update=False, **mk)
        t, m = mine[0]
        step.step({k: v.to(dev) for k, v in t.items()}, m)
        torch.cuda.synchronize()
        dp_grad = opt.flat_grad.clone()
        if rank == 0:
            solo = build().eval()
            sopt = parallel.FlatAdam(solo, lr=1e-3, process_group=None)
            sopt.world = lambda: 1                               # the whole global batch in ONE process, no collective
            gspec = loader.BucketSpec(17, node_gran=256, atom_gran=64, mol_edge_gran=256)
            whole = list(loader.PairBatchLoader(ds, 20_000_000, 16, spec=gspec, shuffle=True, seed=3, pin=False))
            tw, mw = whole[0]
            assert mw["pairs"] == 2 * m["pairs"]
            sstep = training.BucketedTrainStep(solo, sopt, table, launch_mode="eager", update=False, **mk)
            sstep.step({k: v.to(dev) for k, v in tw.items()}, mw)
            torch.cuda.synchronize()
            ref = sopt.flat_grad
            err = float((dp_grad - ref).abs().max() / ref.abs().max())
            out["grad_rel_err"] = err
        # ---- (2) replicas stay bit-identical through real steps (train mode, dropout, captured all-reduce + Adam) ---------------
        model.train()
        tstep = training.BucketedTrainStep(model, opt, table, launch_mode="graph", update=True, collective_in_graph=True, **mk)
        for t, m in mine[:4]:
            tstep.step({k: v.to(dev) for k, v in t.items()}, m)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(opt.flat_param) for _ in range(world)]
        dist.all_gather(gathered, opt.flat_param.detach())
        if rank == 0:
            out["replicas_equal"] = bool(all(torch.equal(gathered[0], g) for g in gathered[1:]))
            out["moved"] = float((gathered[0] - sopt.flat_param).abs().max())
            out["graphs"] = len(tstep.graphs)
        step.close()
        tstep.close()
    finally:
        if not parallel.shutdown():
            os._exit(0 if "replicas_equal" in out or rank != 0 else 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_and_replica_consistency():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    assert out["grad_rel_err"] <= 1e-5, f"DP gradient differs from the single-process gradient: {out['grad_rel_err']:.3e}"
    assert out["replicas_equal"], "replicas diverged"
    assert out["moved"] > 0 and out["graphs"] >= 1
