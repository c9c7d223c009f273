"""world_size-2 gloo test (CPU) of the data-parallel plumbing: flat gradient bucket + one all-reduce gives every
replica the mean gradient, identical bits on both ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caster_dta_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_sync(rank, world, port, ret):
    """Same check for parallel.GradSync (gradients assigned by autograd, one cat + all-reduce + multi-tensor copy)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    parallel.broadcast_parameters(model, 0)
    sync = parallel.GradSync(model)
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    for _ in range(2):                                  # twice: reset() must drop the previous step's gradients
        sync.reset()
        model(data[parallel.shard_pairs(8, rank, world)]).square().mean().backward()
        sync.all_reduce_mean()
    flat = torch.cat([p.grad.flatten() for p in sync.params])
    ref_model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    ref_model.load_state_dict(model.state_dict())
    total = sum(ref_model(data[parallel.shard_pairs(8, r, world)]).square().mean() for r in range(world)) / world
    total.backward()
    ref = torch.cat([p.grad.flatten() for p in ref_model.parameters()])
    ret[rank] = (float((flat - ref).abs().max()), flat.tolist())
    dist.destroy_process_group()


def test_grad_sync_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_sync, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] < 1e-6 and ret[1][0] < 1e-6
    assert ret[0][1] == ret[1][1], "replicas must hold bit-identical reduced gradients"


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123 + rank)                      # replicas start different on purpose
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    parallel.broadcast_parameters(model, 0)
    bucket = parallel.FlatGradBucket(model)
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    mine = parallel.shard_pairs(8, rank, world)
    bucket.zero()
    loss = model(data[mine]).square().mean()
    loss.backward()
    assert bucket.check_views()
    flat = bucket.all_reduce_mean().clone()
    # single-process reference: mean of the two shard losses
    ref_model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    ref_model.load_state_dict(model.state_dict())
    total = sum(ref_model(data[parallel.shard_pairs(8, r, world)]).square().mean() for r in range(world)) / world
    total.backward()
    ref = torch.cat([p.grad.flatten() for p in ref_model.parameters()])
    ret[rank] = (float((flat - ref).abs().max()), flat.tolist())
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] < 1e-6 and ret[1][0] < 1e-6
    assert ret[0][1] == ret[1][1], "replicas must hold bit-identical reduced gradients"


def test_shard_by_cost_balances():
    costs = [100, 90, 20, 20, 15, 10, 5, 5]
    parts = parallel.shard_by_cost(costs, 2)
    loads = [sum(costs[i] for i in p) for p in parts]
    assert sorted(i for p in parts for i in p) == list(range(8))
    assert abs(loads[0] - loads[1]) <= 15


def _worker_flat_adam(rank, world, port, ret):
    """parallel.FlatAdam over gloo: globally-normalised shard losses, SUM all-reduce on the flat buffer, one fused update;
    the replicas end bit-identical and equal to a single process that trains on the whole batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(321 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    parallel.broadcast_parameters(model, 0)
    start = {k: v.clone() for k, v in model.state_dict().items()}
    opt = parallel.FlatAdam(model, lr=1e-2, capturable=False)
    data = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    mine = parallel.shard_pairs(8, rank, world)
    for _ in range(3):
        opt.reset()
        (model(data[mine]).square().sum() / 8).backward()           # weights 1 / GLOBAL batch
        opt.sync()
        opt.step()
    ref_model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    ref_model.load_state_dict(start)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-2)
    for _ in range(3):
        ref_opt.zero_grad(set_to_none=True)
        ref_model(data).square().mean().backward()
        ref_opt.step()
    ref = torch.cat([p.detach().flatten() for p in ref_model.parameters()])
    mine_flat = torch.cat([p.detach().flatten() for p in opt.params])
    ret[rank] = (float((mine_flat - ref).abs().max()), opt.flat_param.detach().tolist())
    dist.destroy_process_group()


def test_flat_adam_world2_matches_single_process():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_flat_adam, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] < 1e-6 and ret[1][0] < 1e-6
    assert ret[0][1] == ret[1][1], "replicas must hold bit-identical weights"
