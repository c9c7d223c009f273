"""Data parallelism for the GVP path: only the mini-batch of protein-ligand pairs shards.

One process per GPU.  Inference is embarrassingly parallel (no communication).  Training does ONE all-reduce per
step over a flat fp32 gradient bucket: every parameter's `.grad` is a view into the bucket, so backward writes
the gradients in place and the collective follows with no repacking.  Works with the `nccl` backend on GPUs
(NVLink 5 / NVSwitch) and with `gloo` on CPU (tests).
"""
import torch
import torch.distributed as dist


def shard_pairs(num_pairs, rank, world_size):
    """Indices of the pairs rank `rank` owns: r, r+W, r+2W, ... (round-robin keeps per-rank edge counts balanced
    when pairs are sorted by size)."""
    return list(range(rank, num_pairs, world_size))


def shard_by_cost(costs, world_size, equal_counts=False):
    """Greedy longest-processing-time partition of pairs by a cost (e.g. protein edge count), the way the
    reference's batch sampler balances batches by edge count (`dataset/dual_dataset.py:476-516`).

    `equal_counts`: every rank gets the same number of pairs (+-1), which keeps the per-rank batch shape -- and with it
    the captured CUDA graph -- the same on all ranks: the costliest remaining pair goes to the least-loaded rank that
    still has a free seat."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads, parts = [0] * world_size, [[] for _ in range(world_size)]
    seats = [len(costs) // world_size + (1 if r < len(costs) % world_size else 0) for r in range(world_size)]
    for i in order:
        free = [k for k in range(world_size) if not equal_counts or len(parts[k]) < seats[k]]
        r = min(free, key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


class FlatGradBucket:
    """All trainable parameters' gradients as views of one contiguous buffer + a single all-reduce."""

    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad and p.numel() > 0]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self.numel = total

    def zero(self):
        self.flat.zero_()

    def check_views(self):
        """Autograd accumulates into an existing .grad in place; verify nobody replaced the views."""
        base = self.flat.data_ptr()
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + off * self.flat.element_size():
                return False
            off += p.numel()
        return True

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat


class GradSync:
    """Gradient handling without the per-parameter accumulate kernels.

    `.grad` is reset to None before every backward, so autograd hands each parameter its freshly computed gradient
    tensor instead of launching one `add_` per parameter into a pre-zeroed buffer (158 tiny kernels per step for
    CASTER-DTA(2,2)).  With more than one rank the gradients are flattened into ONE buffer (a single `cat`), all-reduced
    once over NCCL / NVLink (3.06 MB) and copied back with one multi-tensor copy.  Works under CUDA-graph capture: the
    gradient tensors created while capturing are the graph's own allocations and are rewritten in place on replay."""

    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad and p.numel() > 0]
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)

    def reset(self):
        for p in self.params:
            p.grad = None

    def all_reduce_mean(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        torch._foreach_copy_([g.view(-1) for g in grads], list(flat.split([g.numel() for g in grads])))


class FlatAdam:
    """Parameters, gradients and Adam moments of a module as FOUR flat fp32 buffers.

    * every parameter becomes a view of `flat_param`; the optimizer therefore updates the whole model with ONE fused
      Adam kernel over one tensor (Adam is element-wise, so this is the per-parameter update bit for bit);
    * `.grad` is reset to None before each backward, so autograd ASSIGNS every gradient (no per-parameter accumulate
      kernels); `sync()` packs them into `flat_grad` with one multi-tensor copy, runs ONE all-reduce (sum) over NCCL /
      NVLink on the flat buffer, and the optimizer reads the reduced buffer in place -- no `cat`, no copy back;
    * every call is capturable, so zero-grad + forward + backward + pack + all-reduce + Adam replay as one CUDA graph.

    The loss is expected to be normalised by the GLOBAL number of pairs (each rank's weights are 1 / global pairs), so
    the all-reduce is a plain sum and the result equals the single-process gradient of the global batch."""

    ALIGN = 128          # floats

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None, capturable=True):
        self.params = [p for p in module.parameters() if p.requires_grad and p.numel() > 0]
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        ref = self.params[0]
        # every parameter starts on a 512-byte boundary of the flat buffers (what the caching allocator would give it): the
        # kernels' 16-byte vector loads of weights / LayerNorm vectors keep working; the padding stays zero
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat_param = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.grad_views = []
        for p, off in zip(self.params, offs):
            n = p.numel()
            self.flat_param[off: off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off: off + n].view_as(p)
            self.grad_views.append(self.flat_grad[off: off + n].view_as(p))
        self.master = torch.nn.Parameter(self.flat_param)
        self.master.grad = self.flat_grad
        fused = ref.is_cuda
        self.opt = torch.optim.Adam([self.master], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                    fused=fused, capturable=capturable and fused)
        # create the optimizer state now (a step on a zero gradient moves nothing; the step counter is put back to 0), so
        # that a CUDA-graph capture never meets the lazy state initialisation
        if weight_decay == 0.0:
            self.opt.step()
            st = self.opt.state[self.master]["step"]
            if torch.is_tensor(st):
                st.zero_()
            else:
                self.opt.state[self.master]["step"] = 0

    def reset(self):
        for p in self.params:
            p.grad = None

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def sync(self, collective=True):
        """Pack the freshly assigned gradients into the flat buffer and sum them over the ranks."""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.grad_views, grads)
        if collective and self.world() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat_grad

    def all_reduce(self):
        """The collective alone (sum of the packed flat gradient over the ranks)."""
        if self.world() > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def step(self):
        self.master.grad = self.flat_grad
        self.opt.step()

    def grads(self):
        """Per-parameter views of the (reduced) flat gradient, in `self.params` order."""
        return self.grad_views


def shutdown(timeout=20.0):
    """Destroy the default process group without risking a hang at exit: the caller first drops every CUDA graph that
    captured a collective (`BucketedTrainStep.close()`); the destruction itself runs in a helper thread and is abandoned
    after `timeout` seconds."""
    if not (dist.is_available() and dist.is_initialized()):
        return True
    import threading
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout)
    return not t.is_alive()


def broadcast_parameters(module, src=0, process_group=None):
    """Replicas start from rank `src`'s weights."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        if t.numel():
            dist.broadcast(t.data, src, group=process_group)
