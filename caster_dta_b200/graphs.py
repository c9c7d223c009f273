"""CUDA-graph capture of a fixed-shape step (forward + backward) of the accelerated model.

After the GVP stack is fused, a CASTER-DTA training step is ~500 small launches and CPU-launch-bound; replaying it as
one CUDA graph makes it GPU-bound.  Every C-ABI entry point is capture-safe (no allocation, no synchronisation, all
work on the caller's stream), torch's caching allocator serves the workspaces from the graph's private pool, and the
dropout masks keep following torch's (graph-aware) Philox stream.

Usage: shapes must be fixed -- copy each batch into the `static` tensors the step closes over, then `replay()`.
"""
import torch


class GraphedStep:
    def __init__(self, fn, grads=None, warmup=3):
        """`fn()` runs one step (forward + backward) on static input tensors and returns a tensor (e.g. the loss).
        `grads` (a `parallel.GradSync`) makes the captured backward ASSIGN the parameter gradients: `.grad` is None while
        capturing, the tensors autograd creates are this graph's outputs, and `select()` re-attaches them to the
        parameters when several graphs (e.g. one per input buffer) share one model."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                if grads is not None:
                    grads.reset()
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        if grads is not None:
            grads.reset()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()
        self.grads = grads
        self.grad_tensors = None if grads is None else [p.grad for p in grads.params]

    def select(self):
        if self.grads is not None:
            for p, g in zip(self.grads.params, self.grad_tensors):
                p.grad = g

    def replay(self):
        self.graph.replay()
        return self.out
