"""Data parallelism inside one image (BASELINE.json cfg 4): the coordinate rows of a step are
split evenly over the ranks, every rank holds the full weights and Adam state, the loss is normalised by the GLOBAL
pixel count and the flat fp32 gradient arena is summed with one all-reduce (NCCL over NVLink/NVSwitch on the GPU
box, gloo in the CPU tests).  Independent fits (image x proposal) need none of this: they run one per GPU."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int):
    """Contiguous, balanced [start, stop) row range of `rank`; the first n % world ranks get one extra row."""
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of a flat gradient buffer (no-op for a single process)."""
    if world() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class DataParallelStep:
    """One train step of a row-sharded batch on a Plan: the fused forward + head + backward chain on this rank's rows
    (loss normalised by the global row count), then the weight gradients in a few layer groups, each group's slice of
    the flat fp32 gradient arena all-reduced (NCCL, asynchronously on its own stream) while the next group's GEMMs
    run, then identical Adam on every rank.  Returns this rank's share of the loss (the sum over ranks is the global
    loss) as a device scalar."""

    def __init__(self, plan, n_buckets: int = 3):
        self.plan = plan
        nl = plan.layer_count()
        n_buckets = max(1, min(int(n_buckets), nl))
        # last layers first (any order is correct: the whole backward chain has run before the first group starts)
        edges = [round(nl * k / n_buckets) for k in range(n_buckets + 1)]
        self.groups = [(edges[k], edges[k + 1]) for k in range(n_buckets - 1, -1, -1)]
        self.buckets = [plan.grad_range(lb, le) for lb, le in self.groups]
        self.loss = torch.zeros((), device=plan.device)
        self.launches = 0
        self.last_loss = self.loss
        self.skip_allreduce = False       # timing experiments only: how much of a step is exposed communication

    def __call__(self, coords, target, mask, lr, n_global, step=None):
        plan = self.plan
        n = coords.shape[0]
        plan.step_forward_backward(coords, target, mask, n_norm=n_global)
        pending = []
        multi = world() > 1 and not self.skip_allreduce
        for (lb, le), bucket in zip(self.groups, self.buckets):
            plan.step_wgrad(lb, le, n, n_global)
            if multi:
                pending.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True))
        for w in pending:
            w.wait()                      # the compute stream waits for the collective, the host does not
        plan.step_finish(n_global, lr, self.loss, step=step)
        self.launches = plan.launch_count()
        self.last_loss = self.loss
        return self.loss
