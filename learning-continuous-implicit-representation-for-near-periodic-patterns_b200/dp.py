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
    """One train step of a row-sharded batch on a Plan: forward, masked MSE normalised by the global row count,
    backward, gradient all-reduce, identical Adam on every rank.  Returns this rank's share of the loss (the sum
    over ranks is the global loss)."""

    def __init__(self, plan):
        self.plan = plan

    def __call__(self, coords, target, mask, lr, n_global, step=None):
        plan = self.plan
        logits = plan.forward(coords)
        loss, g, _ = plan.mse(logits, target, mask, n_norm=n_global)
        plan.backward(coords.shape[0], g)
        allreduce_sum_(plan.grads[: plan.trained_floats])
        plan.adam_step(lr, step=step)
        return loss
