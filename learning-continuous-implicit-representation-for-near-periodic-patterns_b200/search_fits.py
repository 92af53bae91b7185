"""Concurrent search-stage fits: the candidate loop of NPP_proposal/search.py:85-148 run side by side.

The reference fits one fresh NPP_Net_light per candidate periodicity, one after the other, with the same seeds
(search.py:91-92), i.e. every candidate sees the same sequence of pixel batches.  A fit at 2048 rows keeps only 16 of
the 148 SMs busy and is bound by the latency of its dependent kernels, so the candidates are independent work that
fits on the GPU at the same time: one plan, one CUDA stream and one host thread per candidate (the C ABI releases the
GIL for the whole `npp_fit_run` call, so the threads enqueue kernels in parallel).  There is no data-path collective.

Only the ``--loss_type l2`` variant of that loop is covered (see `run_fits`); the default adaptive robust loss carries
trained latents across candidates and stays with the drop-in modules.  Scoring the fitted candidates (LPIPS +
contextual loss on the held-out region, search.py:150-196) stays with the caller; `Plan.forward` / `NPP_Net_light.forward` under ``torch.no_grad()`` renders the pixels it needs.
"""
from __future__ import annotations

import os
import threading
from typing import List, Optional, Sequence

import torch

import torch.distributed as dist

from .plan import MODEL_LIGHT, Plan


def assign_candidates(n_candidates: int, rank: int, world: int) -> List[int]:
    """Candidates of `rank` when a search is spread over several GPUs (one process per GPU): round-robin, so that the
    first candidates -- the most plausible periodicities, NPP_proposal/search.py ranks them in that order -- land on
    different GPUs.  Fits are independent; nothing is exchanged while they run."""
    return list(range(int(rank), int(n_candidates), int(world)))


def gather_scores(local: dict) -> dict:
    """{candidate index: score} of every rank merged on every rank (the search's ranking step,
    search.py:199-207, needs all distances); a plain dict in a single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, dict(local))
    merged = {}
    for part in parts:
        for k, v in part.items():
            if k in merged:
                raise ValueError(f"candidate {k} was fitted by two ranks")
            merged[k] = v
    return merged


def gather_batches(image: torch.Tensor, train_coords: torch.Tensor, indices: torch.Tensor):
    """Batches of a whole fit: image [H, W, 3] fp32 CUDA, train_coords [P, 2] (row, col), indices [iters, n] int64 into
    train_coords (the `select_inds` of search.py:116, drawn by the caller -- host RNG for parity with the script, any
    device RNG otherwise).  Returns (coords_all [iters, n, 2] fp32, target_all [iters, n, 3] fp32)."""
    tc = train_coords.to(image.device).long()
    sel = tc[indices.to(image.device)]                                   # [iters, n, 2]
    target = image[sel[..., 0], sel[..., 1], :].contiguous()             # search.py:118
    return sel.float().contiguous(), target.float()


def _run_fits_grouped(plans, coords_all, target_all, per_plan, iters, lrate, lrate_decay, losses):
    import ctypes as C
    from . import _native as nat
    k = len(plans)
    cs = [coords_all[i] if per_plan else coords_all for i in range(k)]
    ts = [target_all[i] if per_plan else target_all for i in range(k)]
    n = int(cs[0].shape[1])
    for c, t in zip(cs, ts):
        if not (c.is_cuda and c.dtype == torch.float32 and c.is_contiguous() and tuple(c.shape) == (iters, n, 2) and
                t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (iters, n, 3)):
            raise ValueError("run_fits needs contiguous fp32 CUDA tensors [iters, n, 2] and [iters, n, 3]")
    if any(n > p.max_rows for p in plans):
        raise ValueError(f"{n} rows exceed a plan's capacity")
    P = C.c_void_p
    handles = (P * k)(*[p.handle for p in plans])
    cptr = (P * k)(*[c.data_ptr() for c in cs])
    tptr = (P * k)(*[t.data_ptr() for t in ts])
    lptr = (P * k)(*[losses[i].data_ptr() for i in range(k)])
    first = (C.c_int64 * k)(*[p.adam_steps + 1 for p in plans])
    nat.check(plans[0].lib.npp_multi_fit_run(handles, k, cptr, tptr, None, n, iters, lrate, 0.1, float(lrate_decay) * 100.0,
                                             0.9, 0.999, 1e-8, first, lptr, nat.current_stream()))
    for p in plans:
        p.adam_steps += iters


def run_fits(plans: Sequence[Plan], coords_all: torch.Tensor, target_all: torch.Tensor, *, lrate: float = 5e-4,
             lrate_decay: float = 500, streams: Optional[Sequence[torch.cuda.Stream]] = None,
             threads: bool = True, loss_type: str = "l2", grouped: Optional[bool] = None) -> torch.Tensor:
    """Fit every plan on the same batches (or on its own, if coords_all / target_all are lists), concurrently.
    Returns the losses [len(plans), iters]; the call returns once the current stream waits for every fit.

    Default execution (``grouped``, NPP_Net_light plans, no caller-supplied streams): npp_multi_fit_run -- one train step
    of every plan captured as parallel branches of ONE CUDA graph that is launched `iters` times, the batch index and
    Adam's scalars following a device-side step counter.  ``grouped=False`` (or NPP_FIT_MULTI=0, or explicit streams):
    one plan, stream and host thread per candidate, each enqueueing its own kernels (npp_fit_run).

    Equivalent to the reference loop only for ``--loss_type l2``: npp_fit_run trains with sigmoid + masked MSE
    (img2mse(pred, gt, 'l2', ...), NPP_proposal/search.py:133).  The scripts' DEFAULT is 'robust_loss_adaptive'
    (options/arg_config.py:34), whose latent alpha / scale are trained along with the network and carried from one
    candidate to the next (the module-level adaptive_pix): that objective can rank candidates differently, so asking for
    it here raises instead of silently fitting another loss.  Run the default loss through the drop-in modules
    (models.helpers.create_npp_net + img2mse), which evaluate it with the fused adaptive-loss kernel."""
    if loss_type != "l2":
        raise NotImplementedError(
            f"run_fits fits with loss_type='l2' only (got {loss_type!r}); the adaptive robust loss of the reference's "
            "default configuration goes through models.helpers.create_npp_net + models.mse_calculator.img2mse")
    k = len(plans)
    per_plan = isinstance(coords_all, (list, tuple))
    iters = int((coords_all[0] if per_plan else coords_all).shape[0])
    dev = plans[0].device
    losses = torch.zeros(k, iters, device=dev)
    if grouped is None:
        grouped = os.environ.get("NPP_FIT_MULTI", "1") != "0"
    if grouped and k >= 1 and iters >= 1 and streams is None and all(p.model == MODEL_LIGHT for p in plans):
        _run_fits_grouped(plans, coords_all, target_all, per_plan, iters, lrate, lrate_decay, losses)
        return losses
    streams = list(streams) if streams is not None else [torch.cuda.Stream(device=dev) for _ in range(k)]
    cur = torch.cuda.current_stream(dev)
    errors: List[BaseException] = []

    def work(i):
        try:
            torch.cuda.set_device(dev)
            streams[i].wait_stream(cur)
            plans[i].fit_run(coords_all[i] if per_plan else coords_all, target_all[i] if per_plan else target_all,
                             lrate=lrate, lrate_decay=lrate_decay, losses=losses[i], stream=streams[i].cuda_stream)
        except BaseException as e:   # surfaced in the calling thread below
            errors.append(e)

    if threads and k > 1:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(k)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    else:
        for i in range(k):
            work(i)
    if errors:
        raise errors[0]
    for s in streams:
        cur.wait_stream(s)
    return losses
