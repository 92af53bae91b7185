"""N > 1 host logic on CPU (gloo, world_size 2): row sharding + global loss normalisation + gradient all-reduce
reproduce the single-process gradient.  The per-rank compute is the numpy oracle standing in for the CUDA kernels
(tests may use oracle/); what is under test is learning-..._b200/dp.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import npp_b200  # noqa: F401
    from npp_b200 import dp
    from oracle import npp_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                     # same data on every rank, each takes its shard
    params = O.init_params(rng, topk=3, width=64)
    enc = rng.standard_normal((n, 1386)).astype(np.float32)
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.4).astype(np.float32)
    a, b = dp.shard_rows(n, rank, world)
    logits, c = O.forward(params, enc[a:b])
    g = O.mse_l2_grad_logits(logits, target[a:b], mask[a:b], n_norm=n)      # GLOBAL normalisation
    grads, _ = O.backward(params, c, g)
    keys = sorted(grads)
    flat = torch.from_numpy(np.concatenate([grads[k].ravel() for k in keys]))
    dp.allreduce_sum_(flat)
    assert dp.world() == world
    if rank == 0:
        np.save(os.path.join(out_dir, "dp.npy"), flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows():
    sys.path.insert(0, ROOT)
    import npp_b200  # noqa: F401
    from npp_b200 import dp
    for n, w in [(10, 2), (11, 2), (262144, 8), (7, 8), (0, 4)]:
        parts = [dp.shard_rows(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_dp_gradient_equals_single_process(tmp_path):
    from oracle import npp_oracle as O
    n, world, port = 101, 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dp.npy")
    rng = np.random.default_rng(0)
    params = O.init_params(rng, topk=3, width=64)
    enc = rng.standard_normal((n, 1386)).astype(np.float32)
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.4).astype(np.float32)
    logits, c = O.forward(params, enc)
    grads, _ = O.backward(params, c, O.mse_l2_grad_logits(logits, target, mask))
    ref = np.concatenate([grads[k].ravel() for k in sorted(grads)])
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5


def _search_worker(rank, world, port, n_candidates, out_dir):
    sys.path.insert(0, ROOT)
    import npp_b200  # noqa: F401
    from npp_b200 import search_fits
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = search_fits.assign_candidates(n_candidates, rank, world)
    scores = search_fits.gather_scores({k: float(10 * k + rank) for k in mine})      # stand-in for the fitted distances
    if rank == 1:
        np.save(os.path.join(out_dir, "scores.npy"), np.array([scores[k] for k in range(n_candidates)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_search_candidates_sharded_over_ranks(tmp_path):
    """Independent search-stage fits over N GPUs: round-robin assignment, no data-path collective, one gather of the
    per-candidate scores at the end (gloo, world_size 2)."""
    sys.path.insert(0, ROOT)
    import npp_b200  # noqa: F401
    from npp_b200 import search_fits
    for n, w in [(9, 2), (9, 8), (3, 4), (0, 2)]:
        parts = [search_fits.assign_candidates(n, r, w) for r in range(w)]
        assert sorted(k for p_ in parts for k in p_) == list(range(n))
        assert max(len(p_) for p_ in parts) - min(len(p_) for p_ in parts) <= 1
    assert search_fits.gather_scores({0: 1.5}) == {0: 1.5}                           # single process: identity
    world, port = 2, 31000 + os.getpid() % 2000
    mp.spawn(_search_worker, args=(world, port, 9, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "scores.npy")
    assert got.tolist() == [10.0 * k + (k % 2) for k in range(9)]
