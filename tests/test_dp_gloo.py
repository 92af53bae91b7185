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
