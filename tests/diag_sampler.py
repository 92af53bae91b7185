"""GridPatchSampler timing on a 512 x 512 completion scene: construction
(reset_patchsize + reset_pool) and one training iteration's sampling -- np.random.choice of N_rand pixels plus
sample_patches(topk=3) -- with patch_size 64, 2 patches, lattice shifts of the synthetic texture.
Ours on the chosen device; the reference's own sampler on the CPU when its checkout is reachable (build container).

    python tests/diag_sampler.py [cpu|cuda]
"""
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")
REF = os.environ.get("NPP_REFERENCE", "/root/reference")


def scene(H=512, W=512, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.random((1, H, W, 3)).astype(np.float32)
    mask = np.ones((1, H, W, 1), np.float32)
    mask[0, H // 4:H // 4 + H // 2, W // 4:W // 4 + W // 2] = 0          # centred hole, 25 % of the area
    train = np.stack(np.nonzero(mask[0, :, :, 0])[:2], 1)
    val = np.stack(np.nonzero(1 - mask[0, :, :, 0])[:2], 1)
    return img * mask, mask, train, val


def timeit(cls, device, iters=20):
    masked, mask, train, val = scene()
    T = lambda a: torch.Tensor(a).to(device)  # noqa: E731
    shifts = [[[42, 5], [-5, 38]], [[21, 3], [-3, 19]], [[84, 10], [-10, 76]]]
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
    kw = dict(N_samples=2, img=T(masked), mask=T(mask), patch_size=64, height=512, width=512, pool_train=T(train),
              pool_val=T(val), selected_shifts=shifts, no_reg_sampling=False)
    cls(**kw)                                   # first construction pays torch's lazy initialisation
    sync()
    np.random.seed(0)
    t0 = time.perf_counter()
    s = cls(**kw)
    sync()
    t_build = time.perf_counter() - t0
    i_train = T(train)
    for _ in range(2):
        s.sample_patches(topk=3, invalid_ratio=0.3)
    sync()
    t_pix = t_patch = 0.0
    for _ in range(iters):
        t0 = time.perf_counter()
        sel = np.random.choice(train.shape[0], size=[8192], replace=False)      # NPP_completion/train.py:154-156
        coords = i_train[sel].long()
        sync()
        t1 = time.perf_counter()
        s.sample_patches(topk=3, invalid_ratio=0.3)
        sync()
        t2 = time.perf_counter()
        t_pix += t1 - t0
        t_patch += t2 - t1
    return t_build * 1e3, t_pix / iters * 1e3, t_patch / iters * 1e3


def main():
    device = sys.argv[1] if len(sys.argv) > 1 else ("cuda" if torch.cuda.is_available() else "cpu")
    sys.path.insert(0, PKG)
    from models.sampler import GridPatchSampler as Ours
    b, p, s = timeit(Ours, device)
    print(f"ours ({device}): construction {b:.1f} ms; per iteration: pixel draw (host np.random.choice + gather) {p:.2f} ms, "
          f"sample_patches {s:.2f} ms")
    if os.path.isdir(os.path.join(REF, "models")):
        for name in ("models", "models.sampler"):
            sys.modules.pop(name, None)
        sys.path.remove(PKG)
        sys.path.insert(0, REF)
        for name in ["torch_dct"]:
            sys.modules.setdefault(name, types.ModuleType(name))
        import importlib
        ref = importlib.import_module("models.sampler")
        torch.autograd.set_detect_anomaly(False)
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self           # the reference calls .cuda() (sampler.py:346)
        try:
            b, p, s = timeit(ref.GridPatchSampler, "cpu", iters=5)
        finally:
            torch.Tensor.cuda = orig_cuda
        print(f"reference (cpu, {os.cpu_count()} cores): construction {b:.1f} ms; per iteration: pixel draw {p:.2f} ms, "
              f"sample_patches {s:.2f} ms")


if __name__ == "__main__":
    main()
