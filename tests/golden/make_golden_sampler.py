"""Golden vectors for the GridPatchSampler surface, generated from the reference's own models/sampler.py on CPU
(run in the build container only; /root/reference is not on the GPU box):

    python tests/golden/make_golden_sampler.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NPP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def scene(seed, H=96, W=128):
    rng = np.random.default_rng(seed)
    img = rng.random((1, H, W, 3)).astype(np.float32)
    mask = np.ones((1, H, W, 1), np.float32)
    mask[0, 30:58, 44:90] = 0                                # unknown hole
    mask[0, 5:12, 100:120] = 0
    masked = img * mask
    train = np.stack(np.nonzero(mask[0, :, :, 0])[:2], 1)
    val = np.stack(np.nonzero(1 - mask[0, :, :, 0])[:2], 1)
    return masked, mask, train, val


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self           # the 'same' branch calls .cuda() (sampler.py:346)
    import models.sampler as ref
    out = {}
    cases = [("int", [[[13, 2], [-3, 15]]], 32, 2, False), ("float", [[[12.5, 2.25], [-3.0, 14.75]]], 32, 3, False),
             ("rand", [[[13, 2], [-3, 15]]], 32, 2, True)]
    for tag, shifts, ps, ns, no_reg in cases:
        masked, mask, train, val = scene(1)
        out[f"{tag}/shifts"] = np.array(shifts, np.float64)
        out[f"{tag}/meta"] = np.array([ps, ns, int(no_reg)])
        s = ref.GridPatchSampler(N_samples=ns, img=torch.Tensor(masked), mask=torch.Tensor(mask), patch_size=ps,
                                 height=masked.shape[1], width=masked.shape[2], pool_train=torch.Tensor(train),
                                 pool_val=torch.Tensor(val), selected_shifts=shifts, no_reg_sampling=no_reg)
        np.random.seed(7)
        for call in range(12):
            if call == 8:                                    # the schedule of NPP_completion/train.py:137-141
                ps, ns = ps // 2, ns * 2
                s.reset_patchsize(img=torch.Tensor(masked), mask=torch.Tensor(mask), N_samples=ns, patch_size=ps)
                s.reset_pool(torch.Tensor(train), torch.Tensor(val))
            r = s.sample_patches(topk=3, invalid_ratio=0.3)
            real, real_mask, fake, fake_mask, coords, source, k, weight = r
            out[f"{tag}/{call}/source"] = np.array(["val", "train", "same", "none"].index(source or "none"))
            out[f"{tag}/{call}/k"] = np.array(k)
            if k == 0:
                continue
            out[f"{tag}/{call}/real"] = real.numpy()
            out[f"{tag}/{call}/real_mask"] = real_mask.numpy()
            out[f"{tag}/{call}/fake"] = fake.numpy()
            out[f"{tag}/{call}/fake_mask"] = fake_mask.numpy()
            out[f"{tag}/{call}/coords"] = coords.numpy()
            if weight is not None:
                out[f"{tag}/{call}/weight"] = weight.numpy()
        out[f"{tag}/rng_after"] = np.array(np.random.randint(0, 1 << 30))   # RNG stream position after the calls
    np.savez_compressed(os.path.join(HERE, "golden_sampler.npz"), **out)
    print("written", len(out), "arrays")


if __name__ == "__main__":
    main()
