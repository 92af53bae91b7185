#!/usr/bin/env python
"""Golden vectors for the adaptive robust pixel loss, generated from the LIVE reference in the build container:

    python tests/golden/make_golden_robust.py        # needs /root/reference; writes tests/golden/golden_robust.npz

Runs models/mse_calculator.py::img2mse(x, y, 'robust_loss_adaptive', AdaptiveLossFunction(num_dims=3, float32, cpu), mask)
on seeded inputs for several latent settings and records the loss, the autograd gradients w.r.t. x and the two latent
parameters, and the reference's spline value of log Z(alpha)."""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "externel_lib"))
sys.path.insert(0, REF)
sys.modules.setdefault("torch_dct", types.ModuleType("torch_dct"))   # only used by DCT helpers
from robust_loss_pytorch.adaptive import AdaptiveLossFunction  # noqa: E402
import models.mse_calculator as ref_mse  # noqa: E402

torch.autograd.set_detect_anomaly(False)
rng = np.random.default_rng(7)
out = {}
cases = [
    ("init", np.zeros(3), np.zeros(3), 257, True),
    ("mid", np.array([-1.3, 0.4, 2.2]), np.array([-0.7, 0.1, 1.5]), 1000, True),
    ("extreme", np.array([-6.0, 6.0, 0.0]), np.array([-3.0, 3.0, -8.0]), 513, False),
]
for name, la, ls, n, use_mask in cases:
    ad = AdaptiveLossFunction(num_dims=3, float_dtype=np.float32, device="cpu")
    with torch.no_grad():
        ad.latent_alpha.copy_(torch.tensor(la, dtype=torch.float32)[None])
        ad.latent_scale.copy_(torch.tensor(ls, dtype=torch.float32)[None])
    x = torch.tensor(rng.random((n, 3), dtype=np.float32), requires_grad=True)
    y = torch.tensor(rng.random((n, 3), dtype=np.float32))
    if name == "extreme":
        y = y + torch.tensor(rng.standard_normal((n, 3)).astype(np.float32)) * 3.0   # large residuals
    mask = torch.tensor((rng.random((n, 1)) > 0.3).astype(np.float32)) if use_mask else None
    loss = ref_mse.img2mse(x, y, "robust_loss_adaptive", ad, mask)
    loss.backward()
    out[f"{name}_x"] = x.detach().numpy()
    out[f"{name}_y"] = y.numpy()
    out[f"{name}_mask"] = mask.numpy() if mask is not None else np.zeros((0, 1), np.float32)
    out[f"{name}_latent_alpha"] = la.astype(np.float32)
    out[f"{name}_latent_scale"] = ls.astype(np.float32)
    out[f"{name}_loss"] = np.float64(loss.item())
    out[f"{name}_gx"] = x.grad.numpy()
    out[f"{name}_g_latent_alpha"] = ad.latent_alpha.grad.numpy().reshape(-1)
    out[f"{name}_g_latent_scale"] = ad.latent_scale.grad.numpy().reshape(-1)
    out[f"{name}_alpha"] = ad.alpha().detach().numpy().reshape(-1)
    out[f"{name}_scale"] = ad.scale().detach().numpy().reshape(-1)
    out[f"{name}_logz"] = ad.distribution.log_base_partition_function(ad.alpha().detach().double()).numpy().reshape(-1)
al = np.linspace(0.001, 1.999, 400)
ad = AdaptiveLossFunction(num_dims=3, float_dtype=np.float32, device="cpu")
out["logz_alpha"] = al
out["logz_ref"] = ad.distribution.log_base_partition_function(torch.tensor(al)).numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_robust.npz"), **out)
print("wrote golden_robust.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "loss" in k or "g_latent" in k})
