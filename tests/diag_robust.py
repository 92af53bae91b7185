"""GPU diagnostic (not a pytest): adaptive robust pixel loss, fused CUDA pass vs the same arithmetic as torch ops
(forward + backward, CUDA events, 200 iterations after warm-up)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import npp_b200  # noqa: F401
from npp_b200.robust_loss import NppAdaptiveLoss

ad = NppAdaptiveLoss(3, device="cuda")
for n in (16384, 59392, 262144):
    x = torch.rand(n, 3, device="cuda", requires_grad=True)
    y = torch.rand(n, 3, device="cuda")
    m = (torch.rand(n, 1, device="cuda") > 0.3).float()

    def fused():
        ad.zero_grad(); x.grad = None
        ad.fused_img2mse(x, y, m).backward()

    def torch_ops():
        ad.zero_grad(); x.grad = None
        d = x - y
        d = d * m + (1 - m) * d * 0.3
        torch.mean(torch.mean(ad.lossfun(d))).backward()

    for name, fn in (("fused CUDA pass", fused), ("torch ops", torch_ops)):
        for _ in range(20):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        print(f"n={n:7d} {name:16s}: {us:8.1f} us per loss fwd+bwd   ({40.0 * n / us / 1e3:7.1f} GB/s of 40 B/row algorithmic)", flush=True)
