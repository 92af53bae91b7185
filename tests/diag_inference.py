"""Full-image inference timing: a 512 x 512 image through the joint NPP_Net (K=3),
(a) Plan.render_into (chunked forward + sigmoid + scatter fused into the head kernel),
(b) the evaluation loop of the reference scripts on the drop-in surface: 20 000-row chunks, sigmoid, index_put
    (NPP_completion/train.py:277-309).
Not a test; run on the GPU box: python tests/diag_inference.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npp_b200  # noqa: E402,F401
from npp_b200.plan import EncoderSpec, Plan  # noqa: E402

RES = (512, 512)
rng = np.random.default_rng(0)
freqs = (rng.standard_normal(10) * 10).astype(np.float32)
enc = EncoderSpec.from_proposals(RES, [[97.0, 187.0]] * 3, [[42.7, 38.4], [21.35, 19.2], [85.4, 76.8]], freqs)
plan = Plan(enc, max_rows=1 << 15, training=False)
plan.reset_parameters(0)
yy, xx = torch.meshgrid(torch.arange(RES[0], device="cuda"), torch.arange(RES[1], device="cuda"), indexing="ij")
coords = torch.stack([yy.reshape(-1), xx.reshape(-1)], 1).float().contiguous()
n = coords.shape[0]
image = torch.zeros(1, RES[0], RES[1], 3, device="cuda")


def fused():
    plan.render_into(coords, image)


def script_loop(chunk=20000):
    for i in range(0, n, chunk):
        c = coords[i:i + chunk]
        pred = torch.sigmoid(plan.forward(c))
        cl = c.long()
        image[:, cl[:, 0], cl[:, 1], :] = pred


for name, fn in (("render_into (fused, wave-aligned 18944-row chunks)", fused), ("script loop (20000-row chunks, sigmoid, index_put)", script_loop)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    flop = n * 21090816 / 3          # forward third of the train-step FLOPs per sample
    print(f"{name}: {ms:.3f} ms per 512x512 image  {n / ms / 1e3:.1f} Mpx/s  {flop / ms / 1e9:.0f} TFLOP/s")
