"""Diagnostic (not a test): a bare loop of fused train steps at the bench shape, for ncu captures.
    python tests/diag_step_loop.py [steps] [rows] [topk]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_parity_gpu import make  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
topk = int(sys.argv[3]) if len(sys.argv) > 3 else 3
plan, params, coords, tabs, freqs, rng = make(topk, n)
target = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
mask = torch.ones(n, 1, device="cuda")
cd = torch.from_numpy(coords).cuda()
loss = torch.zeros((), device="cuda")
for step in range(1, steps + 1):
    plan.train_step(cd, target, mask, 5e-4, loss, step=step)
torch.cuda.synchronize()
print("loss", loss.item())
