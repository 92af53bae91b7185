"""Diagnostic (not a test): steady-state ms per fused train step at the bench shape with per-class timing, after a
warm-up long enough for the power cap to settle.   python tests/diag_step_time.py [rows] [topk] [seconds]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_parity_gpu import make  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
topk = int(sys.argv[2]) if len(sys.argv) > 2 else 3
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
plan, params, coords, tabs, freqs, rng = make(topk, n)
NB = 4
cds = [torch.from_numpy(np.stack([rng.integers(0, 512, n), rng.integers(0, 512, n)], 1).astype(np.float32)).cuda() for _ in range(NB)]
target = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
mask = torch.ones(n, 1, device="cuda")
loss = torch.zeros((), device="cuda")


def step(i):
    plan.prefetch_encode(cds[(i + 1) % NB])
    plan.train_step(cds[i % NB], target, mask, 5e-4, loss)


t0 = time.time()
i = 0
while time.time() - t0 < secs:
    step(i)
    i += 1
    if i % 50 == 0:
        torch.cuda.synchronize()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 400
e0.record()
for j in range(K):
    step(i + j)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
plan.profile(True)
for j in range(K):
    step(i + K + j)
prof = plan.profile_read()
plan.profile(False)
cls = {k: round(v[0] / K * 1e3, 1) for k, v in prof.items() if v[0] > 0}
print(f"{os.environ.get('TAG', '')} rows {n} topk {topk}: {ms * 1e3:.1f} us/step = {n / ms / 1e3:.2f} M samples/s; classes (us) {cls}")
