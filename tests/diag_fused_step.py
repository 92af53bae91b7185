"""Diagnostic (not a test): the fused train step (forward + head + backward in one chain launch, delayed gradient
scale) against the r01 split step (NPP_SPLIT_STEP=1) on identical weights and batches: losses, gradients, weights,
and ms per step.

    python tests/diag_fused_step.py [rows] [topk] [steps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_parity_gpu import make, rel  # noqa: E402


def run(split, n, topk, steps, timing_steps):
    os.environ["NPP_SPLIT_STEP"] = "1" if split else "0"
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    plan.keep_grads(True)
    rng = np.random.default_rng(5)
    target = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
    mask = torch.from_numpy((rng.random((n, 1)) > 0.3).astype(np.float32)).cuda()
    cd = torch.from_numpy(coords).cuda()
    loss = torch.zeros((), device="cuda")
    losses, grads = [], []
    for step in range(1, steps + 1):
        plan.train_step(cd, target, mask, 5e-4, loss, step=step)
        losses.append(loss.item())
        grads.append({k: v.clone() for k, v in plan.grad_views().items()})
    state = plan.state()
    launches = plan.launch_count()
    plan.keep_grads(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):
        plan.train_step(cd, target, mask, 5e-4, loss, step=steps + 1 + i)
    e0.record()
    for i in range(timing_steps):
        plan.train_step(cd, target, mask, 5e-4, loss, step=steps + 21 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / timing_steps
    return losses, grads, state, launches, ms


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    topk = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    t0 = time.time()
    a = run(True, n, topk, steps, 200)
    b = run(False, n, topk, steps, 200)
    print(f"rows {n} topk {topk}: split {a[4]:.4f} ms/step ({a[3]} launches)  fused {b[4]:.4f} ms/step ({b[3]} launches)"
          f"  [{time.time() - t0:.1f} s]")
    for s in range(steps):
        worst = max((rel(b[1][s][k].cpu().numpy(), a[1][s][k].cpu().numpy()), k) for k in a[1][s])
        print(f"  step {s + 1}: loss split {a[0][s]:.8f} fused {b[0][s]:.8f}  worst gradient rel diff {worst[0]:.3e} ({worst[1]})")
    worst = max((float((a[2][k] - b[2][k]).abs().max()), k) for k in a[2])
    print(f"  weights after {steps} steps: max |split - fused| = {worst[0]:.3e} ({worst[1]})")


if __name__ == "__main__":
    main()
