"""GPU diagnostic (not a pytest): run-to-run bit reproducibility of the chain kernels (forward logits, weight
gradients) over many repetitions and row counts -- a stress test for the forwarding / cluster synchronisation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_parity_gpu as T

reps = int(os.environ.get("REPS", "300"))
for topk, n in ((3, 16384), (3, 5000), (1, 16384), (3, 40000)):
    plan, params, coords, tabs, freqs, rng = T.make(topk, n)
    cd = torch.from_numpy(coords).cuda()
    g = torch.randn(n, 3, device="cuda") * 1e-4
    ref_logits = plan.forward(cd).clone()
    plan.backward(n, g)
    ref_grads = {k: v.clone() for k, v in plan.grad_views().items() if k.endswith(".weight") and not k.startswith("rgb_linear")}
    bad_f = bad_b = 0
    for i in range(reps):
        logits = plan.forward(cd)
        if not torch.equal(logits, ref_logits):
            bad_f += 1
        plan.backward(n, g)
        gv = plan.grad_views()
        if any(not torch.equal(gv[k], ref_grads[k]) for k in ref_grads):
            bad_b += 1
    print(f"topk={topk} n={n}: {reps} repetitions, forward mismatches {bad_f}, weight-gradient mismatches {bad_b}", flush=True)
