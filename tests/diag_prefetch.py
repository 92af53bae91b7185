"""GPU diagnostic (not a pytest): run the same 7 train steps with and without npp_encode_prefetch, several times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_parity_gpu as T

n = 3000
rng = np.random.default_rng(11)
batches = [torch.from_numpy(np.stack([rng.integers(0, T.RES[0], n), rng.integers(0, T.RES[1], n)], 1).astype(np.float32)).cuda() for _ in range(4)]
targets = [torch.rand(n, 3, device="cuda") for _ in range(4)]
mask = torch.ones(n, 1, device="cuda")

def run(prefetch, extra):
    plan, *_ = T.make(3, n)
    loss_d = torch.zeros((), device="cuda")
    losses = []
    for step in range(1, 8):
        b = (step - 1) % 4
        if prefetch:
            plan.prefetch_encode(batches[step % 4])
            if extra and step == 4:
                plan.prefetch_encode(batches[(step + 2) % 4])
        plan.train_step(batches[b], targets[b], mask, 5e-4, loss_d, step=step)
        losses.append(loss_d.item())
    return np.array(losses), {k: v.clone() for k, v in plan.state().items()}

for trial in range(8):
    mode = [(False, False), (True, False), (True, True)][trial % 3]
    l0, s0 = run(False, False)
    l1, s1 = run(*mode)
    dl = np.abs(l0 - l1) / np.abs(l0)
    dw = max((s0[k] - s1[k]).abs().max().item() for k in s0)
    print(f"trial {trial} mode {mode}: max rel loss diff {dl.max():.3e} at step {dl.argmax()+1}; max weight diff {dw:.3e}", flush=True)
