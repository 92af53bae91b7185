"""GPU diagnostic (not a pytest): throughput of the drop-in ``models`` surface as the unchanged reference scripts use it
(render -> img2mse -> loss.backward() -> optimizer.step(), NPP_completion/train.py:187-263) at 16 384 rows."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import numpy as np, torch
import models.helpers as H
from models.mse_calculator import img2mse

args = argparse.Namespace(multires=10, i_embed=0, p_topk=3, freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5],
                          angle_offsets=[0], netdepth=8, netwidth=512, activation='snake', netchunk=1024 * 4096,
                          lrate=5e-4, lrate_decay=500, normalize_type=1, loss_type='l2')
res = (512, 512)
angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
periods = torch.Tensor([[17.2, 14.9], [8.6, 7.45], [34.4, 29.8]])
kw, _, _, grad_vars, optimizer, embedder, per = H.create_npp_net(args, angles, periods, res, None)
n = 16384
coords = torch.stack([torch.randint(0, res[0], (n,)), torch.randint(0, res[1], (n,))], 1).float().cuda()
emb = torch.cat([embedder.embed(e.embed(coords.clone())) for e in per], 1)
gt = torch.rand(n, 3, device="cuda"); mask = torch.ones(n, 1, device="cuda")
for loss_type in ("l2", "robust_loss_adaptive"):
    args.loss_type = loss_type
    def step():
        pred = H.render(None, emb, args, **kw)
        optimizer.zero_grad()
        loss = img2mse(pred, gt, loss_type, H.adaptive_pix, mask)
        loss.backward()
        optimizer.step()
        return loss
    for _ in range(20): step()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(200): step()
    torch.cuda.synchronize(); dt = (time.time() - t0) / 200
    print(f"autograd path, {loss_type:22s}: {dt*1e3:.3f} ms/step  {n/dt/1e6:.2f} M samples/s", flush=True)

# host-side cost per phase (no synchronisation inside the loop: this is pure CPU time spent enqueueing)
args.loss_type = "l2"
import collections
acc = collections.OrderedDict((k, 0.0) for k in ("render", "zero_grad", "img2mse", "backward", "optimizer.step"))
for _ in range(200):
    t = time.perf_counter(); pred = H.render(None, emb, args, **kw); acc["render"] += time.perf_counter() - t
    t = time.perf_counter(); optimizer.zero_grad(); acc["zero_grad"] += time.perf_counter() - t
    t = time.perf_counter(); loss = img2mse(pred, gt, "l2", H.adaptive_pix, mask); acc["img2mse"] += time.perf_counter() - t
    t = time.perf_counter(); loss.backward(); acc["backward"] += time.perf_counter() - t
    t = time.perf_counter(); optimizer.step(); acc["optimizer.step"] += time.perf_counter() - t
torch.cuda.synchronize()
print("host us per step:", {k: round(v / 200 * 1e6, 1) for k, v in acc.items()}, flush=True)
