"""Timing of search-stage fits (NPP_Net_light, 2048 rows/step; NPP_proposal/search.py:85-148):
one fit alone, and 9 candidate fits (the reference ranks up to 9 candidates one after the other) interleaved on 9
streams with one plan each.  Not a test; run on the GPU box: python tests/diag_search_fit.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npp_b200  # noqa: E402,F401
from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT  # noqa: E402

RES = (512, 512)
N, ITERS = 2048, 300


def make(seed, period):
    rng = np.random.default_rng(seed)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, [[83.0, 172.5]], [[period, 0.9 * period]], freqs, include_input=False)
    plan = Plan(enc, depth=4, width=256, skip_layer=-1, max_rows=N, model=MODEL_LIGHT)
    plan.reset_parameters(seed)
    return plan


def batches(k):
    g = torch.Generator(device="cuda").manual_seed(0)
    return [(torch.stack([torch.randint(0, RES[0], (N,), device="cuda", generator=g),
                          torch.randint(0, RES[1], (N,), device="cuda", generator=g)], 1).float().contiguous(),
             torch.rand(N, 3, device="cuda", generator=g)) for _ in range(k)]


def main():
    data = batches(8)
    plan = make(0, 42.7)
    loss = torch.zeros((), device="cuda")
    for i in range(20):
        plan.train_step(*data[i % 8], None, 5e-4, loss)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(ITERS):
        plan.train_step(*data[i % 8], None, 5e-4, loss)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / ITERS
    print(f"one search fit: {ms * 1e3:.1f} us/step  {N / ms / 1e3:.2f} M samples/s  "
          f"({ITERS} iterations = {ms * ITERS:.1f} ms; launches/step {plan.launch_count()})")
    plan.profile(True)
    for i in range(50):
        plan.train_step(*data[i % 8], None, 5e-4, loss)
    prof = plan.profile_read()
    plan.profile(False)
    print("  per class us/step:", {k: round(v[0] / 50 * 1e3, 1) for k, v in prof.items() if v[1]})

    # 9 candidates: one plan and one stream each, iterations issued round-robin
    K = 9
    plans = [make(s, 30.0 + 3 * s) for s in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    losses = [torch.zeros((), device="cuda") for _ in range(K)]
    torch.cuda.synchronize()
    for it in range(10):
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                plans[k].train_step(*data[(it + k) % 8], None, 5e-4, losses[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for it in range(ITERS):
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                plans[k].train_step(*data[(it + k) % 8], None, 5e-4, losses[k])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = a.elapsed_time(b)
    print(f"{K} candidate fits on {K} streams: {ms:.1f} ms for {ITERS} iterations each (wall {wall:.1f} ms) = "
          f"{ms / ITERS / K * 1e3:.1f} us per fit-step, {K * ITERS * N / ms / 1e3:.2f} M samples/s")

    print("losses:", [round(l.item(), 5) for l in losses])

    # the same nine fits through search_fits.run_fits: one npp_fit_run call per candidate, one host thread each
    from npp_b200.search_fits import run_fits
    coords_all = torch.stack([d[0] for d in data] * ((ITERS + 7) // 8))[:ITERS].contiguous()
    target_all = torch.stack([d[1] for d in data] * ((ITERS + 7) // 8))[:ITERS].contiguous()
    for threads in (False, True):
        plans2 = [make(s, 30.0 + 3 * s) for s in range(K)]
        run_fits(plans2, coords_all[:10], target_all[:10], streams=streams, threads=threads)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        out = run_fits(plans2, coords_all, target_all, streams=streams, threads=threads)
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = a.elapsed_time(b)
        print(f"run_fits, {K} candidates x {ITERS} iterations, host threads={threads}: {ms:.1f} ms (wall {wall:.1f} ms) = "
              f"{ms / ITERS / K * 1e3:.1f} us per fit-step, {K * ITERS * N / ms / 1e3:.2f} M samples/s; "
              f"final losses {[round(v, 4) for v in out[:, -1].tolist()][:3]}...")
    one = make(0, 42.7)
    one.fit_run(coords_all[:10], target_all[:10])
    torch.cuda.synchronize()
    a.record()
    one.fit_run(coords_all, target_all)
    b.record()
    torch.cuda.synchronize()
    print(f"fit_run, one candidate: {a.elapsed_time(b):.1f} ms for {ITERS} iterations = {a.elapsed_time(b) / ITERS * 1e3:.1f} us/step")

    # the numpy port of the same step on the host cores (oracle/npp_oracle.py train_step_light; a bounded sample)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import npp_oracle as O
    rng = np.random.default_rng(0)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    table = O.encoder_tables([83.0, 172.5], [42.7, 38.4], [1], [0, -1, 1, 0.5, -0.5], [0])
    p_ = O.init_params_light(rng)
    m_ = {k: np.zeros_like(v) for k, v in p_.items()}
    v_ = {k: np.zeros_like(v) for k, v in p_.items()}
    cpu_coords = data[0][0].cpu().numpy()
    cpu_target = data[0][1].cpu().numpy()
    t0 = time.perf_counter()
    steps = 20
    for it in range(1, steps + 1):
        pos, per = O.encode_search(cpu_coords, table, freqs, RES)
        O.train_step_light(p_, m_, v_, it, pos, per, cpu_target, 5e-4)
    dt = (time.perf_counter() - t0) / steps
    print(f"numpy port on {os.cpu_count()} host cores (encode + step): {dt * 1e3:.2f} ms/step  {N / dt / 1e6:.3f} M samples/s")

    # the same layer stack in plain torch (eager fp32 and TF32) on this GPU: what the reference's NPP_Net_light costs
    # per iteration of NPP_proposal/search.py:112-146 once its encodings are precomputed
    import torch.nn as nn

    class Light(nn.Module):
        def __init__(self, W=256, D=4):
            super().__init__()
            self.trunk = nn.ModuleList([nn.Linear(20, W)] + [nn.Linear(W, W) for _ in range(D - 1)])
            self.f1, self.pos, self.rgb = nn.Linear(W, W), nn.Linear(W + 42, W // 2), nn.Linear(W // 2, 3)

        def forward(self, x, xp):
            h = xp
            for l in self.trunk:
                h = l(h)
                h = h + torch.sin(h) ** 2
            h = self.pos(torch.cat([self.f1(h), x], -1))
            return self.rgb(h + torch.sin(h) ** 2)

    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        net = Light().cuda()
        opt = torch.optim.Adam(net.parameters(), lr=5e-4)
        x, xp, y = torch.randn(N, 42, device="cuda"), torch.randn(N, 20, device="cuda"), torch.rand(N, 3, device="cuda")

        def step():
            pred = torch.sigmoid(net(x, xp))
            opt.zero_grad()
            loss = torch.mean((pred - y) ** 2)
            loss.backward()
            opt.step()
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        a.record()
        for _ in range(100):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 100
        print(f"torch eager {'tf32' if tf32 else 'fp32'} (same layer stack, encodings precomputed): {ms * 1e3:.0f} us/step  "
              f"{N / ms / 1e3:.2f} M samples/s")


if __name__ == "__main__":
    main()
