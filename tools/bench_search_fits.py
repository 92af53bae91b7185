"""Supplementary bench line for the search-stage fits; bench.py keeps BASELINE.json's metric.

One "step" is a whole candidate search: K candidate periodicities x ITERS iterations of N_rand pixels each
(NPP_proposal/search.py:85-148 with the defaults of options/arg_config.py:114-143: NPP_Net_light D=4 W=256, 2048
pixels, 300 iterations, up to 9 candidates), fitted side by side by search_fits.run_fits.  Prints ONE JSON line in
bench.py's format:
  value     coordinate samples/s with the batches resident in HBM,
  e2e       the same with the pixel indices drawn on the host and coordinates + targets copied from pinned host
            memory inside the timed region, and the final losses read back,
  roofline  algorithmic FLOPs (1.83 MFLOP per training sample) against the measured dense peak: the fits are latency
            bound, the fraction says how far from tensor-bound they are,
  cpu_baseline  the numpy port of one step (oracle/npp_oracle.py) on the host cores, a bounded sample.

    python tools/bench_search_fits.py [--candidates 9] [--iters 300] [--steps 5] [--warmup 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (peaks, ClockSampler, synthetic_image)
import npp_b200  # noqa: E402,F401
from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT  # noqa: E402
from npp_b200.search_fits import run_fits, gather_batches  # noqa: E402

RES = (512, 512)
N_RAND = 2048
FLOP_PER_SAMPLE = 3 * 2 * (20 * 256 + 4 * 256 * 256 + 298 * 128 + 128 * 3)   # fwd + dgrad + wgrad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=9)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="grouped", choices=["grouped", "threads"],
                    help="grouped: one re-launched CUDA graph with a branch per candidate (npp_multi_fit_run); threads: one "
                         "stream + host thread per candidate (npp_fit_run)")
    args = ap.parse_args()
    K, iters = args.candidates, args.iters
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    img_np, p = B.synthetic_image(seed=0)
    image = torch.from_numpy(img_np).to(dev)
    H, W = RES
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    train_coords = torch.from_numpy(np.stack([yy.reshape(-1), xx.reshape(-1)], 1)).to(dev)
    rng = np.random.default_rng(0)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)

    def plans():
        out = []
        for k in range(K):          # candidate k: the true period scaled by (1 + k/8), like the shift hypotheses of the search
            per = [p * (1 + k / 8), 0.9 * p * (1 + k / 8)]
            enc = EncoderSpec.from_proposals(RES, [[97.0, 187.0]], [per], freqs, include_input=False)
            pl = Plan(enc, depth=4, width=256, skip_layer=-1, max_rows=N_RAND, model=MODEL_LIGHT)
            pl.reset_parameters(seed=0)
            out.append(pl)
        return out

    def host_indices(seed):
        r = np.random.default_rng(seed)
        return torch.from_numpy(r.integers(0, H * W, (iters, N_RAND)))

    streams = [torch.cuda.Stream(device=dev) for _ in range(K)] if args.mode == "threads" else None
    idx = host_indices(0).to(dev)
    coords_all, target_all = gather_batches(image, train_coords, idx)
    fits = plans()
    for _ in range(args.warmup):
        run_fits(fits, coords_all, target_all, streams=streams)
    torch.cuda.synchronize()
    sampler = B.ClockSampler(0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    a.record()
    for _ in range(args.steps):
        losses = run_fits(fits, coords_all, target_all, streams=streams)
    b.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = a.elapsed_time(b) / args.steps
    clocks = sampler.stop(t0, t1)
    samples = K * iters * N_RAND
    value = samples / (ms * 1e-3)

    # end to end: indices drawn on the host, coordinates and targets gathered on the host image and copied from pinned
    # memory, final losses read back -- every search
    img_host = torch.from_numpy(img_np)
    tc_host = train_coords.cpu()
    pin_c = torch.empty(iters, N_RAND, 2, dtype=torch.float32).pin_memory()
    pin_t = torch.empty(iters, N_RAND, 3, dtype=torch.float32).pin_memory()
    dev_c, dev_t = torch.empty_like(pin_c, device=dev), torch.empty_like(pin_t, device=dev)

    def search_e2e(seed):
        sel = tc_host[host_indices(seed)]
        pin_c.copy_(sel.float())
        pin_t.copy_(img_host[sel[..., 0], sel[..., 1], :])
        dev_c.copy_(pin_c, non_blocking=True)
        dev_t.copy_(pin_t, non_blocking=True)
        ls = run_fits(fits, dev_c, dev_t, streams=streams)
        return ls[:, -1].cpu()

    for s in range(2):
        search_e2e(100 + s)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for s in range(args.steps):
        final = search_e2e(s)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps

    # CPU: the numpy port of one search iteration on the host cores
    from oracle import npp_oracle as O
    table = O.encoder_tables([97.0, 187.0], [p, 0.9 * p], [1], [0, -1, 1, 0.5, -0.5], [0])
    pp = O.init_params_light(np.random.default_rng(0))
    mm = {k: np.zeros_like(v) for k, v in pp.items()}
    vv = {k: np.zeros_like(v) for k, v in pp.items()}
    cc, tt = coords_all[0].cpu().numpy(), target_all[0].cpu().numpy()
    cpu_steps = 40
    c0 = time.perf_counter()
    for it in range(1, cpu_steps + 1):
        pos, per = O.encode_search(cc, table, freqs, RES)
        O.train_step_light(pp, mm, vv, it, pos, per, tt, 5e-4)
    cpu_dt = (time.perf_counter() - c0) / cpu_steps

    pk, src = B.peaks()
    peak = float(pk.get("bf16_tflops", 1607.5))
    achieved = value * FLOP_PER_SAMPLE / 1e12
    line = {
        "metric": "coord samples/sec NPP_Net_light search fits (fwd+bwd+Adam), candidates fitted side by side",
        "value": value, "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (f32 master weights, loss, Adam)", "data": "synthetic",
        "config": {"workload": f"search: 512x512 synthetic near-periodic texture, {K} candidate periodicities x {iters} "
                               f"iterations x {N_RAND} pixels, NPP_Net_light D=4 W=256, loss l2; one step = one whole search",
                   "l2_policy": "every iteration reads a different batch; working set per fit 4 MB (L2 resident by design)",
                   "concurrency": ("one re-launched CUDA graph with one branch per candidate (npp_multi_fit_run)" if args.mode == "grouped" else "one plan, CUDA stream and host thread per candidate (npp_fit_run)")},
        "clocks": clocks,
        "e2e": {"value": samples / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_search": e2e_ms,
                "h2d_bytes_per_step": int(pin_c.numel() * 4 + pin_t.numel() * 4), "d2h_bytes_per_step": int(K * 4),
                "note": "includes drawing the pixel indices and gathering coordinates / targets on the host"},
        "gpu_launches": int(args.steps * sum(pl.launch_count() for pl in fits)),
        "final_losses": [round(float(v), 5) for v in final.tolist()],
        "roofline": {"bound": "tensor", "kernel": "npp_gemm_kmajor (forward + dgrad chains of NPP_Net_light)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": f"{src} bf16 dense (burst)", "traffic": None,
                     "flop_per_sample": FLOP_PER_SAMPLE,
                     "note": "latency bound: six dependent 256x256x256 GEMMs per chain on 8 CTA pairs per fit"},
        "cpu_baseline": {"value": N_RAND / cpu_dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{cpu_steps} iterations x {N_RAND} rows of one candidate: numpy oracle, encode + fwd + l2 + bwd + Adam"},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
