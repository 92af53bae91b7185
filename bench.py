#!/usr/bin/env python
"""Headline benchmark: coordinate samples/sec of one NPP-Net train step (encode + MLP forward + masked
MSE + backward + Adam) on B200, BASELINE.json's metric, on its configs[1] workload:

    NPP completion, 512x512 synthetic near-periodic texture, top-3 proposals, joint NPP_Net (K=3, D=8,
    W=512), N_rand 8192 pixel rows + 2 x 64^2 patch rows = 16384 coordinate rows per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (launched by torchrun, one rank per GPU) runs one independent fit per GPU -- the (image x proposal)
units shard with no data-path collective, so scaling is "weak" and `value` is the
job aggregate.  `--workload cfg4` instead splits one 2^18-row batch data-parallel with an NCCL all-reduce
of the gradient arena.

`--impl reference` times the reference's CPU path for the same step.  The reference is Python/PyTorch and
cannot travel to the GPU box, so this arm runs the numpy restatement in oracle/ (kind "port") on all host
cores.  bench.py and tests/ are the only places allowed to execute oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs (reference arm, cpu_baseline) must use all host
# cores, so the BLAS pools are sized before numpy is imported.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = {3: 21090816, 1: 15266304}      # SURVEY.md section 8a (fwd + wgrad + needed dgrad, x2)
GEMM_KMAJOR_MAC = {3: 3830528 - 768 + 2884352 - 768, 1: 2702080 - 768 + 2228992 - 768}  # fwd+dgrad without the 256x3 head
RES = (512, 512)
ROWS = 8192 + 2 * 64 * 64


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def kernel_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), averaged
    over its two launches per step (forward chain, dgrad chain).  None if no capture has been recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)["npp_gemm_kmajor"]
        return 0.5 * (t["forward_chain_bytes"] + t["dgrad_chain_bytes"])
    except Exception:
        return None


def synthetic_image(res=RES, seed=0):
    """Near-periodic RGB texture: lattice motif + illumination gradient + noise."""
    rng = np.random.default_rng(seed)
    H, W = res
    p = H / 12.0
    th = np.deg2rad(7.0)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    u = (xx * np.cos(th) + yy * np.sin(th)) / p
    v = (-xx * np.sin(th) + yy * np.cos(th)) / (0.9 * p)
    fu, fv = u - np.floor(u), v - np.floor(v)
    img = np.zeros((H, W, 3), np.float32)
    for _ in range(6):
        cu, cv, s = rng.random(), rng.random(), 0.05 + 0.1 * rng.random()
        col = rng.random(3).astype(np.float32)
        du = np.minimum(np.abs(fu - cu), 1 - np.abs(fu - cu))
        dv = np.minimum(np.abs(fv - cv), 1 - np.abs(fv - cv))
        img += np.exp(-(du * du + dv * dv) / (2 * s * s))[..., None] * col
    img *= (1.0 + 0.15 * (xx / W - 0.5))[..., None]
    img += rng.normal(0, 0.01, img.shape).astype(np.float32)
    return np.clip(img, 0, 1).astype(np.float32), p


def proposals(p, topk):
    angles = [[97.0, 187.0]] * 3
    periods = [[p, 0.9 * p], [p / 2, 0.45 * p], [2 * p, 1.8 * p]]
    return angles[:topk], periods[:topk]


def fourier_freqs():
    import torch
    g = torch.Generator().manual_seed(0)
    return (torch.randn(10, 1, generator=g) * 10).reshape(-1).numpy()          # embedder.py:26


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(topk, rows, steps, warmup, threads=None):
    """The reference's per-iteration CPU work restated in numpy (oracle/): gather rows of the precomputed
    encoding table (NPP_completion/train.py:178-181), forward, sigmoid + l2, backward, Adam."""
    from oracle import npp_oracle as O
    threads = threads or os.cpu_count()
    img, p = synthetic_image()
    angles, periods = proposals(p, topk)
    freqs = fourier_freqs()
    tabs = [O.encoder_tables(a, pr, [1], [0, -1, 1, 0.5, -0.5], [0]) for a, pr in zip(angles, periods)]
    rng = np.random.default_rng(0)
    pool = 4 * rows
    coords = np.stack([rng.integers(0, RES[0], pool), rng.integers(0, RES[1], pool)], 1).astype(np.float32)
    table = O.encode(coords, tabs, freqs, RES)                      # one-time table build, not timed
    target_all = img[coords[:, 0].astype(int), coords[:, 1].astype(int)]
    params = O.init_params(rng, topk=topk)
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(x) for k, x in params.items()}
    mask = np.ones((rows, 1), np.float32)
    t0 = None
    for it in range(1, warmup + steps + 1):
        if it == warmup + 1:
            t0 = time.perf_counter()
        sel = rng.choice(pool, rows, replace=False)
        O.train_step(params, m, v, it, table[sel], target_all[sel], mask, O.lr_schedule(it), topk_model=topk > 1)
    dt = time.perf_counter() - t0
    return rows * steps / dt, dt / steps * 1e3, threads


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg2-top1", "cfg4"])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    topk = 1 if args.workload == "cfg2-top1" else 3
    rows = args.rows or (1 << 18 if args.workload == "cfg4" else ROWS)
    metric = "coord samples/sec NPP-Net train step (fwd+bwd+Adam)"
    config = {"workload": {"cfg2": "cfg2: completion 512x512 synthetic near-periodic texture, top-3 proposals, joint NPP_Net "
                                   "K=3 D=8 W=512, 16384 coordinate rows/step (8192 pixel + 2x64^2 patch), loss l2; one independent fit per GPU",
                           "cfg2-top1": "cfg2 (one proposal per GPU): NPP_Net_top1 K=1 D=8 W=512, 16384 rows/step, loss l2",
                           "cfg4": "cfg4: remapping 2048x2048, K=3, 2^18-row batches split data-parallel, NCCL grad all-reduce"}[args.workload],
              "rows_per_step_per_gpu": rows if args.workload != "cfg4" else rows // max(world, 1),
              "l2_policy": "per-step working set (activations+deltas+split-K slabs, ~0.7 GB at 16384 rows) exceeds the 126 MB L2; "
                           "8 rotating coordinate batches",
              "input_pipeline": "GPU arm: the next batch's coordinates are encoded on the plan's low-priority side stream while "
                                "a step runs (npp_encode_prefetch, counted in gpu_launches); steps run on a priority -1 stream"}

    if args.impl == "reference":
        if rank != 0:
            return
        ref_rows = min(rows, 16384)
        steps = max(1, min(args.steps, 20))
        warm = max(1, min(args.warmup, 3))
        val, ms, threads = cpu_reference(topk, ref_rows, steps, warm)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port",
                                 "sample": f"{steps} steps x {ref_rows} rows of the same workload, numpy oracle (oracle/npp_oracle.py), "
                                           "table gather + fwd + l2 + bwd + Adam"},
                "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan

    torch.cuda.set_device(local_rank)
    # The train steps run on a high-priority stream: the plan's input-prefetch stream (default = lowest priority) then
    # only gets the SMs the step's kernels leave idle (the chain kernels occupy 128 of the 148).
    torch.cuda.set_stream(torch.cuda.Stream(device=local_rank, priority=-1))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    img_np, p = synthetic_image(seed=rank if args.workload != "cfg4" else 0)
    angles, periods = proposals(p, topk)
    enc = EncoderSpec.from_proposals(RES, angles, periods, fourier_freqs())
    dp = args.workload == "cfg4" and world > 1
    my_rows = rows // world if dp else rows
    plan = Plan(enc, max_rows=my_rows)
    plan.reset_parameters(seed=0)        # nn.Linear default init, like the reference (helpers.py:140-141)
    img = torch.from_numpy(img_np).to(dev)

    NB = 8
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_coords = [torch.stack([torch.randint(0, RES[0], (my_rows,), generator=g),
                                torch.randint(0, RES[1], (my_rows,), generator=g)], 1).float().pin_memory() for _ in range(NB)]
    host_target = [img_np[c[:, 0].long().numpy(), c[:, 1].long().numpy()] for c in host_coords]
    host_target = [torch.from_numpy(t).pin_memory() for t in host_target]
    host_mask = torch.ones(my_rows, 1).pin_memory()
    dev_coords = [c.to(dev) for c in host_coords]
    dev_target = [t.to(dev) for t in host_target]
    dev_mask = host_mask.to(dev)
    loss_d = torch.zeros((), device=dev)
    lr = 5e-4
    n_norm = rows if dp else my_rows

    def step_resident(i):
        if dp:
            b = i % NB
            logits = plan.forward(dev_coords[b])
            _, gl, _ = plan.mse(logits, dev_target[b], dev_mask, n_norm=n_norm)
            plan.backward(my_rows, gl)
            dist.all_reduce(plan.grads[: plan.trained_floats])
            plan.adam_step(lr)
        else:
            b = i % NB
            # input pipelining: the next batch is encoded on the plan's side stream while this step runs
            plan.prefetch_encode(dev_coords[(i + 1) % NB])
            plan.train_step(dev_coords[b], dev_target[b], dev_mask, lr, loss_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi needs ~1 s to come up
    # W warm-up steps, and keep going until the SM clock has had ~0.5 s of load to settle
    t_w = time.time()
    i = 0
    while i < args.warmup or time.time() - t_w < 0.5:
        step_resident(i)
        i += 1
        if i % 50 == 0:
            torch.cuda.synchronize()
    launches_per_step = plan.launch_count() if not dp else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = ms.item()
    total_rows = (rows if dp else rows * world) * args.steps
    value = total_rows / (total_ms * 1e-3)
    final_loss = loss_d.item()

    # ---- per-kernel-class timing (CUDA events on the launch stream) over another K steps
    prof = None
    if not dp:
        plan.profile(True)
        for i in range(args.steps):
            step_resident(i)
        prof = plan.profile_read()
        plan.profile(False)

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs, D2H of the loss
    # Inputs go through two device slots filled by a copy stream (the usual double-buffered loader): the H2D copies of
    # step i+1 are issued before the host blocks on the loss of step i, every copy and every loss read stays inside
    # the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    slots = [(torch.empty_like(dev_coords[0]), torch.empty_like(dev_target[0]), torch.empty_like(dev_mask))
             for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for ev in ev_free:
        ev.record(main_stream)

    def issue_copies(i):
        b, k = i % NB, i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])          # the step that last used this slot has finished
            slots[k][0].copy_(host_coords[b], non_blocking=True)
            slots[k][1].copy_(host_target[b], non_blocking=True)
            slots[k][2].copy_(host_mask, non_blocking=True)
            if not dp:
                plan.prefetch_encode(slots[k][0])       # encoded as soon as the coordinates have landed
            ev_ready[k].record(copy_stream)

    def step_e2e(i):
        k = i % 2
        main_stream.wait_event(ev_ready[k])
        c, t, mk = slots[k]
        if dp:
            logits = plan.forward(c)
            l, gl, _ = plan.mse(logits, t, mk, n_norm=n_norm)
            plan.backward(my_rows, gl)
            dist.all_reduce(plan.grads[: plan.trained_floats])
            plan.adam_step(lr)
            ev_free[k].record(main_stream)
            issue_copies(i + 1)
            return l.item()
        plan.train_step(c, t, mk, lr, loss_d)
        ev_free[k].record(main_stream)
        issue_copies(i + 1)                             # next step's inputs: copies + encoding overlap this step
        return loss_d.item()

    issue_copies(0)
    for i in range(3):
        step_e2e(i)
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t2 = time.time()
    e0.record()
    for i in range(3, 3 + args.steps):   # the pipeline keeps running: each timed step issues the next step's copies
        step_e2e(i)
    e1.record()
    barrier()
    # the end-to-end loop blocks on the loss every step: the short idle gaps let a power-capped GPU clock higher
    # inside the kernels than the back-to-back loop above does, which is why E can exceed `value`
    clocks_e2e = sampler2.stop(t2, time.time()) if sampler2 else None
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = total_rows / (ms2.item() * 1e-3)
    h2d = my_rows * (2 + 3 + 1) * 4
    d2h = 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    sustained = total_ms > 2000.0      # timed region longer than 2 s -> compare with the sustained peak
    peak = pk["bf16_tflops_sustained"] if sustained else pk["bf16_tflops"]
    roofline = None
    if prof is not None:
        gemm_ms = prof["gemm_fwd"][0] + prof["gemm_dgrad"][0]
        gemm_launches = prof["gemm_fwd"][1] + prof["gemm_dgrad"][1]
        flops = 2.0 * GEMM_KMAJOR_MAC[topk] * rows * args.steps            # algorithmic FLOPs of those launches
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        step_ms_prof = sum(v[0] for v in prof.values()) / args.steps
        roofline = {"bound": "tensor", "kernel": "npp_gemm_kmajor (forward + dgrad GEMMs, tcgen05 kind::f16)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": f"{pk_kind} bf16 dense, {'sustained' if sustained else 'burst'} (fp16 and bf16 share the kind::f16 pipe)",
                    "traffic": kernel_traffic(),
                    "launches_per_step": gemm_launches / args.steps,
                    "avg_launch_us": gemm_ms * 1e3 / gemm_launches,
                    "flop_per_launch_avg": flops / gemm_launches,
                    "class_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                    "class_share_of_step": {k: v[0] / args.steps / step_ms_prof for k, v in prof.items()},
                    "whole_step": {"achieved": value / world * FLOP_PER_SAMPLE[topk] / 1e12,
                                   "frac": value / world * FLOP_PER_SAMPLE[topk] / 1e12 / peak,
                                   "flop_per_sample": FLOP_PER_SAMPLE[topk]}}
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if dp else "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (f32 master weights, loss, Adam)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "sm_mhz": (clocks_e2e or {}).get("sm_mhz")},
            "gpu_launches": (launches_per_step or 0) * args.steps, "final_loss": final_loss, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        cval, cms, cthreads = cpu_reference(topk, 8192, 6, 2)
        line["cpu_baseline"] = {"value": cval, "unit": "samples/s", "cores": cthreads, "kind": "port",
                                "sample": "6 steps x 8192 rows of the same workload on the host cores: numpy oracle, "
                                          "table gather + fwd + l2 + bwd + Adam"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
