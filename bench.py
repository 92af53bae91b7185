#!/usr/bin/env python
"""Headline benchmark: coordinate samples/sec of one NPP-Net train step (encode + MLP forward + masked
MSE + backward + Adam) on B200, BASELINE.json's metric, on its configs[1] workload:

    NPP completion, 512x512 synthetic near-periodic texture, top-3 proposals, joint NPP_Net (K=3, D=8,
    W=512), N_rand 8192 pixel rows + 2 x 64^2 patch rows = 16384 coordinate rows per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|...]

N > 1 (launched by torchrun, one rank per GPU) runs one independent fit per GPU -- the (image x proposal)
units shard with no data-path collective, so scaling is "weak" and `value` is the
job aggregate.  The same JSON line carries short extra records under "also": the other BASELINE configs
(cfg2-top1, cfg3, cfg5 in fits/s), at N > 1 the one path that HAS a collective (cfg4: a 2^18-row batch split
data-parallel with an overlapped NCCL all-reduce of the gradients, strong scaling), and at N = 1 the
reference's own modules run eagerly on the same B200 ("reference_gpu").  `--workload X` makes X the headline.

`--impl reference` times the reference's CPU path for the same step on the host cores: the UNMODIFIED
reference modules from baseline/_ref (a git-ignored copy made by __graft_entry__.build(), which travels to
the GPU box; kind "reference"), or, when that copy is missing, the numpy restatement in oracle/ (kind "port").
bench.py and tests/ are the only places allowed to execute oracle/.
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
METRIC = "coord samples/sec NPP-Net train step (fwd+bwd+Adam)"

WORKLOADS = {
    "cfg2": dict(res=(512, 512), topk=3, rows=ROWS, mask="ones",
                 desc="cfg2: completion 512x512 synthetic near-periodic texture, top-3 proposals, joint NPP_Net K=3 D=8 "
                      "W=512, 16384 coordinate rows/step (8192 pixel + 2x64^2 patch), loss l2; one independent fit per GPU"),
    "cfg2-top1": dict(res=(512, 512), topk=1, rows=ROWS, mask="ones",
                      desc="cfg2 (one proposal per GPU): NPP_Net_top1 K=1 D=8 W=512, 16384 rows/step, loss l2"),
    "cfg3": dict(res=(1024, 1024), topk=3, rows=8192 + 2 * 96 * 96, mask="binary",
                 desc="cfg3: segmentation 1024x1024 synthetic tiled facade, K=3, 26624 rows/step (8192 pixel + 2x96^2 "
                      "patch), masked l2 (0/1 mask, 20 % occluder)"),
    "cfg4": dict(res=(2048, 2048), topk=3, rows=1 << 18, mask="ones",
                 desc="cfg4: remapping 2048x2048, K=3, 2^18-row batches split data-parallel over the ranks, NCCL "
                      "all-reduce of the gradients overlapped with the weight-gradient GEMMs"),
    "cfg5": dict(res=(512, 512), topk=1, rows=ROWS, mask="ones",
                 desc="cfg5: 64 synthetic 512x512 images x top-3 proposals = 192 independent K=1 fits dealt round-robin "
                      "to the ranks, 16384 rows/step"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def kernel_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json).  None if
    no capture has been recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)["npp_gemm_kmajor"]
        if "fused_chain_bytes" in t:
            return t["fused_chain_bytes"]
        return t["forward_chain_bytes"] + t["dgrad_chain_bytes"]
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


def occluder_mask(res, seed=0):
    """0/1 loss mask of cfg3: 1 on the periodic region, 0 on ~20 % of the image covered by rectangular occluders."""
    rng = np.random.default_rng(seed + 77)
    H, W = res
    m = np.ones((H, W), np.float32)
    while m.mean() > 0.8:
        h, w = int(H * (0.08 + 0.12 * rng.random())), int(W * (0.08 + 0.12 * rng.random()))
        y, x = rng.integers(0, H - h), rng.integers(0, W - w)
        m[y:y + h, x:x + w] = 0.0
    return m


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


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_port(topk, rows, steps, warmup, res=RES):
    """The reference's per-iteration CPU work restated in numpy (oracle/): gather rows of the precomputed
    encoding table (NPP_completion/train.py:178-181), forward, sigmoid + l2, backward, Adam."""
    from oracle import npp_oracle as O
    img, p = synthetic_image(res)
    angles, periods = proposals(p, topk)
    freqs = fourier_freqs()
    tabs = [O.encoder_tables(a, pr, [1], [0, -1, 1, 0.5, -0.5], [0]) for a, pr in zip(angles, periods)]
    rng = np.random.default_rng(0)
    pool = 4 * rows
    coords = np.stack([rng.integers(0, res[0], pool), rng.integers(0, res[1], pool)], 1).astype(np.float32)
    table = O.encode(coords, tabs, freqs, res)                      # one-time table build, not timed
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
    return rows * steps / dt, dt / steps * 1e3


def reference_fit(topk, rows, device="cpu", res=RES, anomaly=True):
    """A ReferenceFit (the reference's own modules, baseline/reference_arm.py) on the synthetic workload, or None
    when no reference checkout is available on this box."""
    from baseline import reference_arm as R
    if R.reference_root() is None:
        return None
    img, p = synthetic_image(res)
    angles, periods = proposals(p, topk)
    rng = np.random.default_rng(0)
    pool = 4 * rows
    coords = np.stack([rng.integers(0, res[0], pool), rng.integers(0, res[1], pool)], 1).astype(np.float32)
    return R.ReferenceFit(res, angles, periods, img, coords, topk=topk, device=device, anomaly=anomaly)


def cpu_reference(topk, rows, steps, warmup, res=RES):
    """(samples/s, ms/step, threads, kind, description) of the reference's CPU train step on all host cores."""
    threads = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(threads)
        fit = reference_fit(topk, rows, "cpu", res, anomaly=True)
    except Exception as e:                     # the copy is missing or does not import on this box: the numpy port
        print(f"[bench] reference modules unavailable ({type(e).__name__}: {e}); using the numpy port", file=sys.stderr)
        fit = None
    if fit is not None:
        from baseline import reference_arm as R
        val, ms, _ = R.time_fit(fit, rows, steps, warmup)
        return val, ms, threads, "reference", (
            f"{steps} steps x {rows} rows of the same workload: UNMODIFIED reference modules (models/networks.py, "
            f"helpers.render, img2mse 'l2', torch.optim.Adam; torch {torch.__version__} fp32 eager, anomaly detection on "
            f"as shipped), table gather + fwd + loss + bwd + Adam; one-time table build {fit.table_build_s:.1f} s not timed")
    val, ms = cpu_port(topk, rows, steps, warmup, res)
    return val, ms, threads, "port", (f"{steps} steps x {rows} rows of the same workload: numpy oracle "
                                      "(oracle/npp_oracle.py), table gather + fwd + l2 + bwd + Adam")


def reference_on_gpu(topk, rows, dev_index, steps=20, warmup=3):
    """The reference's own modules, eager torch on this B200 ('beat this on the same box')."""
    import torch
    from baseline import reference_arm as R
    out = {}
    dev = f"cuda:{dev_index}"
    variants = (("fp32_anomaly_on_as_shipped", True, False, None), ("fp32", False, False, None),
                ("tf32", False, True, None), ("bf16_autocast", False, True, torch.bfloat16))
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, anomaly, tf32, autocast in variants:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            fit = reference_fit(topk, rows, dev, anomaly=anomaly)
            if fit is None:
                return None
            val, ms, loss = R.time_fit(fit, rows, steps, warmup, autocast=autocast)
            out[name] = {"value": val, "ms_per_step": ms}
            del fit
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev[0]
        torch.autograd.set_detect_anomaly(False)
    out["unit"] = "samples/s"
    out["note"] = (f"reference models/networks.py + helpers.render + img2mse + torch.optim.Adam on cuda:{dev_index}, "
                   f"{steps} steps x {rows} rows, K={topk}, encoding table resident (its build is not timed)")
    return out


# ------------------------------------------------------------------------------------------ GPU arm
class Fit:
    """One plan + resident synthetic batches of a workload on the current device."""

    def __init__(self, name, rank, dev, my_rows=None, nb=8, image_seed=None):
        import torch
        from npp_b200.plan import EncoderSpec, Plan
        w = WORKLOADS[name]
        self.name, self.w, self.dev, self.torch = name, w, dev, torch
        self.res, self.topk = w["res"], w["topk"]
        self.rows = my_rows or w["rows"]
        img_np, p = synthetic_image(self.res, seed=rank if image_seed is None else image_seed)
        angles, periods = proposals(p, self.topk)
        self.plan = Plan(EncoderSpec.from_proposals(self.res, angles, periods, fourier_freqs()), max_rows=self.rows)
        self.plan.reset_parameters(seed=0)        # nn.Linear default init, like the reference (helpers.py:140-141)
        g = torch.Generator(device="cpu").manual_seed(1234 + rank)
        self.host_coords = [torch.stack([torch.randint(0, self.res[0], (self.rows,), generator=g),
                                         torch.randint(0, self.res[1], (self.rows,), generator=g)], 1).float().pin_memory()
                            for _ in range(nb)]
        self.host_target = [torch.from_numpy(img_np[c[:, 0].long().numpy(), c[:, 1].long().numpy()]).pin_memory()
                            for c in self.host_coords]
        if w["mask"] == "binary":
            m = occluder_mask(self.res)
            self.host_mask = [torch.from_numpy(m[c[:, 0].long().numpy(), c[:, 1].long().numpy()][:, None].copy()).pin_memory()
                              for c in self.host_coords]
        else:
            self.host_mask = [torch.ones(self.rows, 1).pin_memory()] * nb
        self.coords = [c.to(dev) for c in self.host_coords]
        self.target = [t.to(dev) for t in self.host_target]
        self.mask = [m.to(dev) for m in self.host_mask]
        self.loss = torch.zeros((), device=dev)
        self.nb = nb
        self.lr = 5e-4

    def step(self, i):
        b = i % self.nb
        # input pipelining: the next batch is encoded on the plan's side stream while this step runs
        self.plan.prefetch_encode(self.coords[(i + 1) % self.nb])
        self.plan.train_step(self.coords[b], self.target[b], self.mask[b], self.lr, self.loss)


def timed_loop(step, steps, warmup, barrier, rewarm=100, settle_s=0.5):
    """W warm-up steps (and at least settle_s of load so the power-capped clock has settled), a barrier, `rewarm`
    un-synchronised steps so that the timed region starts on a busy GPU (a barrier lets the chip idle and then burst),
    then exactly `steps` steps between two CUDA events."""
    import torch
    t_w = time.time()
    i = 0
    while i < warmup or time.time() - t_w < settle_s:
        step(i)
        i += 1
        if i % 50 == 0:
            torch.cuda.synchronize()
    barrier()
    for j in range(rewarm):
        step(i + j)
    i += rewarm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for j in range(steps):
        step(i + j)
    e1.record()
    barrier()
    t1 = time.time()
    return e0.elapsed_time(e1), t0, t1, i + steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the records under 'also'")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    topk = wl["topk"]
    rows = args.rows or wl["rows"]
    dp = args.workload == "cfg4" and world > 1
    config = {"workload": wl["desc"],
              "rows_per_step_per_gpu": rows // world if dp else rows,
              "l2_policy": "per-step working set (activations+deltas+split-K slabs, ~0.7 GB at 16384 rows) exceeds the 126 MB L2; "
                           "8 rotating coordinate batches",
              "input_pipeline": "GPU arm: the next batch's coordinates are encoded on the plan's low-priority side stream while "
                                "a step runs (npp_encode_prefetch, counted in gpu_launches); steps run on a priority -1 stream",
              "timing": "warm-up, barrier, 100 un-synchronised re-warm steps, then K steps between two CUDA events on the "
                        "launching stream, max over ranks"}

    if args.impl == "reference":
        if rank != 0:
            return
        ref_rows = min(rows, 16384)
        steps = max(1, min(args.steps, 20))
        warm = max(1, min(args.warmup, 3))
        val, ms, threads, kind, sample = cpu_reference(topk, ref_rows, steps, warm, wl["res"])
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import npp_b200  # noqa: F401

    torch.cuda.set_device(local_rank)
    # The train steps run on a high-priority stream: the plan's input-prefetch stream (default = lowest priority) then
    # only gets the SMs the step's kernels leave idle (the chain kernels occupy 128 of the 148).
    torch.cuda.set_stream(torch.cuda.Stream(device=local_rank, priority=-1))
    if world > 1:
        import datetime
        # a rank that falls out of step must fail within minutes, not after NCCL's 10-minute default
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ------------------------------------------------------------------ headline
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi needs ~1 s to come up
    launches_per_step = None
    if dp:
        from npp_b200.dp import DataParallelStep
        fit = Fit("cfg4", rank, dev, my_rows=rows // world, image_seed=0)
        dps = DataParallelStep(fit.plan)

        def step(i):
            b = i % fit.nb
            dps(fit.coords[b], fit.target[b], fit.mask[b], fit.lr, rows)
    elif args.workload == "cfg5":
        fit = Fit("cfg5", rank, dev)
        step = fit.step
    else:
        fit = Fit(args.workload, rank, dev, my_rows=rows)
        step = fit.step
    plan = fit.plan
    my_rows = fit.rows
    total_ms, t0, t1, done = timed_loop(step, args.steps, args.warmup, barrier, settle_s=0.0 if dp else 0.5)
    launches_per_step = plan.launch_count() if not dp else dps.launches
    clocks = sampler.stop(t0, t1) if sampler else None
    total_ms = max_over_ranks(total_ms)
    total_rows = (rows if dp else rows * world) * args.steps
    value = total_rows / (total_ms * 1e-3)
    final_loss = fit.loss.item() if not dp else float(dps.last_loss)

    # ---- per-kernel-class timing (CUDA events on the launch stream) over another K steps
    prof = None
    if not dp:
        plan.profile(True)
        for i in range(args.steps):
            step(done + i)
        prof = plan.profile_read()
        plan.profile(False)

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs, D2H of the loss
    # Inputs go through two device slots filled by a copy stream (the usual double-buffered loader): the H2D copies of
    # step i+1 are issued before the host blocks on the loss of step i, every copy and every loss read stays inside
    # the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    slots = [(torch.empty_like(fit.coords[0]), torch.empty_like(fit.target[0]), torch.empty_like(fit.mask[0]))
             for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for ev in ev_free:
        ev.record(main_stream)

    def issue_copies(i):
        b, k = i % fit.nb, i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])          # the step that last used this slot has finished
            slots[k][0].copy_(fit.host_coords[b], non_blocking=True)
            slots[k][1].copy_(fit.host_target[b], non_blocking=True)
            slots[k][2].copy_(fit.host_mask[b], non_blocking=True)
            if not dp:
                plan.prefetch_encode(slots[k][0])       # encoded as soon as the coordinates have landed
            ev_ready[k].record(copy_stream)

    def step_e2e(i):
        k = i % 2
        main_stream.wait_event(ev_ready[k])
        c, t, mk = slots[k]
        if dp:
            l = dps(c, t, mk, fit.lr, rows)
            ev_free[k].record(main_stream)
            issue_copies(i + 1)
            return l.item()
        plan.train_step(c, t, mk, fit.lr, fit.loss)
        ev_free[k].record(main_stream)
        issue_copies(i + 1)                             # next step's inputs: copies + encoding overlap this step
        return fit.loss.item()

    issue_copies(0)
    for i in range(3):
        step_e2e(i)
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    for i in range(3, 23):                               # re-warm after the barrier
        step_e2e(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2 = time.time()
    e0.record()
    for i in range(23, 23 + args.steps):   # the pipeline keeps running: each timed step issues the next step's copies
        step_e2e(i)
    e1.record()
    barrier()
    # the end-to-end loop blocks on the loss every step: the short idle gaps let a power-capped GPU clock higher
    # inside the kernels than the back-to-back loop above does, which is why E can exceed `value`
    clocks_e2e = sampler2.stop(t2, time.time()) if sampler2 else None
    e2e_value = total_rows / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    h2d = my_rows * (2 + 3 + 1) * 4
    d2h = 4

    # ------------------------------------------------------------------ extra records
    also = {}
    if not args.no_extras:
        extra_steps = max(20, min(args.steps, 200))
        names = [n for n in ("cfg2", "cfg2-top1", "cfg3") if n != args.workload]
        del fit, plan
        torch.cuda.empty_cache()
        for name in names:
            f = Fit(name, rank, dev)
            ms, _, _, _ = timed_loop(f.step, extra_steps, 10, barrier, rewarm=50, settle_s=0.2)
            ms = max_over_ranks(ms)
            v = f.rows * world * extra_steps / (ms * 1e-3)
            also[name] = {"value": v, "unit": "samples/s", "ms_per_step": ms / extra_steps, "steps": extra_steps,
                          "rows_per_step_per_gpu": f.rows, "scaling": "weak", "workload": WORKLOADS[name]["desc"],
                          "whole_step_tflops_per_gpu": v / world * FLOP_PER_SAMPLE[f.topk] / 1e12}
            del f
            torch.cuda.empty_cache()
        if args.workload != "cfg5":
            also["cfg5"] = bench_cfg5(rank, world, dev, barrier, max_over_ranks)
        if not dp:      # N = 1: the same 2^18-row step on one GPU, the denominator of the strong-scaling curve
            also["dp_cfg4"] = bench_dp_cfg4(rank, world, dev, barrier, max_over_ranks)
        if world == 1:
            try:
                ref_gpu = reference_on_gpu(topk, my_rows, local_rank)
                if ref_gpu is not None:
                    also["reference_gpu"] = ref_gpu
            except Exception as e:
                also["reference_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}

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
        roofline = {"bound": "tensor",
                    "kernel": "npp_gemm_kmajor (one launch per step: forward GEMM chain + RGB head/loss epilogue + dgrad GEMM "
                              "chain, tcgen05 kind::f16)",
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
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if dp else "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (f32 master weights, loss, Adam)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "sm_mhz": (clocks_e2e or {}).get("sm_mhz")},
            "gpu_launches": (launches_per_step or 0) * args.steps, "final_loss": final_loss, "roofline": roofline,
            "also": also}
    if world == 1 and not args.no_cpu_baseline:
        cval, cms, cthreads, kind, sample = cpu_reference(topk, 8192, 6, 2, wl["res"])
        line["cpu_baseline"] = {"value": cval, "unit": "samples/s", "cores": cthreads, "kind": kind, "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_cfg5(rank, world, dev, barrier, max_over_ranks, fits=192, iters=20):
    """192 independent K=1 fits dealt round-robin over the ranks, no collective: every fit re-initialises the weights
    and Adam state on the device and runs `iters` steps.  The re-initialisation is timed on its own as well, so that
    fits/s can be quoted for the reference's 2001 iterations per fit (options/arg_config.py:96) as
    world / (t_init + 2001 * t_step) from the two measured times."""
    import torch
    f = Fit("cfg5", rank, dev)
    mine = [k for k in range(fits) if k % world == rank]
    for i in range(30):
        f.step(i)
    f.plan.reset_parameters(seed=0, on_device=True)
    barrier()
    for i in range(50):
        f.step(i)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    i = 0
    for k in mine:
        f.plan.reset_parameters(seed=k, on_device=True)
        f.plan.adam_steps = 0
        for _ in range(iters):
            f.step(i)
            i += 1
    e1.record()
    for k in mine:                        # the re-initialisations alone
        f.plan.reset_parameters(seed=k, on_device=True)
    e2.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_init = max_over_ranks(e1.elapsed_time(e2))
    n = max(len(mine), 1)
    t_init = ms_init / n
    t_step = max(ms - ms_init, 0.0) / n / iters
    return {"value": fits * iters * f.rows / (ms * 1e-3), "unit": "samples/s", "fits": fits, "iters_per_fit_timed": iters,
            "ms_per_step": t_step, "ms_per_reinit": t_init,
            "fits_per_s_at_2001_iters": world / ((t_init + 2001 * t_step) * 1e-3),
            "scaling": "weak (independent fits, no collective)", "workload": WORKLOADS["cfg5"]["desc"]}


def bench_dp_cfg4(rank, world, dev, barrier, max_over_ranks, steps=30):
    """cfg4 at this world size: one 2^18-row batch per step split over the ranks (strong scaling), gradients summed by
    NCCL all-reduce overlapped with the weight-gradient GEMMs (npp_b200.dp.DataParallelStep)."""
    import torch
    from npp_b200.dp import DataParallelStep
    rows = WORKLOADS["cfg4"]["rows"]
    f = Fit("cfg4", rank, dev, my_rows=rows // world, nb=4, image_seed=0)
    dps = DataParallelStep(f.plan)

    def step(i):
        b = i % f.nb
        dps(f.coords[b], f.target[b], f.mask[b], f.lr, rows)
    # every rank must run the SAME number of steps (each one holds collectives): no time-based warm-up here
    ms, _, _, _ = timed_loop(step, steps, 5, barrier, rewarm=10, settle_s=0.0)
    ms = max_over_ranks(ms)
    exposed = None
    if world > 1:       # the same steps without the collectives (the ranks' weights drift apart: timing only, done last)
        dps.skip_allreduce = True
        ms0, _, _, _ = timed_loop(step, steps, 3, barrier, rewarm=10, settle_s=0.0)
        exposed = (ms - max_over_ranks(ms0)) / steps
    return {"allreduce_exposed_ms_per_step": exposed, "value": rows * steps / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms / steps, "steps": steps,
            "n_gpus": world, "rows_per_step": rows, "rows_per_step_per_gpu": rows // world, "scaling": "strong",
            "collective": "ncclAllReduce(sum) of the fp32 gradients in layer-group buckets on a side stream, overlapped "
                          "with the remaining weight-gradient GEMM groups; identical Adam on every rank",
            "launches_per_step": dps.launches, "workload": WORKLOADS["cfg4"]["desc"]}


if __name__ == "__main__":
    main()
