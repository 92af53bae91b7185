"""The UNMODIFIED reference modules as the baseline arm of bench.py (test infrastructure, never on the product path).

`__graft_entry__.build()` copies the reference checkout (when /root/reference is present, i.e. in the build
container) into ``baseline/_ref/`` -- git-ignored, so no reference source enters the history, but not
gpurun-ignored, so it travels to the GPU box.  This module imports the reference's own

    models/embedder.py   get_embedder, Embedder, Embedder_periodic      (encoding table, train.py:93-105)
    models/networks.py   NPP_Net / NPP_Net_top1                          (networks.py:56-95, 145-173)
    models/helpers.py    render                                          (helpers.py:41-62)
    models/mse_calculator.py  img2mse(..., 'l2', ...)                    (mse_calculator.py:13-27)

from there with the import shims SURVEY.md section 8c lists, builds the per-iteration body of
NPP_completion/train.py:164-263 around them (row gather from the precomputed encoding table, render, img2mse,
backward, torch.optim.Adam step, learning-rate rewrite) and times it.  Anomaly detection stays as the reference
ships it (torch.autograd.set_detect_anomaly(True) at models/embedder.py:2) unless asked otherwise.
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")


def reference_root():
    for cand in (os.environ.get("NPP_REFERENCE_ROOT"), REF_COPY, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "models", "networks.py")):
            return cand
    return None


def install_copy(src="/root/reference", dst=REF_COPY):
    """Copy the reference checkout next to this file (build container only).  Returns the path or None."""
    import shutil
    if not os.path.isfile(os.path.join(src, "models", "networks.py")):
        return dst if os.path.isdir(dst) else None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "teaser.jpg"))
    return dst


_mods = None


def import_reference():
    """(embedder, networks, mse_calculator, render) of the reference, imported from its own files."""
    global _mods
    if _mods is not None:
        return _mods
    import torch
    root = reference_root()
    if root is None:
        raise ImportError("no reference checkout: neither baseline/_ref nor /root/reference exists")
    for name in ("torch_dct",):                               # only used by DCT helpers of robust_loss_pytorch/util.py
        sys.modules.setdefault(name, types.ModuleType(name))
    for k in ("float", "int", "bool"):                       # numpy aliases removed in numpy >= 1.24 (utils/miscs.py:29)
        if not hasattr(np, k):
            setattr(np, k, {"float": float, "int": int, "bool": bool}[k])
    own = [p for p in sys.path if os.path.isdir(os.path.join(p, "models")) and p != root]
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path[:] = [p for p in sys.path if p not in own]
    sys.path.insert(0, os.path.join(root, "externel_lib"))
    sys.path.insert(0, root)
    import contextlib

    def _import_all():
        import models.embedder as emb
        import models.networks as net
        import models.mse_calculator as mse          # prints os.getcwd() at import (mse_calculator.py:9)
        render = None
        try:
            if not torch.cuda.is_available():
                # models/helpers.py:8-9 builds AdaptiveLossFunction(device=0) at import: without a GPU that is a CUDA
                # call; the adaptive loss is not on the 'l2' path, so it is stubbed for the import only
                import externel_lib.robust_loss_pytorch as rl
                rl.AdaptiveLossFunction = lambda *a, **k: torch.nn.Module()
            import models.helpers as helpers
            render = helpers.render
        except Exception as e:                               # pragma: no cover - environment dependent
            print(f"[reference_arm] models/helpers.py not importable here ({type(e).__name__}: {e}); "
                  "render = sigmoid(network_query_fn(...)) restated from helpers.py:41-62", file=sys.stderr)
        return emb, net, mse, render

    try:
        with contextlib.redirect_stdout(sys.stderr):         # keep bench.py's stdout to its one JSON line
            emb, net, mse, render = _import_all()
    finally:
        ref_mods = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = [p for p in sys.path if p not in (root, os.path.join(root, "externel_lib"))] + []
        for p in own:
            if p not in sys.path:
                sys.path.insert(0, p)
    if render is None:
        def render(x, xp, args, network_query_fn, network_fn):
            raw = network_query_fn(x, xp, network_fn)
            return torch.sigmoid(raw) if args.normalize_type == 1 else torch.tanh(raw)
    _mods = (emb, net, mse, render)
    return _mods


class ReferenceFit:
    """One image fit with the reference modules: the state NPP_completion/train.py holds between iterations."""

    def __init__(self, res, angles, periods, image, pool_coords, *, topk=3, device="cpu", seed=0, anomaly=True,
                 freq_scales=(1,), freq_offsets=(0, -1, 1, 0.5, -0.5), angle_offsets=(0,), netdepth=8, netwidth=512,
                 multires=10, lrate=5e-4, lrate_decay=500):
        import torch
        emb, net, mse, render = import_reference()
        torch.autograd.set_detect_anomaly(bool(anomaly))
        self.torch, self.mse, self.render = torch, mse, render
        self.device = torch.device(device)
        self.lrate, self.lrate_decay = lrate, lrate_decay
        torch.manual_seed(seed)
        # create_npp_net (helpers.py:75-175) without the adaptive-loss / perceptual parameter lists
        embedder, freq_nerf = emb.get_embedder(multires, 0, res)
        sel_a, sel_p = torch.Tensor(angles), torch.Tensor(periods)
        periodic, chans = [], []
        for i in range(topk):
            e, ch = emb.get_embedder(multires, 0, res, selected_angles=sel_a[i], selected_periods=sel_p[i],
                                     freq_scales=list(freq_scales), freq_offsets=list(freq_offsets),
                                     angle_offsets=list(angle_offsets))
            periodic.append(e)
            chans.append(ch)
        chans = np.array(chans)
        kw = dict(D=netdepth, W=netwidth, freq_nerf=freq_nerf, input_ch_periodic=chans[:1].sum(),
                  freq_scales=list(freq_scales), freq_offsets=list(freq_offsets), angle_offsets=list(angle_offsets),
                  output_ch=3, skips=[4], activation="snake")
        if topk > 1:
            self.model = net.NPP_Net(input_ch_periodic_aux=chans[1:].sum(), **kw).to(self.device)
        else:
            self.model = net.NPP_Net_top1(**kw).to(self.device)
        self.optimizer = torch.optim.Adam(params=list(self.model.parameters()), lr=lrate, betas=(0.9, 0.999))
        self.args = types.SimpleNamespace(normalize_type=1)
        self.query = lambda x, xp, fn: fn(x, xp)              # run_network with one chunk (helpers.py:26-36)
        # encoding table of the pixel pool (train.py:93-105): periodic encoder, then the Fourier encoder, per proposal
        t0 = time.perf_counter()
        pool = torch.as_tensor(pool_coords, dtype=torch.float32)
        cols = [embedder.embed(periodic[i].embed(pool.clone())) for i in range(topk)]
        self.table = torch.cat(cols, 1).to(self.device)
        self.table_build_s = time.perf_counter() - t0
        img = torch.as_tensor(image, dtype=torch.float32)
        self.target = img[pool[:, 0].long(), pool[:, 1].long()].to(self.device)
        self.global_step = 0

    def step(self, sel, mask=None, autocast=None):
        """Iteration body of NPP_completion/train.py:164-263 for the rows `sel` of the pool, loss_type 'l2'."""
        torch = self.torch
        x = self.table[sel]                                   # train.py:178-181
        y = self.target[sel]
        ctx = torch.autocast(self.device.type, dtype=autocast) if autocast is not None else _null()
        with ctx:
            pred = self.render(None, x, self.args, self.query, self.model)
            loss = self.mse.img2mse(pred.float(), y, "l2", None, mask)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        self.global_step += 1
        new_lrate = self.lrate * (0.1 ** (self.global_step / (self.lrate_decay * 1000)))   # train.py:258-263
        for g in self.optimizer.param_groups:
            g["lr"] = new_lrate
        return loss


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def time_fit(fit, rows, steps, warmup, seed=0, autocast=None):
    """samples/s of `steps` iterations of `rows` random pool rows each (host pixel draw as in train.py:172 kept
    outside: only the model step is timed, like the GPU arm whose inputs are resident)."""
    torch = fit.torch
    rng = np.random.default_rng(seed)
    n_pool = fit.table.shape[0]
    sels = [torch.as_tensor(rng.choice(n_pool, rows, replace=False), device=fit.device) for _ in range(steps + warmup)]
    cuda = fit.device.type == "cuda"
    for i in range(warmup):
        fit.step(sels[i], autocast=autocast)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    last = None
    for i in range(warmup, warmup + steps):
        last = fit.step(sels[i], autocast=autocast)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
    else:
        dt = time.perf_counter() - t0
    return rows * steps / dt, dt / steps * 1e3, float(last.detach())
