"""CPU oracle for the NPP-Net train-step hot path -- TEST INFRASTRUCTURE ONLY.

This is a plain-numpy (fp32) restatement of the reference's algorithm for the path
``encoding -> MLP forward -> sigmoid + masked MSE -> backward -> Adam``.  It exists to check
the CUDA kernels; nothing in the product package imports it (only tests/, the cpu_baseline /
--impl reference legs of bench.py and tools/bench_search_fits.py, and __graft_entry__.smoke() do).

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against the reference itself: tests/golden/make_golden.py imports the
reference modules from /root/reference, runs them on seeded inputs and stores the results in
tests/golden/*.npz (make_golden_light.py for the search-stage network and encoders); tests/test_oracle_golden.py checks every function below against them.

Every function cites the reference file:line (relative to the reference root) it restates.
The backward pass is derived by hand (the reference relies on torch autograd), which makes it an
independent check of the fused dgrad/wgrad kernels rather than a re-run of the same code.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- encoding
def encoder_tables(selected_angles, selected_periods, freq_scales, freq_offsets, angle_offsets):
    """Per (direction, augmentation) constants of one proposal: cos(theta), sin(theta), period.

    models/embedder.py:117-124 -- loop order freq_scale -> freq_offset -> idx -> angle_offset;
    freq = (period[idx] + freq_offset) * freq_scale; theta = deg2rad(angle[idx] + angle_offset).
    Returns three float32 arrays of shape [2, n_aug]."""
    n_aug = len(freq_scales) * len(freq_offsets) * len(angle_offsets)
    cos_t = np.zeros((2, n_aug), F32)
    sin_t = np.zeros((2, n_aug), F32)
    period = np.zeros((2, n_aug), F32)
    for idx in range(2):
        a = 0
        for fs in freq_scales:
            for fo in freq_offsets:
                for ao in angle_offsets:
                    p = (F32(selected_periods[idx]) + F32(fo)) * F32(fs)
                    theta = (F32(selected_angles[idx]) + F32(ao)) * F32(np.pi / 180.0)
                    cos_t[idx, a] = np.cos(theta, dtype=F32)
                    sin_t[idx, a] = np.sin(theta, dtype=F32)
                    period[idx, a] = p
                    a += 1
    return cos_t, sin_t, period


def torch_remainder(a, b):
    """torch.remainder for floats (the % at models/embedder.py:127): fmod, then sign of divisor."""
    r = np.fmod(a, b).astype(F32)
    fix = (r != 0) & ((r < 0) != (b < 0))
    return np.where(fix, (r + b).astype(F32), r).astype(F32)


def encode_periodic(coords, cos_t, sin_t, period, res, include_input=True):
    """Embedder_periodic.embed (models/embedder.py:102-148) for one proposal.

    coords [N,2] = (row y, col x); res = (H, W).  Output [N, 2*(include_input + 2*n_aug)]:
    fn_x list ([x_norm], sin, cos, ...) then fn_y list."""
    coords = np.asarray(coords, F32)
    y = coords[:, 0:1]
    x = coords[:, 1:2]
    lists = []
    for idx in range(2):
        cols = []
        if include_input:
            v = x / F32(res[1]) if idx == 0 else y / F32(res[0])          # embedder.py:107-108
            cols.append(((v - F32(0.5)) * F32(2)).astype(F32))
        for a in range(cos_t.shape[1]):
            proj = (y * cos_t[idx, a]).astype(F32) + (x * sin_t[idx, a]).astype(F32)
            frac = (torch_remainder(proj.astype(F32), period[idx, a]) / period[idx, a]).astype(F32)
            phi = ((frac * F32(2)).astype(F32) * F32(np.pi)).astype(F32)   # embedder.py:127
            cols.append(np.sin(phi, dtype=F32))
            cols.append(np.cos(phi, dtype=F32))
        lists.append(np.concatenate(cols, -1))
    return np.concatenate(lists, -1).astype(F32)


def encode_fourier(u, freqs, include_input=True):
    """Embedder.embed with input_dims=1, sampling='gaussian' (models/embedder.py:41-44,56):
    cat([u, sin(u f_0), cos(u f_0), ...], -1) applied to the whole [N,B] matrix."""
    u = np.asarray(u, F32)
    outs = [u] if include_input else []
    for f in np.asarray(freqs, F32).reshape(-1):
        a = (u * f).astype(F32)
        outs.append(np.sin(a, dtype=F32))
        outs.append(np.cos(a, dtype=F32))
    return np.concatenate(outs, -1).astype(F32)


def encode(coords, tables, freqs, res):
    """Full network input (NPP_completion/train.py:93-105): per proposal periodic -> Fourier, then
    concatenated along features.  tables = list over proposals of (cos_t, sin_t, period)."""
    return np.concatenate(
        [encode_fourier(encode_periodic(coords, c, s, p, res), freqs) for (c, s, p) in tables], 1)


# ------------------------------------------------------------------------------ model
def snake(z):
    """SnakeActivation with a=1 (models/activations.py:29-35): x + sin(x)^2."""
    return (z + np.square(np.sin(z, dtype=F32))).astype(F32)


def snake_grad(z):
    return (F32(1) + np.sin(F32(2) * z, dtype=F32)).astype(F32)


def relu(z):
    """F.relu (models/networks.py:66-67), used when activation != 'snake'."""
    return np.maximum(z, F32(0)).astype(F32)


def relu_grad(z):
    return (z > 0).astype(F32)


def _act(activation):
    """(function, derivative) of the hidden activation: 'snake' or, like the reference, relu for anything else."""
    return (snake, snake_grad) if activation == "snake" else (relu, relu_grad)


def _lin(p, name, x):
    return (x @ p[name + ".weight"].T + p[name + ".bias"]).astype(F32)


def forward(p, enc, depth=8, skips=(4,), topk_model=True, ch1=None, activation="snake"):
    """NPP_Net.forward (models/networks.py:56-95) / NPP_Net_top1.forward (:145-173).

    p: dict of float32 arrays keyed like the reference state_dict.  enc [N, K*462].
    Returns (logits, cache) where cache holds every pre-/post-activation (for the backward and
    for per-layer parity checks)."""
    enc = np.asarray(enc, F32)
    snake = _act(activation)[0]        # shadows the module-level function inside this forward
    if ch1 is None:
        ch1 = p["periodic_linears.0.weight"].shape[1]
    enc1, enc_aux = enc[:, :ch1], enc[:, ch1:]
    c = {"enc1": enc1, "enc_aux": enc_aux, "a": {}, "z": {}, "h": {}}
    h = enc1
    for i in range(depth):
        name = f"periodic_linears.{i}"
        c["a"][name] = h
        z = _lin(p, name, h)
        c["z"][name] = z
        h = snake(z)
        c["h"][name] = h
        if i in skips:
            h = np.concatenate([enc1, h], -1)                      # networks.py:70-71
    c["a"]["feature_linear1"] = h
    f1 = _lin(p, "feature_linear1", h)                             # networks.py:73
    c["h"]["feature_linear1"] = f1
    if topk_model:
        a = np.concatenate([f1, enc_aux], -1)                      # networks.py:76
        c["a"]["scale_linears.0"] = a
        z = _lin(p, "scale_linears.0", a)
        c["z"]["scale_linears.0"] = z
        hs = snake(z)
        c["h"]["scale_linears.0"] = hs
        c["a"]["feature_linear2"] = hs
        f2 = _lin(p, "feature_linear2", hs)                        # networks.py:84
        c["h"]["feature_linear2"] = f2
        a = np.concatenate([f1, f2], -1)                           # networks.py:85
    else:
        a = f1                                                     # networks.py:159
    c["a"]["pos_linears.0"] = a
    z = _lin(p, "pos_linears.0", a)
    c["z"]["pos_linears.0"] = z
    hp = snake(z)
    c["h"]["pos_linears.0"] = hp
    c["a"]["rgb_linear"] = hp
    logits = _lin(p, "rgb_linear", hp)                             # networks.py:94
    c["logits"] = logits
    return logits, c


def sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def mse_l2(pred, target, mask=None):
    """img2mse(x, y, 'l2', adaptive, mask) (models/mse_calculator.py:13-27)."""
    d = (pred - target).astype(F32)
    if mask is not None:
        d = (d * mask + (F32(1) - mask) * d * F32(0.3)).astype(F32)
    return F32(np.mean(np.square(d), dtype=np.float64))


def mse_l2_grad_logits(logits, target, mask=None, n_norm=None):
    """d mean((w (sigmoid(z) - y))^2) / dz, w = m + 0.3 (1 - m); mean over n_norm*3 elements."""
    yh = sigmoid(logits)
    n_norm = logits.shape[0] if n_norm is None else n_norm
    w = F32(1) if mask is None else (mask + (F32(1) - mask) * F32(0.3)).astype(F32)
    g_pred = (F32(2) * w * w * (yh - target) / F32(3 * n_norm)).astype(F32)
    return (g_pred * yh * (F32(1) - yh)).astype(F32)


def backward(p, c, g_logits, depth=8, skips=(4,), topk_model=True, activation="snake"):
    """Hand-derived backward of `forward` (the reference uses autograd: loss.backward(),
    NPP_completion/train.py:253).  Returns (grads dict, deltas dict) -- deltas[name] is dL/dz of
    that layer (dL/d output for the activation-free feature_linear1/2)."""
    ch1 = c["enc1"].shape[1]
    W = p["feature_linear1.weight"].shape[0]
    snake_grad = _act(activation)[1]
    grads, deltas = {}, {}

    def lin_bwd(name, delta):
        deltas[name] = delta
        grads[name + ".weight"] = (delta.T @ c["a"][name]).astype(F32)
        grads[name + ".bias"] = delta.sum(0, dtype=np.float64).astype(F32)
        return (delta @ p[name + ".weight"]).astype(F32)

    d_hp = lin_bwd("rgb_linear", np.asarray(g_logits, F32))
    d_a = lin_bwd("pos_linears.0", (d_hp * snake_grad(c["z"]["pos_linears.0"])).astype(F32))
    if topk_model:
        d_f1 = d_a[:, :W].copy()
        d_f2 = d_a[:, W:]
        d_hs = lin_bwd("feature_linear2", d_f2)
        d_a = lin_bwd("scale_linears.0", (d_hs * snake_grad(c["z"]["scale_linears.0"])).astype(F32))
        d_f1 = (d_f1 + d_a[:, :W]).astype(F32)                      # fan-out of feature1
    else:
        d_f1 = d_a
    d_h = lin_bwd("feature_linear1", d_f1)
    for i in reversed(range(depth)):
        name = f"periodic_linears.{i}"
        if i in skips:
            d_h = d_h[:, ch1:]                                       # cat([enc1, h]): enc needs no grad
        d_h = lin_bwd(name, (d_h * snake_grad(c["z"][name])).astype(F32))
    return grads, deltas


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (torch/optim/adam.py, built at models/helpers.py:164):
    m.lerp_(g, 1-b1); v = b2 v + (1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  In place."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = F32(lr / bc1)
    inv_sqrt_bc2 = F32(1.0 / np.sqrt(bc2))
    for k in g:
        m[k] += (g[k] - m[k]) * F32(1.0 - beta1)
        v[k] *= F32(beta2)
        v[k] += F32(1.0 - beta2) * g[k] * g[k]
        p[k] -= step_size * (m[k] / (np.sqrt(v[k]) * inv_sqrt_bc2 + F32(eps)))


def lr_schedule(step_index, lrate=5e-4, lrate_decay=500):
    """LR actually used by optimizer.step() number `step_index` (1-based) in NPP_completion/train.py:
    the rewrite at :258-263 happens after step() with the pre-increment global_step, so step k uses
    lrate * 0.1 ** (max(k-2, 0) / (lrate_decay*100))."""
    return lrate * (0.1 ** (max(step_index - 2, 0) / (lrate_decay * 100)))


def train_step(p, m, v, step, enc, target, mask, lr, depth=8, skips=(4,), topk_model=True, activation="snake"):
    """One whole reference iteration with --loss_type l2 (NPP_completion/train.py:187-254)."""
    logits, c = forward(p, enc, depth, skips, topk_model, activation=activation)
    pred = sigmoid(logits)
    loss = mse_l2(pred, target, mask)
    g = mse_l2_grad_logits(logits, target, mask)
    grads, _ = backward(p, c, g, depth, skips, topk_model, activation=activation)
    adam_step(p, grads, m, v, step, lr)
    return loss, pred


# ----------------------------------------------------------------------- initialisation
def init_params(rng, topk=3, width=512, depth=8, skips=(4,), ch=462):
    """nn.Linear default init, U(-1/sqrt(in), 1/sqrt(in)) for weight and bias
    (weights_init_normal only touches Conv/BatchNorm, models/helpers.py:65-71,140-141)."""
    p = {}

    def lin(name, out, inp):
        b = 1.0 / np.sqrt(inp)
        p[name + ".weight"] = rng.uniform(-b, b, (out, inp)).astype(F32)
        p[name + ".bias"] = rng.uniform(-b, b, (out,)).astype(F32)

    lin("periodic_linears.0", width, ch)
    for i in range(1, depth):
        lin(f"periodic_linears.{i}", width, width + ch if (i - 1) in skips else width)
    if topk > 1:
        lin("scale_linears.0", width, width + ch * (topk - 1))
        lin("pos_linears.0", width // 2, 2 * width)
    else:
        lin("pos_linears.0", width // 2, width)
    lin("feature_linear1", width, width)
    lin("feature_linear2", width, width)
    lin("alpha_linear", 1, width)
    lin("rgb_linear", 3, width // 2)
    return p


# ------------------------------------------------------------- search stage (NPP_Net_light)
def encode_search_positional(coords, freqs, res):
    """Embedder.embed in search mode (models/embedder.py:51-56 with input_dims=2, :76-80):
    coordinates normalised in place, row by res[0] and column by res[1], then
    cat([u, sin(u f_0), cos(u f_0), ...], -1) on the [N,2] matrix -> [N, 2 + 4*n_freq]."""
    c = np.asarray(coords, F32)
    u = np.stack([((c[:, 0] / F32(res[0])).astype(F32) - F32(0.5)) * F32(2),
                  ((c[:, 1] / F32(res[1])).astype(F32) - F32(0.5)) * F32(2)], 1).astype(F32)
    return encode_fourier(u, freqs)


def encode_search(coords, table, freqs, res):
    """The two tensors NPP_proposal/search.py:104-108 builds for one candidate periodicity:
    (embedder.embed(coords) [N,42], embedder_periodic.embed(coords) [N,20]); the periodic encoder has
    neither the raw input nor the Fourier expansion in search mode (embedder.py:93-95)."""
    cos_t, sin_t, period = table
    return (encode_search_positional(coords, freqs, res),
            encode_periodic(coords, cos_t, sin_t, period, res, include_input=False))


def forward_light(p, pos, per, depth=4, skips=(4,)):
    """NPP_Net_light.forward with len(freq_scales) == 1 (models/networks.py:222-263): the scale MLP is
    skipped, `pos` (x) joins after feature_linear1.  Same cache layout as `forward`."""
    pos = np.asarray(pos, F32)
    per = np.asarray(per, F32)
    c = {"enc1": per, "pos": pos, "a": {}, "z": {}, "h": {}}
    h = per
    for i in range(depth):
        name = f"periodic_linears.{i}"
        c["a"][name] = h
        z = _lin(p, name, h)
        c["z"][name] = z
        h = snake(z)
        c["h"][name] = h
        if i in skips:
            h = np.concatenate([per, h], -1)                       # networks.py:234-235
    c["a"]["feature_linear1"] = h
    f1 = _lin(p, "feature_linear1", h)                             # networks.py:237
    c["h"]["feature_linear1"] = f1
    a = np.concatenate([f1, pos], -1)                              # networks.py:250
    c["a"]["pos_linears.0"] = a
    z = _lin(p, "pos_linears.0", a)
    c["z"]["pos_linears.0"] = z
    hp = snake(z)
    c["h"]["pos_linears.0"] = hp
    c["a"]["rgb_linear"] = hp
    logits = _lin(p, "rgb_linear", hp)                             # networks.py:262
    c["logits"] = logits
    return logits, c


def backward_light(p, c, g_logits, depth=4, skips=(4,)):
    """Hand-derived backward of `forward_light` (reference: loss.backward(), NPP_proposal/search.py:135)."""
    ch1 = c["enc1"].shape[1]
    W = p["feature_linear1.weight"].shape[0]
    grads, deltas = {}, {}

    def lin_bwd(name, delta):
        deltas[name] = delta
        grads[name + ".weight"] = (delta.T @ c["a"][name]).astype(F32)
        grads[name + ".bias"] = delta.sum(0, dtype=np.float64).astype(F32)
        return (delta @ p[name + ".weight"]).astype(F32)

    d_hp = lin_bwd("rgb_linear", np.asarray(g_logits, F32))
    d_a = lin_bwd("pos_linears.0", (d_hp * snake_grad(c["z"]["pos_linears.0"])).astype(F32))
    d_h = lin_bwd("feature_linear1", d_a[:, :W])                   # the positional columns need no gradient
    for i in reversed(range(depth)):
        name = f"periodic_linears.{i}"
        if i in skips:
            d_h = d_h[:, ch1:]
        d_h = lin_bwd(name, (d_h * snake_grad(c["z"][name])).astype(F32))
    return grads, deltas


def train_step_light(p, m, v, step, pos, per, target, lr, depth=4, skips=(4,)):
    """One search iteration with --loss_type l2 and no mask (NPP_proposal/search.py:112-136)."""
    logits, c = forward_light(p, pos, per, depth, skips)
    pred = sigmoid(logits)
    loss = mse_l2(pred, target, None)
    g = mse_l2_grad_logits(logits, target, None)
    grads, _ = backward_light(p, c, g, depth, skips)
    adam_step(p, grads, m, v, step, lr)
    return loss, pred


def init_params_light(rng, width=256, depth=4, skips=(4,), ch=20, ch_pos=42):
    """nn.Linear default init for NPP_Net_light (models/networks.py:198-214, scale_dim == 0)."""
    p = {}

    def lin(name, out, inp):
        b = 1.0 / np.sqrt(inp)
        p[name + ".weight"] = rng.uniform(-b, b, (out, inp)).astype(F32)
        p[name + ".bias"] = rng.uniform(-b, b, (out,)).astype(F32)

    lin("periodic_linears.0", width, ch)
    for i in range(1, depth):
        lin(f"periodic_linears.{i}", width, width + ch if (i - 1) in skips else width)
    lin("scale_linears.0", width, width)
    lin("pos_linears.0", width // 2, ch_pos + width)
    lin("feature_linear1", width, width)
    lin("feature_linear2", width, width)
    lin("alpha_linear", 1, width)
    lin("rgb_linear", 3, width // 2)
    return p
