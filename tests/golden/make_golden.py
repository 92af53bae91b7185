"""Generate the golden vectors that pin oracle/npp_oracle.py to the reference.

Run in the build container (the reference is mounted read-only at /root/reference and is NOT
available on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own models/embedder.py, models/networks.py and models/mse_calculator.py
(with the import shims SURVEY.md section 8c lists), runs them with torch 2.11 fp32 on CPU on seeded
inputs and writes small .npz fixtures next to this file:

  golden_encoding.npz   Embedder_periodic + Embedder outputs for integer coordinates
  golden_topk.npz       NPP_Net      (K=3, W=64): per-layer activations, loss, grads, Adam trajectory
  golden_top1.npz       NPP_Net_top1 (K=1, W=64): same
  golden_relu.npz       NPP_Net (K=3, W=64) with activation='relu' (models/networks.py:51-54,66-69): same

W=64 keeps the fixtures small; the reference modules and the oracle are width-agnostic.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NPP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ["torch_dct"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "externel_lib"))
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    import models.embedder as emb
    import models.networks as net
    import models.mse_calculator as mse
    os.chdir(cwd)
    torch.autograd.set_detect_anomaly(False)
    return emb, net, mse


def tables_from_embedder(e, n_aug):
    """Read (theta, period) back out of the reference closures and evaluate cos/sin with torch,
    exactly as the lambda at models/embedder.py:127 does at call time."""
    cos_t = np.zeros((2, n_aug), np.float32)
    sin_t = np.zeros((2, n_aug), np.float32)
    per = np.zeros((2, n_aug), np.float32)
    for idx, fns in enumerate([e.embed_fns_x, e.embed_fns_y]):
        fns = fns[1:] if e.kwargs["include_input"] else fns
        for a in range(n_aug):
            p_fn, freq, theta = fns[2 * a].__defaults__
            cos_t[idx, a] = torch.cos(theta).item()
            sin_t[idx, a] = torch.sin(theta).item()
            per[idx, a] = freq.item()
    return cos_t, sin_t, per


def main():
    emb, net, mse = import_reference()
    freq_scales, freq_offsets, angle_offsets = [1], [0, -1, 1, 0.5, -0.5], [0]
    n_aug = 5

    # ---------------------------------------------------------------- encoding
    torch.manual_seed(0)
    np.random.seed(0)
    res = (211, 325)
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
    periods = torch.Tensor([[27.2, 24.9], [13.6, 12.45], [54.4, 49.8]])
    nerf, out_dim = emb.get_embedder(10, 0, res)
    assert out_dim == 21
    freqs = np.array([fn.__defaults__[1].item() for fn in nerf.embed_fns[1::2]], np.float32)
    periodic = [emb.get_embedder(10, 0, res, selected_angles=angles[i], selected_periods=periods[i],
                                 freq_scales=freq_scales, freq_offsets=freq_offsets,
                                 angle_offsets=angle_offsets)[0] for i in range(3)]
    n = 40
    coords = np.stack([np.random.randint(0, res[0], n), np.random.randint(0, res[1], n)], 1).astype(np.float32)
    coords[0] = (0, 0)
    coords[1] = (res[0] - 1, res[1] - 1)
    base = [e.embed(torch.from_numpy(coords).clone()) for e in periodic]
    full = torch.cat([nerf.embed(b) for b in base], 1)
    tabs = [tables_from_embedder(e, n_aug) for e in periodic]
    if not os.environ.get("NPP_GOLDEN_ONLY"):
        np.savez_compressed(
            os.path.join(HERE, "golden_encoding.npz"), coords=coords, res=np.array(res), freqs=freqs,
            angles=angles.numpy(), periods=periods.numpy(),
            cos_t=np.stack([t[0] for t in tabs]), sin_t=np.stack([t[1] for t in tabs]),
            period=np.stack([t[2] for t in tabs]),
            base=torch.cat(base, 1).numpy(), full=full.numpy())

    # ------------------------------------------------------------------- models
    for tag, topk in (("topk", 3), ("top1", 1), ("relu", 3)):
        only = os.environ.get("NPP_GOLDEN_ONLY")       # regenerate one fixture without touching the others
        if only and only != tag:
            continue
        activation = "relu" if tag == "relu" else "snake"
        torch.manual_seed(1)
        np.random.seed(1)
        W = 64
        if topk > 1:
            model = net.NPP_Net(D=8, W=W, freq_nerf=21, input_ch_periodic=22, input_ch_periodic_aux=22 * (topk - 1),
                                freq_scales=freq_scales, freq_offsets=freq_offsets, angle_offsets=angle_offsets,
                                output_ch=3, skips=[4], activation=activation)
        else:
            model = net.NPP_Net_top1(D=8, W=W, freq_nerf=21, input_ch_periodic=22, freq_scales=freq_scales,
                                     freq_offsets=freq_offsets, angle_offsets=angle_offsets, output_ch=3,
                                     skips=[4], activation=activation)
        enc = full[:, : 462 * topk].clone()
        target = torch.rand(n, 3)
        mask = (torch.rand(n, 1) > 0.3).float()
        out = {"enc": enc.numpy(), "target": target.numpy(), "mask": mask.numpy()}
        for k, v in model.state_dict().items():
            out["init/" + k] = v.numpy().copy()

        acts = {}
        hooks = []
        for name, mod in model.named_modules():
            if isinstance(mod, torch.nn.Linear):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, name=name: acts.__setitem__(name, o.detach().numpy().copy())))
        opt = torch.optim.Adam(params=list(model.parameters()), lr=5e-4, betas=(0.9, 0.999))
        global_step = 0
        losses = []
        for it in range(1, 4):
            raw = model(None, enc)
            pred = torch.sigmoid(raw)                                     # models/helpers.py:55-56
            opt.zero_grad()
            loss = mse.img2mse(pred, target, "l2", None, mask)            # models/mse_calculator.py:13
            loss.backward()
            if it == 1:
                for k, v in acts.items():
                    out["z/" + k] = v
                out["logits"] = raw.detach().numpy().copy()
                out["pred"] = pred.detach().numpy().copy()
                for k, prm in model.named_parameters():
                    if prm.grad is not None:
                        out["grad/" + k] = prm.grad.numpy().copy()
            opt.step()
            new_lr = 5e-4 * (0.1 ** (global_step / (500 * 100)))          # NPP_completion/train.py:258-263
            for g in opt.param_groups:
                g["lr"] = new_lr
            global_step += 1
            losses.append(loss.item())
        for h in hooks:
            h.remove()
        out["losses"] = np.array(losses, np.float32)
        for k, v in model.state_dict().items():
            out["final/" + k] = v.numpy().copy()
        np.savez_compressed(os.path.join(HERE, f"golden_{tag}.npz"), **out)
        print(tag, "losses", losses, "params", sum(p.numel() for p in model.parameters()))


if __name__ == "__main__":
    main()
