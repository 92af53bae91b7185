"""Golden vectors for the search-stage network (NPP_Net_light) and the search-mode encoders.

Run in the build container (reference mounted at /root/reference; not available on the GPU box):

    python tests/golden/make_golden_light.py

Imports the reference's models/embedder.py, models/networks.py and models/mse_calculator.py, builds
the encoders the way create_npp_net(is_search=True) does (models/helpers.py:87-103) and runs three
iterations of the loop body of NPP_proposal/search.py:112-146 (img2mse 'l2', Adam, LR rewrite) on
seeded inputs.  W=64 keeps the fixture small; D=4 and skips=[4] are the search defaults
(options/arg_config.py:114-116), so no skip connection is active.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, tables_from_embedder  # noqa: E402


def main():
    emb, net, mse = import_reference()
    freq_scales, freq_offsets, angle_offsets = [1], [0, -1, 1, 0.5, -0.5], [0]
    torch.manual_seed(2)
    np.random.seed(2)
    res = (211, 325)
    angles = torch.Tensor([83.0, 172.5])
    periods = torch.Tensor([27.2, 24.9])
    embedder, ch_pos = emb.get_embedder(10, 0, res, is_search=True)
    periodic, ch_per = emb.get_embedder(10, 0, res, selected_angles=angles, selected_periods=periods,
                                        freq_scales=freq_scales, freq_offsets=freq_offsets,
                                        angle_offsets=angle_offsets, is_search=True)
    assert (ch_pos, ch_per) == (42, 20)
    freqs = np.array([fn.__defaults__[1].item() for fn in embedder.embed_fns[1::2]], np.float32)
    n = 48
    coords = np.stack([np.random.randint(0, res[0], n), np.random.randint(0, res[1], n)], 1).astype(np.float32)
    coords[0] = (0, 0)
    coords[1] = (res[0] - 1, res[1] - 1)
    pos = embedder.embed(torch.from_numpy(coords).clone())          # search.py:104 (normalises the clone in place)
    per = periodic.embed(torch.from_numpy(coords))                  # search.py:107
    cos_t, sin_t, period = tables_from_embedder(periodic, 5)

    W = 64
    model = net.NPP_Net_light(D=4, W=W, input_ch=ch_pos, input_ch_periodic=ch_per, freq_scales=freq_scales,
                              freq_offsets=freq_offsets, angle_offsets=angle_offsets, output_ch=3, skips=[4],
                              activation="snake")
    target = torch.rand(n, 3)
    out = {"coords": coords, "res": np.array(res), "freqs": freqs, "angles": angles.numpy(),
           "periods": periods.numpy(), "cos_t": cos_t, "sin_t": sin_t, "period": period,
           "pos": pos.numpy(), "per": per.numpy(), "target": target.numpy()}
    for k, v in model.state_dict().items():
        out["init/" + k] = v.numpy().copy()
    acts, hooks = {}, []
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Linear):
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, name=name: acts.__setitem__(name, o.detach().numpy().copy())))
    opt = torch.optim.Adam(params=list(model.parameters()), lr=5e-4, betas=(0.9, 0.999))
    global_step, losses = 0, []
    for it in range(1, 4):
        raw = model(pos, per)
        pred = torch.sigmoid(raw)                                    # models/helpers.py:55-56
        opt.zero_grad()
        loss = mse.img2mse(pred, target, "l2", None, None)           # search.py:134 with --loss_type l2
        loss.backward()
        if it == 1:
            for k, v in acts.items():
                out["z/" + k] = v
            out["logits"] = raw.detach().numpy().copy()
            for k, prm in model.named_parameters():
                if prm.grad is not None:
                    out["grad/" + k] = prm.grad.numpy().copy()
        opt.step()
        new_lr = 5e-4 * (0.1 ** (global_step / (500 * 100)))         # search.py:139-144
        for g in opt.param_groups:
            g["lr"] = new_lr
        global_step += 1
        losses.append(loss.item())
    for h in hooks:
        h.remove()
    out["losses"] = np.array(losses, np.float32)
    for k, v in model.state_dict().items():
        out["final/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "golden_light.npz"), **out)
    print("light losses", losses, "trained grads", sorted(k for k in out if k.startswith("grad/")))


if __name__ == "__main__":
    main()
