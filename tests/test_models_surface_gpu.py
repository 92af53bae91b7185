"""The reference's train-loop body (NPP_completion/train.py:93-105,164-195,253-263) run on the drop-in ``models``
package and checked step by step against the CPU oracle started from the same weights."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

from oracle import npp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")

EXPECTED_KEYS_TOPK = sorted(
    [f"periodic_linears.{i}.{s}" for i in range(8) for s in ("weight", "bias")] +
    [f"{m}.{s}" for m in ("scale_linears.0", "pos_linears.0", "feature_linear1", "feature_linear2", "alpha_linear",
                          "rgb_linear") for s in ("weight", "bias")])


def _args(topk):
    return argparse.Namespace(multires=10, i_embed=0, p_topk=topk, freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5],
                              angle_offsets=[0], netdepth=8, netwidth=512, activation='snake', netchunk=1024 * 4096,
                              lrate=5e-4, lrate_decay=500, normalize_type=1, loss_type='l2')


@pytest.mark.parametrize("topk", [3, 1])
def test_train_loop_on_dropin_models(topk, monkeypatch):
    monkeypatch.delenv("NPP_B200_EMBED", raising=False)
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render
    from models.mse_calculator import img2mse

    torch.manual_seed(0)
    np.random.seed(0)
    dev = torch.device("cuda")
    res = (96, 128)
    args = _args(topk)
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])[:topk]
    periods = torch.Tensor([[17.2, 14.9], [8.6, 7.45], [34.4, 29.8]])[:topk]
    kw, _, start, grad_vars, optimizer, embedder, embedder_periodic = create_npp_net(args, angles, periods, res, None)
    model = kw['network_fn']
    assert sorted(model.state_dict().keys()) == (EXPECTED_KEYS_TOPK if topk > 1 else
                                                 [k for k in EXPECTED_KEYS_TOPK if not k.startswith("scale_linears")])
    assert sum(p.numel() for p in model.parameters()) == (3836932 if topk > 1 else 2970116)

    # oracle twin from the same initial weights and encoder constants
    p = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    tabs = [(e.cos_t, e.sin_t, e.period) for e in embedder_periodic]
    freqs = embedder.freqs

    # "table build" exactly as the script does it -- in coords mode it is just the coordinates
    img = torch.rand(1, res[0], res[1], 3, device=dev)
    h, w = torch.meshgrid(torch.arange(0, res[0]), torch.arange(0, res[1]), indexing="ij")
    i_all = torch.stack([h, w], dim=-1).reshape(-1, 2).float().to(dev)
    i_train = i_all[torch.randperm(i_all.shape[0], device=dev)[:6000]]
    embs = []
    for i in range(args.p_topk):
        embs.append(embedder.embed(embedder_periodic[i].embed(i_train.clone())))
    i_train_emb = torch.cat(embs, 1)
    assert i_train_emb.shape == (6000, 2)

    N_rand = 2048
    global_step = start
    for it in range(1, 7):
        select_inds = np.random.choice(i_train.shape[0], size=[N_rand], replace=False)
        select_coords = i_train[select_inds].long()
        gt_rgb = img[0, select_coords[:, 0], select_coords[:, 1], :]
        gt_mask = torch.ones_like(gt_rgb[:, :1])
        pred_rgb = render(None, i_train_emb[select_inds], args, **kw)
        optimizer.zero_grad()
        loss = img2mse(pred_rgb, gt_rgb, args.loss_type, None, gt_mask)
        loss.backward()
        optimizer.step()
        new_lrate = args.lrate * (0.1 ** (global_step / (args.lrate_decay * 100)))
        for g in optimizer.param_groups:
            g['lr'] = new_lrate
        global_step += 1
        enc = O.encode(select_coords.float().cpu().numpy(), tabs, freqs, res)
        ref_loss, _ = O.train_step(p, m, v, it, enc, gt_rgb.cpu().numpy(), gt_mask.cpu().numpy(), O.lr_schedule(it),
                                   topk_model=topk > 1)
        assert abs(loss.item() - ref_loss) < 1e-3 * ref_loss, (it, loss.item(), ref_loss)
    sd = model.state_dict()
    for k in p:
        assert np.abs(sd[k].cpu().numpy() - p[k]).max() < 1e-3, k

    # chunked no-grad inference like the i_testset branch (train.py:277-309)
    with torch.no_grad():
        out = torch.cat([render(None, i_all[j:j + 5000], args, **kw) for j in range(0, i_all.shape[0], 5000)])
    assert out.shape == (res[0] * res[1], 3) and torch.isfinite(out).all()


def test_table_mode_matches_coords_mode(monkeypatch):
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render
    torch.manual_seed(0)
    res = (64, 80)
    args = _args(3)
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
    periods = torch.Tensor([[17.2, 14.9], [8.6, 7.45], [34.4, 29.8]])
    kw, _, _, _, _, embedder, per = create_npp_net(args, angles, periods, res, None)
    coords = torch.stack([torch.randint(0, res[0], (500,)), torch.randint(0, res[1], (500,))], 1).float().cuda()
    monkeypatch.setenv("NPP_B200_EMBED", "coords")
    with torch.no_grad():
        a = render(None, torch.cat([embedder.embed(e.embed(coords.clone())) for e in per], 1), args, **kw)
    monkeypatch.setenv("NPP_B200_EMBED", "table")
    table = torch.cat([embedder.embed(e.embed(coords.clone())) for e in per], 1)
    assert table.shape == (500, 1386)
    tabs = [(e.cos_t, e.sin_t, e.period) for e in per]
    np.testing.assert_allclose(table.cpu().numpy(), O.encode(coords.cpu().numpy(), tabs, embedder.freqs, res), atol=5e-5)
    with torch.no_grad():
        b = render(None, table, args, **kw)
    assert (a - b).abs().max().item() < 2e-3


def test_foreign_parameters_and_external_weight_edit():
    """Adam over a mixed list (our arena + foreign tiny parameters, models/helpers.py:144-151) and shadow-weight
    refresh after someone edits the fp32 parameters in place."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render

    class Adaptive(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.latent = torch.nn.Parameter(torch.zeros(1, 3, device="cuda"))

    class Percep:
        adaptive_perceps = [Adaptive()]

    torch.manual_seed(0)
    args = _args(1)
    args.use_adaptive_perceptual_loss = True
    res = (64, 64)
    kw, _, _, grad_vars, opt, embedder, per = create_npp_net(args, torch.Tensor([[90.0, 180.0]]),
                                                            torch.Tensor([[12.0, 11.0]]), res, Percep())
    foreign = Percep.adaptive_perceps[0].latent
    assert any(g is foreign for g in grad_vars)
    coords = torch.stack([torch.randint(0, 64, (256,)), torch.randint(0, 64, (256,))], 1).float().cuda()
    pred = render(None, coords, args, **kw)
    loss = ((pred - 0.5) ** 2).mean() + (foreign ** 2).sum() + foreign.sum()
    opt.zero_grad()
    loss.backward()
    w_before = kw['network_fn'].rgb_linear.weight.detach().clone()
    opt.step()
    assert (foreign.detach().abs() > 0).all()                       # foreign parameter moved
    assert not torch.equal(w_before, kw['network_fn'].rgb_linear.weight.detach())
    # external in-place edit -> next forward must see it
    with torch.no_grad():
        out1 = render(None, coords, args, **kw)
        kw['network_fn'].rgb_linear.bias.add_(1.0)
        out2 = render(None, coords, args, **kw)
    assert (out2 > out1).all()
    # double forward then backward of the stale graph is rejected
    a = render(None, coords, args, **kw)
    _ = render(None, coords, args, **kw)
    with pytest.raises(RuntimeError):
        a.sum().backward()


def test_default_adaptive_robust_loss_iterations():
    """The reference's DEFAULT --loss_type (robust_loss_adaptive, options/arg_config.py:34) through the drop-in surface:
    `adaptive_pix` from models.helpers, `img2mse(pred, gt, 'robust_loss_adaptive', adaptive_pix, mask)` (train.py:205-208),
    its latent parameters in `grad_vars` (helpers.py:144) updated by the optimizer next to the network.  Checked per
    iteration against the numpy oracle (network oracle for the prediction, robust oracle for the loss)."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import models.helpers as H
    from models.mse_calculator import img2mse
    from oracle import robust_oracle as R

    torch.manual_seed(1)
    np.random.seed(1)
    res = (80, 96)
    args = _args(3)
    args.loss_type = 'robust_loss_adaptive'
    adaptive_pix = H.adaptive_pix
    assert adaptive_pix is not None and sorted(n for n, _ in adaptive_pix.named_parameters()) == ["latent_alpha", "latent_scale"]
    with torch.no_grad():   # module-level singleton (like the reference's): start from its initial state
        adaptive_pix.latent_alpha.zero_()
        adaptive_pix.latent_scale.zero_()
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
    periods = torch.Tensor([[17.2, 14.9], [8.6, 7.45], [34.4, 29.8]])
    kw, _, _, grad_vars, optimizer, embedder, per = H.create_npp_net(args, angles, periods, res, None)
    assert any(q is adaptive_pix.latent_alpha for q in grad_vars) and any(q is adaptive_pix.latent_scale for q in grad_vars)
    model = kw['network_fn']
    p = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    tabs = [(e.cos_t, e.sin_t, e.period) for e in per]
    n = 3000
    coords = torch.stack([torch.randint(0, res[0], (n,)), torch.randint(0, res[1], (n,))], 1).float().cuda()
    emb = torch.cat([embedder.embed(e.embed(coords.clone())) for e in per], 1)
    gt = torch.rand(n, 3, device="cuda")
    mask = (torch.rand(n, 1, device="cuda") > 0.2).float()

    # iteration 1: loss and latent gradients against the oracles, from the same weights
    pred = H.render(None, emb, args, **kw)
    optimizer.zero_grad()
    loss = img2mse(pred, gt, args.loss_type, adaptive_pix, mask)
    loss.backward()
    enc = O.encode(coords.cpu().numpy(), tabs, embedder.freqs, res)
    logits = O.forward(p, enc, topk_model=True)[0]
    pred_ref = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    eloss, _, ega, egs = R.adaptive_img2mse(pred_ref, gt.cpu().numpy(), mask.cpu().numpy(), np.zeros(3), np.zeros(3))
    assert abs(loss.item() - eloss) < 1e-3 * abs(eloss), (loss.item(), eloss)
    np.testing.assert_allclose(adaptive_pix.latent_alpha.grad.cpu().numpy().ravel(), ega, rtol=2e-2, atol=1e-5)
    np.testing.assert_allclose(adaptive_pix.latent_scale.grad.cpu().numpy().ravel(), egs, rtol=2e-2, atol=1e-5)
    optimizer.step()
    # a few more iterations: the NLL goes down and the foreign parameters move with Adam's step size
    first = loss.item()
    for _ in range(5):
        pred = H.render(None, emb, args, **kw)
        optimizer.zero_grad()
        loss = img2mse(pred, gt, args.loss_type, adaptive_pix, mask)
        loss.backward()
        optimizer.step()
    assert torch.isfinite(loss) and loss.item() < first
    assert 1e-3 < adaptive_pix.latent_scale.abs().max().item() < 6.5 * args.lrate


def test_small_parameter_adam_matches_torch():
    """Foreign parameters (adaptive_pix latents, adaptive LPIPS / style heads: models/helpers.py:144-151) are stepped by
    npp_adam_flat: same trajectory as torch.optim.Adam, including a rewritten lr and a step without a gradient."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.optim import NppAdam
    torch.manual_seed(3)
    shapes = [(1, 3), (1, 3), (5, 7), (1, 1)]
    ours = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = NppAdam(ours, lr=5e-4, betas=(0.9, 0.999), net=None)
    topt = torch.optim.Adam(ref, lr=5e-4, betas=(0.9, 0.999))
    for it in range(6):
        lr = 5e-4 * (0.9 ** it)
        for g in opt.param_groups:
            g['lr'] = lr
        for g in topt.param_groups:
            g['lr'] = lr
        for i, (a, b) in enumerate(zip(ours, ref)):
            if it == 2 and i == 1:
                a.grad = b.grad = None            # skipped by both, step counts stay aligned
                continue
            gr = torch.randn_like(a) * (10.0 ** (i - 1))
            a.grad, b.grad = gr.clone(), gr.clone()
        opt.step()
        topt.step()
        for a, b in zip(ours, ref):
            torch.testing.assert_close(a.detach(), b.detach(), rtol=2e-6, atol=1e-8)


def test_relu_activation_on_the_dropin_surface(monkeypatch):
    """args.activation='relu' (anything but 'snake', reference networks.py:51-54): NPP_Net builds, trains and has no
    `snakes` module, like the reference object."""
    monkeypatch.delenv("NPP_B200_EMBED", raising=False)
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render
    from models.mse_calculator import img2mse
    args = _args(3)
    args.activation = 'relu'
    torch.manual_seed(0)
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
    periods = torch.Tensor([[42.7, 38.4], [21.35, 19.2], [85.4, 76.8]])
    kw, _, _, grad_vars, optimizer, embedder, embedder_periodics = create_npp_net(args, angles, periods, (512, 512), None)
    model = kw["network_fn"]
    assert model.snakes is None
    coords = torch.stack([torch.randint(0, 512, (4096,)), torch.randint(0, 512, (4096,))], 1).float().cuda()
    target = (0.5 + 0.4 * torch.sin(coords[:, :1] * 0.147 + torch.arange(3, device="cuda"))).contiguous()
    enc = torch.cat([e.embed(coords.clone()) for e in embedder_periodics], 1)
    losses = []
    for _ in range(30):
        pred = render(None, enc, args, **kw)
        optimizer.zero_grad()
        loss = img2mse(pred, target, 'l2', None, None)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert losses[-1] < 0.8 * losses[0], losses[::6]


def test_backward_accumulates_without_zero_grad():
    """Two forward/backward passes without zero_grad() in between (micro-batch accumulation, foreign training loops):
    .grad must hold the SUM, as autograd's AccumulateGrad gives for the reference modules -- the kernels overwrite
    the gradient arena, so the drop-in layer has to add the earlier contents back."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render
    from models.mse_calculator import img2mse
    torch.manual_seed(0)
    res = (64, 80)
    args = _args(3)
    angles = torch.Tensor([[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]])
    periods = torch.Tensor([[17.2, 14.9], [8.6, 7.45], [34.4, 29.8]])
    kw, _, _, _, optimizer, embedder, per = create_npp_net(args, angles, periods, res, None)
    model = kw['network_fn']
    coords = [torch.stack([torch.randint(0, res[0], (700,)), torch.randint(0, res[1], (700,))], 1).float().cuda()
              for _ in range(2)]
    targets = [torch.rand(700, 3, device="cuda") for _ in range(2)]

    def one(k):
        x = torch.cat([embedder.embed(e.embed(coords[k].clone())) for e in per], 1)
        loss = img2mse(render(None, x, args, **kw), targets[k], 'l2', None, None)
        loss.backward()

    names = [n for n, p in model.named_parameters() if not n.startswith("alpha_linear")]
    singles = []
    for k in range(2):
        optimizer.zero_grad()
        one(k)
        singles.append({n: p.grad.clone() for n, p in model.named_parameters() if n in names})
    optimizer.zero_grad()
    one(0)
    one(1)                      # no zero_grad in between
    for n, p in model.named_parameters():
        if n not in names:
            continue
        want = singles[0][n] + singles[1][n]
        err = (p.grad - want).norm() / (want.norm() + 1e-30)
        assert err < 1e-4, (n, err.item())     # identical kernels, only atomics order differs


def test_unrunnable_skip_index_is_refused():
    """skips[0] == D-1 concatenates after the last trunk layer in the reference forward, which then fails in
    feature_linear1 (networks.py:70-73); the drop-in refuses it instead of training another network."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.embedder import get_embedder
    from models.networks import NPP_Net_top1
    res = (32, 32)
    emb, freq_nerf = get_embedder(10, 0, res)
    e, ch = get_embedder(10, 0, res, selected_angles=torch.Tensor([83.0, 172.5]), selected_periods=torch.Tensor([9.0, 7.0]),
                         freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0])
    with pytest.raises(ValueError):
        NPP_Net_top1(D=4, W=256, freq_nerf=freq_nerf, input_ch_periodic=ch, freq_scales=[1],
                     freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0], output_ch=3, skips=[3], activation='snake')
    NPP_Net_top1(D=4, W=256, freq_nerf=freq_nerf, input_ch_periodic=ch, freq_scales=[1],
                 freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0], output_ch=3, skips=[4], activation='snake')
