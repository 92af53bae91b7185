"""Adaptive robust pixel loss (reference models/mse_calculator.py:24-25 + externel_lib/robust_loss_pytorch).

CPU: the numpy oracle, the package's log Z table and the torch-op path against golden vectors generated from the live
reference (tests/golden/make_golden_robust.py).  GPU: the fused CUDA pass against the oracle, through the drop-in
``models.mse_calculator.img2mse`` surface."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import robust_oracle as R  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_robust.npz"))
CASES = ("init", "mid", "extreme")


def _case(name):
    m = G[f"{name}_mask"]
    return (G[f"{name}_x"], G[f"{name}_y"], None if m.shape[0] == 0 else m, G[f"{name}_latent_alpha"],
            G[f"{name}_latent_scale"])


def _close(a, b, rtol, atol):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    x, y, m, la, ls = _case(name)
    loss, gx, ga, gs = R.adaptive_img2mse(x, y, m, la, ls)
    alpha, scale, _, _ = R.adaptive_params(la, ls)
    _close(alpha, G[f"{name}_alpha"], 2e-7, 1e-7)
    _close(scale, G[f"{name}_scale"], 3e-7, 1e-9)
    _close(loss, G[f"{name}_loss"], 5e-7, 0)                       # reference runs in fp32
    _close(gx, G[f"{name}_gx"], 2e-5, 1e-6 * np.abs(G[f"{name}_gx"]).max())
    # the reference differentiates its spline of log Z, the oracle the integral itself: ~1e-4 relative
    _close(ga, G[f"{name}_g_latent_alpha"], 2e-3, 2e-6)
    _close(gs, G[f"{name}_g_latent_scale"], 1e-5, 1e-7)


def test_log_partition_table_matches_reference_spline():
    import npp_b200  # noqa: F401
    from npp_b200.robust_loss import log_partition
    al = torch.tensor(G["logz_alpha"], dtype=torch.float64)
    lz = log_partition(al).numpy()
    assert np.abs(lz - G["logz_ref"]).max() < 2e-6                 # the reference spline itself is "accurate to 1e-6"
    for a in (0.0059403, 0.42890172, 1.0, 1.79969853, 1.9940597):
        t = torch.tensor([a], dtype=torch.float64, requires_grad=True)
        v = log_partition(t)
        v.backward()
        ev, ed = R.log_partition(a)
        assert abs(v.item() - ev) < 2e-7 and abs(t.grad.item() - ed) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_torch_path_matches_reference_golden(name):
    import npp_b200  # noqa: F401
    from npp_b200.robust_loss import NppAdaptiveLoss
    x, y, m, la, ls = _case(name)
    ad = NppAdaptiveLoss(3, device="cpu")
    assert [tuple(p.shape) for p in ad.parameters()] == [(1, 3), (1, 3)]
    assert sorted(n for n, _ in ad.named_parameters()) == ["latent_alpha", "latent_scale"]
    with torch.no_grad():
        ad.latent_alpha.copy_(torch.tensor(la)[None])
        ad.latent_scale.copy_(torch.tensor(ls)[None])
    xt = torch.tensor(x, requires_grad=True)
    diff = xt - torch.tensor(y)
    if m is not None:
        mt = torch.tensor(m)
        diff = diff * mt + (1 - mt) * diff * 0.3
    loss = torch.mean(torch.mean(ad.lossfun(diff)))
    loss.backward()
    _close(loss.item(), G[f"{name}_loss"], 2e-6, 0)
    _close(xt.grad.numpy(), G[f"{name}_gx"], 1e-4, 1e-6 * np.abs(G[f"{name}_gx"]).max())
    _close(ad.latent_alpha.grad.numpy().ravel(), G[f"{name}_g_latent_alpha"], 5e-3, 5e-6)
    _close(ad.latent_scale.grad.numpy().ravel(), G[f"{name}_g_latent_scale"], 1e-3, 1e-6)


def test_default_initialisation_is_the_reference_one():
    import npp_b200  # noqa: F401
    from npp_b200.robust_loss import NppAdaptiveLoss
    ad = NppAdaptiveLoss(3, device="cpu")
    _close(ad.alpha().detach().numpy().ravel(), G["init_alpha"], 1e-6, 0)       # (lo + hi) / 2 = 1
    _close(ad.scale().detach().numpy().ravel(), G["init_scale"], 1e-6, 0)       # scale_init = 1
    with pytest.raises(ValueError):
        NppAdaptiveLoss(3, device="cpu", alpha_lo=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_cuda_matches_oracle(name):
    import npp_b200  # noqa: F401
    from npp_b200.robust_loss import NppAdaptiveLoss
    sys.path.insert(0, os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200"))
    from models.mse_calculator import img2mse
    x, y, m, la, ls = _case(name)
    ad = NppAdaptiveLoss(3, device="cuda")
    with torch.no_grad():
        ad.latent_alpha.copy_(torch.tensor(la)[None])
        ad.latent_scale.copy_(torch.tensor(ls)[None])
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    yt = torch.tensor(y, device="cuda")
    mt = None if m is None else torch.tensor(m, device="cuda")
    loss = img2mse(xt, yt, "robust_loss_adaptive", ad, mt) * 2.0     # upstream gradient 2 goes through backward
    loss.backward()
    eloss, egx, ega, egs = R.adaptive_img2mse(x, y, m, la, ls)
    _close(loss.item() / 2.0, eloss, 3e-6, 0)
    _close(xt.grad.cpu().numpy() / 2.0, egx, 2e-4, 2e-6 * np.abs(egx).max())
    _close(ad.latent_alpha.grad.cpu().numpy().ravel() / 2.0, ega, 5e-3, 5e-6)
    _close(ad.latent_scale.grad.cpu().numpy().ravel() / 2.0, egs, 1e-3, 1e-6)


@pytest.mark.gpu
def test_fused_cuda_full_size_is_additive():
    """2^18 rows (cfg 4): the loss and its gradients over the batch equal the mean of the two halves."""
    import npp_b200  # noqa: F401
    from npp_b200.robust_loss import NppAdaptiveLoss
    n = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(n, 3, device="cuda", generator=g)
    y = torch.rand(n, 3, device="cuda", generator=g)
    m = (torch.rand(n, 1, device="cuda", generator=g) > 0.5).float()
    ad = NppAdaptiveLoss(3, device="cuda")

    def run(lo, hi):
        ad.zero_grad()
        xs = x[lo:hi].clone().requires_grad_(True)
        loss = ad.fused_img2mse(xs, y[lo:hi], m[lo:hi])
        loss.backward()
        return loss.item(), ad.latent_alpha.grad.clone(), ad.latent_scale.grad.clone()

    lf, af, sf = run(0, n)
    l0, a0, s0 = run(0, n // 2)
    l1, a1, s1 = run(n // 2, n)
    assert abs(lf - 0.5 * (l0 + l1)) < 1e-5 * abs(lf)
    assert (af - 0.5 * (a0 + a1)).abs().max().item() < 1e-5 and (sf - 0.5 * (s0 + s1)).abs().max().item() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("n,use_mask", [(1, True), (1000, True), (16384, False)])
def test_fused_l2_matches_formula(n, use_mask):
    """img2mse(x, y, 'l2', None, mask) fused (npp_l2_fwd_bwd) against the reference formula evaluated in float64
    (models/mse_calculator.py:14-27): value and gradient, with an upstream factor through backward."""
    sys.path.insert(0, os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200"))
    from models.mse_calculator import img2mse
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand(n, 3, device="cuda", generator=g, requires_grad=True)
    y = torch.rand(n, 3, device="cuda", generator=g)
    m = (torch.rand(n, 1, device="cuda", generator=g) > 0.4).float() if use_mask else None
    loss = img2mse(x, y, "l2", None, m) * 3.0
    loss.backward()
    xd, yd = x.detach().double(), y.double()
    d = xd - yd
    if m is not None:
        d = d * m.double() + (1 - m.double()) * d * 0.3
    ref = (d ** 2).mean()
    w = torch.ones(n, 1, device="cuda", dtype=torch.float64) if m is None else m.double() + (1 - m.double()) * 0.3
    gref = 2 * d * w / (3 * n)
    assert abs(loss.item() / 3.0 - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-12
    assert (x.grad.double() / 3.0 - gref).abs().max().item() <= 1e-6 * gref.abs().max().item() + 1e-12
