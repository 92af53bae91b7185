"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical coordinates and weights.

Tolerance (BASELINE.json north_star): per-layer activations and gradients within 1e-3 relative with
fp32 accumulation.  "Relative" is the relative Frobenius error ||x - ref|| / ||ref|| of each tensor.
The tensor-core operands are fp16 (11-bit significand, same as TF32), accumulators, biases, snake,
loss, master weights and Adam are fp32."""
import numpy as np
import pytest
import torch

from oracle import npp_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3
RES = (512, 512)
ANGLES = [[83.0, 172.5], [90.0, 180.0], [41.3, 127.9]]
PERIODS = [[42.7, 38.4], [21.35, 19.2], [85.4, 76.8]]


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def make(topk, n, seed=0, max_rows=None):
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(seed)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, ANGLES[:topk], PERIODS[:topk], freqs)
    plan = Plan(enc, max_rows=max_rows or max(n, 128))
    params = O.init_params(rng, topk=topk)
    plan.load_state(params)
    coords = np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)
    tabs = [(enc.cos_t[j], enc.sin_t[j], enc.period[j]) for j in range(topk)]
    return plan, params, coords, tabs, freqs, rng


def test_encoding_matches_reference_golden():
    """The fp32 materialised encoding reproduces the reference's own Embedder outputs."""
    import os
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_encoding.npz"))
    enc = EncoderSpec(res=tuple(g["res"]), cos_t=g["cos_t"], sin_t=g["sin_t"], period=g["period"], freqs=g["freqs"])
    plan = Plan(enc, max_rows=128, training=False)
    out = plan.encode(torch.from_numpy(g["coords"]).cuda()).cpu().numpy()
    assert out.shape == g["full"].shape
    # base features (first 22 columns of each proposal) to a few ulp, expanded ones amplified by |f|<=22
    for j in range(3):
        np.testing.assert_allclose(out[:, 462 * j: 462 * j + 22], g["full"][:, 462 * j: 462 * j + 22], atol=2e-6)
    np.testing.assert_allclose(out, g["full"], atol=5e-5)
    # table built by the product's own host code equals the reference closures bit for bit
    enc2 = EncoderSpec.from_proposals(tuple(g["res"]), g["angles"], g["periods"], g["freqs"])
    np.testing.assert_array_equal(enc2.cos_t, g["cos_t"])
    np.testing.assert_array_equal(enc2.sin_t, g["sin_t"])
    np.testing.assert_array_equal(enc2.period, g["period"])


@pytest.mark.parametrize("topk,n", [(3, 1000), (1, 777), (3, 128), (3, 1), (1, 129)])
def test_forward_backward_parity(topk, n):
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, enc, topk_model=topk > 1)
    cd = torch.from_numpy(coords).cuda()
    logits = plan.forward(cd)
    torch.cuda.synchronize()

    # fp16 encoding that feeds the first GEMM (half rounding: 2^-11 relative)
    e1 = plan.debug("enc1", n).cpu().numpy()
    assert rel(e1[:, :462], enc[:, :462]) < 5e-4
    assert np.all(e1[:, 462:] == 0)
    report = {}
    for i, name in enumerate(plan.layer_names):
        h = plan.debug(f"h{i}", n).cpu().numpy()
        report[name] = rel(h, c["h"][name])
        assert report[name] < TOL, (name, report)
        if name in c["z"]:
            d = plan.debug(f"d{i}", n).cpu().numpy()
            assert rel(d, O.snake_grad(c["z"][name])) < TOL, name
    assert rel(logits.cpu().numpy(), logits_ref) < TOL, report

    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.3).astype(np.float32)
    g_ref = O.mse_l2_grad_logits(logits_ref, target, mask)
    grads_ref, deltas_ref = O.backward(params, c, g_ref, topk_model=topk > 1)
    plan.backward(n, torch.from_numpy(g_ref).cuda())
    torch.cuda.synchronize()
    scale = plan.grad_scale()
    assert scale > 1.0
    for i, name in enumerate(plan.layer_names):
        dl = plan.debug(f"delta{i}", n).cpu().numpy() / scale
        assert rel(dl, deltas_ref[name]) < TOL, ("delta", name)
    gv = plan.grad_views()
    assert sorted(gv.keys()) == sorted(grads_ref.keys())
    for k, ref in grads_ref.items():
        assert rel(gv[k].cpu().numpy(), ref) < TOL, ("grad", k)


def test_mse_kernel():
    plan, params, coords, tabs, freqs, rng = make(1, 333)
    logits = torch.randn(333, 3, device="cuda") * 2
    target = torch.rand(333, 3, device="cuda")
    mask = (torch.rand(333, 1, device="cuda") > 0.5).float()
    for m in (mask, None):
        loss, g, pred = plan.mse(logits, target, m, want_pred=True)
        mn = None if m is None else m.cpu().numpy()
        ref_pred = O.sigmoid(logits.cpu().numpy())
        np.testing.assert_allclose(pred.cpu().numpy(), ref_pred, atol=1e-6)
        assert abs(loss.item() - O.mse_l2(ref_pred, target.cpu().numpy(), mn)) < 1e-6
        assert rel(g.cpu().numpy(), O.mse_l2_grad_logits(logits.cpu().numpy(), target.cpu().numpy(), mn)) < 1e-5
    # global normalisation used under data parallelism
    loss2, g2, _ = plan.mse(logits, target, None, n_norm=666)
    loss1, g1, _ = plan.mse(logits, target, None)
    assert abs(loss2.item() * 2 - loss1.item()) < 1e-6
    assert rel(g2.cpu().numpy() * 2, g1.cpu().numpy()) < 1e-6


def test_adam_matches_torch_semantics():
    plan, params, coords, tabs, freqs, rng = make(3, 64)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    trained = set(plan.grad_views().keys())
    for step in range(1, 5):
        g = {k: (rng.standard_normal(p[k].shape) * 1e-3).astype(np.float32) for k in trained}
        for k, t in plan.grad_views().items():
            t.copy_(torch.from_numpy(g[k]))
        lr = O.lr_schedule(step)
        plan.adam_step(lr, step=step)
        O.adam_step(p, g, m, v, step, lr)
    got = plan.state()
    for k in p:
        np.testing.assert_allclose(got[k].cpu().numpy(), p[k], rtol=0, atol=2e-7, err_msg=k)
    # untrained tensors (alpha_linear) never move
    np.testing.assert_array_equal(got["alpha_linear.weight"].cpu().numpy(), params["alpha_linear.weight"])


@pytest.mark.parametrize("topk", [3, 1])
def test_train_steps_follow_oracle(topk):
    n = 2048
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    target = rng.random((n, 3), dtype=np.float32)
    mask = np.ones((n, 1), np.float32)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    cd, td, md = (torch.from_numpy(a).cuda() for a in (coords, target, mask))
    loss_d = torch.zeros((), device="cuda")
    for step in range(1, 9):
        lr = O.lr_schedule(step)
        plan.train_step(cd, td, md, lr, loss_d, step=step)
        ref_loss, _ = O.train_step(p, m, v, step, enc, target, mask, lr, topk_model=topk > 1)
        assert abs(loss_d.item() - ref_loss) < 1e-3 * ref_loss, (step, loss_d.item(), ref_loss)
    assert plan.launch_count() > 20
    # after 8 Adam steps of size ~5e-4 the weights agree to a fraction of one step
    got = plan.state()
    for k in plan.grad_views():
        assert np.abs(got[k].cpu().numpy() - p[k]).max() < 1e-3, k


def test_linearity_of_backward_in_grad():
    """Size-independent property: the weight gradient is linear in dL/dlogits, including the fp16 delta
    scaling (a power of two, so scaling g by 2^k must scale every gradient exactly by 2^k)."""
    n = 16384
    plan, params, coords, tabs, freqs, rng = make(3, n)
    cd = torch.from_numpy(coords).cuda()
    plan.forward(cd)
    g = torch.randn(n, 3, device="cuda") * 1e-5
    plan.backward(n, g)
    g1 = plan.grads.clone()
    plan.backward(n, g * 1024.0)
    g2 = plan.grads.clone()
    assert torch.equal(g1 * 1024.0, g2)
    assert torch.isfinite(g1).all() and g1.abs().max() > 0


def test_rejects_bad_input():
    import npp_b200
    plan, *_ = make(1, 64, max_rows=256)
    with pytest.raises(ValueError):
        plan.forward(torch.zeros(300, 2, device="cuda"))
    with pytest.raises(npp_b200._native.NppError):
        npp_b200._native.check(plan.lib.npp_forward(plan.handle, None, 10, None, None))
