"""Search-stage fits: NPP_Net_light and the search-mode encoders on the CUDA path,
through the C ABI, against the CPU oracle (oracle/npp_oracle.py *_light / encode_search, pinned to the reference by
tests/golden/golden_light.npz) on identical coordinates and weights.

Tolerances as in test_parity_gpu.py: fp16 tensor-core operands with fp32 accumulation, per-layer 1e-3 relative
(Frobenius), 2e-3 for quantities accumulated over the whole chain."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import npp_oracle as O

pytestmark = pytest.mark.gpu

TOL, TOL_CUM = 1e-3, 2e-3
RES = (211, 325)
ANGLES, PERIODS = [83.0, 172.5], [27.2, 24.9]
G = os.path.join(os.path.dirname(__file__), "golden")
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def make(n, seed=0, width=256, depth=4, max_rows=None, training=True):
    import npp_b200  # noqa: F401
    from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT
    rng = np.random.default_rng(seed)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, [ANGLES], [PERIODS], freqs, include_input=False)
    plan = Plan(enc, depth=depth, width=width, skip_layer=-1, max_rows=max_rows or max(n, 128), model=MODEL_LIGHT,
                training=training)
    params = O.init_params_light(rng, width=width, depth=depth)
    plan.load_state(params)
    coords = np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)
    table = (enc.cos_t[0], enc.sin_t[0], enc.period[0])
    return plan, params, coords, table, freqs, rng


def test_search_encoders_match_reference_golden():
    """npp_encode in search mode reproduces the reference's two search-mode embedders ([periodic | positional])."""
    from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT
    g = np.load(os.path.join(G, "golden_light.npz"))
    enc = EncoderSpec(res=tuple(g["res"]), cos_t=g["cos_t"][None], sin_t=g["sin_t"][None], period=g["period"][None],
                      freqs=g["freqs"], include_input=False)
    plan = Plan(enc, depth=4, width=256, skip_layer=-1, max_rows=128, model=MODEL_LIGHT, training=False)
    assert plan.encoding_width == 62
    out = plan.encode(torch.from_numpy(g["coords"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(out[:, :20], g["per"], atol=2e-6)
    np.testing.assert_allclose(out[:, 20:], g["pos"], atol=1e-5)     # sin/cos of arguments up to ~25
    enc2 = EncoderSpec.from_proposals(tuple(g["res"]), [g["angles"]], [g["periods"]], g["freqs"], include_input=False)
    np.testing.assert_array_equal(enc2.cos_t[0], g["cos_t"])
    np.testing.assert_array_equal(enc2.sin_t[0], g["sin_t"])
    np.testing.assert_array_equal(enc2.period[0], g["period"])


def test_state_dict_layout_is_the_reference_one():
    """Same tensor names and shapes as NPP_Net_light(D=4, W=256, input_ch=42, input_ch_periodic=20).state_dict()."""
    plan, params, *_ = make(8)
    ours = {s.name: tuple(s.shape) for s in plan.slots}
    assert ours == {k: tuple(v.shape) for k, v in params.items()}
    g = np.load(os.path.join(G, "golden_light.npz"))
    assert sorted(ours) == sorted(k[5:] for k in g.files if k.startswith("init/"))
    trained = sorted(s.name for s in plan.slots if s.trained)
    assert trained == sorted(k[5:] for k in g.files if k.startswith("grad/"))


@pytest.mark.parametrize("n,width", [(2048, 256), (777, 256), (1, 256), (300, 512)])
def test_light_forward_backward_parity(n, width):
    plan, params, coords, table, freqs, rng = make(n, seed=n, width=width)
    pos, per = O.encode_search(coords, table, freqs, RES)
    logits_ref, c = O.forward_light(params, pos, per)
    cd = torch.from_numpy(coords).cuda()
    logits = plan.forward(cd)
    torch.cuda.synchronize()
    W = width
    idx = {name: i for i, name in enumerate(plan.layer_names)}
    assert list(idx) == [f"periodic_linears.{i}" for i in range(4)] + ["feature_linear1", "pos_linears.0"]
    buf = {}

    def ours(name, cols=None):
        if name not in buf:
            buf[name] = plan.debug(name, n).cpu().numpy()
        return buf[name] if cols is None else buf[name][:, :cols]

    e1, ep = ours("enc1"), ours("enc_aux")
    assert rel(e1[:, :20], per) < 5e-4 and np.all(e1[:, 20:] == 0)
    assert rel(ep[:, :42], pos) < 5e-4 and np.all(ep[:, 42:] == 0)
    out_w = {name: (W // 2 if name == "pos_linears.0" else W) for name in idx}

    def our_input(name):
        if name == "periodic_linears.0":
            return e1[:, :20]
        if name.startswith("periodic_linears."):
            return ours(f"h{int(name.split('.')[1]) - 1}")
        if name == "feature_linear1":
            return ours("h3")
        return np.concatenate([ours(f"h{idx['feature_linear1']}"), ep[:, :42]], 1)

    rep = {"fwd_layer": {}, "fwd_cum": {}, "bwd_layer": {}, "bwd_cum": {}, "grad_layer": {}, "grad_cum": {}}
    for name, i in idx.items():
        z = (our_input(name) @ params[name + ".weight"].T + params[name + ".bias"]).astype(np.float32)
        snake_layer = name in c["z"]
        h = ours(f"h{i}", out_w[name])
        rep["fwd_layer"][name] = rel(h, O.snake(z) if snake_layer else z)
        rep["fwd_cum"][name] = rel(h, c["h"][name])
        if snake_layer:
            rep["fwd_layer"][name + "/snake_grad"] = rel(ours(f"d{i}", out_w[name]), O.snake_grad(z))
    hp = ours(f"h{idx['pos_linears.0']}", W // 2)
    lg = logits.cpu().numpy()
    rep["fwd_layer"]["rgb_linear"] = rel(lg, hp @ params["rgb_linear.weight"].T + params["rgb_linear.bias"])
    rep["fwd_cum"]["rgb_linear"] = rel(lg, logits_ref)
    # the materialised-encoding entry point runs the same network
    enc62 = torch.from_numpy(np.concatenate([per, pos], 1)).cuda()
    assert rel(plan.forward_encoded(enc62).cpu().numpy(), logits_ref) < TOL
    plan.forward(cd)      # backward below belongs to the coordinate forward

    target = rng.random((n, 3), dtype=np.float32)
    g_ref = O.mse_l2_grad_logits(logits_ref, target, None)
    grads_ref, deltas_ref = O.backward_light(params, c, g_ref)
    plan.backward(n, torch.from_numpy(g_ref).cuda())
    torch.cuda.synchronize()
    scale = plan.grad_scale()
    buf.clear()
    dl = {name: ours(f"delta{i}", out_w[name]) / scale for name, i in idx.items()}
    buf.clear()
    gv = {k: v.cpu().numpy() for k, v in plan.grad_views().items()}
    assert sorted(gv) == sorted(grads_ref)

    def dact(name):
        return ours(f"d{idx[name]}", out_w[name]) if name in c["z"] else 1.0

    iso = {"pos_linears.0": (g_ref @ params["rgb_linear.weight"]) * dact("pos_linears.0")}
    iso["feature_linear1"] = dl["pos_linears.0"] @ params["pos_linears.0.weight"][:, :W]
    iso["periodic_linears.3"] = (dl["feature_linear1"] @ params["feature_linear1.weight"]) * dact("periodic_linears.3")
    for i in range(2, -1, -1):
        iso[f"periodic_linears.{i}"] = (dl[f"periodic_linears.{i + 1}"] @ params[f"periodic_linears.{i + 1}.weight"]) * \
            dact(f"periodic_linears.{i}")
    for name in idx:
        rep["bwd_layer"][name] = rel(dl[name], iso[name])
        rep["bwd_cum"][name] = rel(dl[name], deltas_ref[name])
        rep["grad_layer"][name + ".weight"] = rel(gv[name + ".weight"], dl[name].T.astype(np.float64) @ our_input(name))
        rep["grad_layer"][name + ".bias"] = rel(gv[name + ".bias"], dl[name].sum(0, dtype=np.float64))
    rep["grad_layer"]["rgb_linear.weight"] = rel(gv["rgb_linear.weight"], g_ref.T.astype(np.float64) @ hp)
    rep["grad_layer"]["rgb_linear.bias"] = rel(gv["rgb_linear.bias"], g_ref.sum(0, dtype=np.float64))
    for k, ref in grads_ref.items():
        rep["grad_cum"][k] = rel(gv[k], ref)
    d = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(d):
        import json
        with open(os.path.join(d, f"parity_light_w{width}_n{n}.json"), "w") as fh:
            json.dump(rep, fh, indent=1, sort_keys=True)
    # the zero-padded half of the 128-unit pos_linears.0 tile never leaks into a gradient
    if width == 256:
        full = plan.debug(f"delta{idx['pos_linears.0']}", n).cpu().numpy()
        assert np.all(full[:, 128:] == 0)
    for sect in ("fwd_layer", "bwd_layer", "grad_layer"):
        for k, v in rep[sect].items():
            assert v < TOL, (sect, k, v)
    for sect in ("fwd_cum", "bwd_cum", "grad_cum"):
        for k, v in rep[sect].items():
            assert v < TOL_CUM, (sect, k, v)
    assert rep["fwd_cum"]["rgb_linear"] < TOL


def test_light_fused_train_steps_follow_the_oracle():
    """npp_train_step on the search-stage network: N_rand = 2048 rows, no mask (NPP_proposal/search.py:112-146)."""
    n = 2048
    plan, params, coords, table, freqs, rng = make(n, seed=5)
    pos, per = O.encode_search(coords, table, freqs, RES)
    target = rng.random((n, 3), dtype=np.float32)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    cd, td = torch.from_numpy(coords).cuda(), torch.from_numpy(target).cuda()
    loss = torch.zeros((), device="cuda")
    for step in range(1, 7):
        lr = O.lr_schedule(step)
        plan.train_step(cd, td, None, lr, loss, step=step)
        l_ref, _ = O.train_step_light(p, m, v, step, pos, per, target, lr)
        assert abs(loss.item() - l_ref) < 1e-3 * l_ref, (step, loss.item(), l_ref)
    assert plan.launch_count() == 4      # encode, fused chain (forward + head + loss + backward), wgrad, update
    got = plan.state()
    for k in p:
        a = got[k].cpu().numpy()
        # six Adam steps move every weight by <= 6 lr; sign flips of near-zero gradients bound the difference
        assert np.abs(a - p[k]).max() < 7 * 5e-4, k
        if p[k].size >= 1024:
            assert np.abs(a - p[k]).mean() < 5e-5, k
    for k in ("scale_linears.0.weight", "feature_linear2.weight", "alpha_linear.weight"):
        np.testing.assert_array_equal(got[k].cpu().numpy(), params[k])     # never trained (networks.py:236-250)


def test_create_npp_net_search_mode_runs_the_search_loop_body(monkeypatch):
    """create_npp_net(is_search=True) + the loop body of NPP_proposal/search.py:112-146 on the drop-in `models` surface."""
    monkeypatch.delenv("NPP_B200_EMBED", raising=False)
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net, render
    from models.mse_calculator import img2mse
    args = types.SimpleNamespace(multires=10, i_embed=0, freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5],
                                 angle_offsets=[0], netdepth=4, netwidth=256, activation='snake', lrate=5e-4,
                                 netchunk=1024 * 64, normalize_type=1, p_topk=1)
    torch.manual_seed(0)
    np.random.seed(0)
    H, W_ = RES
    yy, xx = np.meshgrid(np.arange(H), np.arange(W_), indexing="ij")
    img = np.stack([0.5 + 0.4 * np.sin(2 * np.pi * xx / 27.2), 0.5 + 0.4 * np.cos(2 * np.pi * yy / 24.9),
                    0.5 + 0.3 * np.sin(2 * np.pi * (xx + yy) / 27.2)], -1).astype(np.float32)
    masked_img = torch.from_numpy(img)[None].cuda()
    i_train = torch.Tensor(np.stack([yy.reshape(-1), xx.reshape(-1)], 1))
    kw_train, kw_test, start, grad_vars, optimizer, embedder, embedder_periodic = create_npp_net(
        args, torch.Tensor(ANGLES), torch.Tensor(PERIODS), RES, percep_net=None, is_search=True)
    model = kw_train["network_fn"]
    assert type(model).__name__ == "NPP_Net_light" and (embedder.out_dim, embedder_periodic.out_dim) == (42, 20)
    sd = model.state_dict()
    assert sd["pos_linears.0.weight"].shape == (128, 298) and sd["periodic_linears.0.weight"].shape == (256, 20)
    i_train_emb = embedder.embed(i_train.clone())
    i_train_emb_periodic = embedder_periodic.embed(i_train)
    losses, global_step = [], 0
    for i in range(1, 41):
        select_inds = np.random.choice(i_train.shape[0], size=[2048], replace=False)
        select_coords = i_train[select_inds].long()
        gt_rgb = masked_img[0, select_coords[:, 0], select_coords[:, 1], :]
        pred_rgb = render(i_train_emb[select_inds], i_train_emb_periodic[select_inds], args, **kw_train)
        optimizer.zero_grad()
        loss = img2mse(pred_rgb, gt_rgb, 'l2', None, None)
        loss.backward()
        optimizer.step()
        new_lrate = args.lrate * (0.1 ** (global_step / (500 * 100)))
        for group in optimizer.param_groups:
            group['lr'] = new_lrate
        global_step += 1
        losses.append(loss.item())
    assert losses[-1] < 0.5 * losses[0], losses[::8]
    with torch.no_grad():      # validation render of search.py:152-170, 20000-row chunks
        out = render(i_train_emb[:20000], i_train_emb_periodic[:20000], args, **kw_train)
    assert out.shape == (20000, 3) and torch.isfinite(out).all()


def _fit_data(iters, n, seed=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    coords = torch.stack([torch.randint(0, RES[0], (iters, n), device="cuda", generator=g),
                          torch.randint(0, RES[1], (iters, n), device="cuda", generator=g)], -1).float().contiguous()
    target = (0.5 + 0.4 * torch.sin(coords[..., :1] * 0.23 + coords[..., 1:] * 0.11 +
                                    torch.arange(3, device="cuda") * 2.0)).contiguous()
    return coords, target


def test_fit_run_equals_stepwise_train_steps():
    """npp_fit_run is the loop of train steps with the scripts' learning-rate rewrite (search.py:139-144)."""
    iters, n = 12, 2048
    coords, target = _fit_data(iters, n)
    a, params, *_ = make(n, seed=11)
    b, _, *_ = make(n, seed=11)
    loss = torch.zeros((), device="cuda")
    step_losses = []
    for i in range(iters):
        a.train_step(coords[i], target[i], None, O.lr_schedule(i + 1, lrate=5e-4, lrate_decay=0.02), loss)
        step_losses.append(loss.item())
    run_losses = b.fit_run(coords, target, lrate=5e-4, lrate_decay=0.02)     # decays 10x every 2 steps: LR rule visible
    assert b.launch_count() == 4 * iters and b.adam_steps == iters
    np.testing.assert_allclose(run_losses.cpu().numpy(), step_losses, rtol=1e-4)
    sa, sb = a.state(), b.state()
    for k in sa:
        assert (sa[k] - sb[k]).abs().max().item() < 2e-4, k
    assert step_losses[-1] < step_losses[0]


def test_concurrent_candidate_fits_match_fits_run_alone():
    """search_fits.run_fits: one plan / stream / host thread per candidate periodicity, same batches for all
    (search.py:91-92 re-seeds per candidate), against the same fits run one after the other."""
    from npp_b200.search_fits import run_fits, gather_batches
    from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT
    iters, n, K = 60, 2048, 5
    rng = np.random.default_rng(0)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    H, W_ = RES
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W_, device="cuda"), indexing="ij")
    image = torch.stack([0.5 + 0.4 * torch.sin(2 * np.pi * xx / 27.2), 0.5 + 0.4 * torch.cos(2 * np.pi * yy / 24.9),
                         0.5 + 0.3 * torch.sin(2 * np.pi * (xx + yy) / 27.2)], -1).float()
    train_coords = torch.stack([yy.reshape(-1), xx.reshape(-1)], 1)
    idx = torch.from_numpy(np.stack([rng.choice(H * W_, n, replace=False) for _ in range(iters)]))
    coords_all, target_all = gather_batches(image, train_coords, idx)
    assert coords_all.shape == (iters, n, 2) and target_all.shape == (iters, n, 3)
    assert torch.equal(target_all[3, 5], image[int(coords_all[3, 5, 0]), int(coords_all[3, 5, 1])])

    def plans():
        out = []
        for k in range(K):
            enc = EncoderSpec.from_proposals(RES, [ANGLES], [[27.2 + 4 * k, 24.9 - 3 * k]], freqs, include_input=False)
            p = Plan(enc, depth=4, width=256, skip_layer=-1, max_rows=n, model=MODEL_LIGHT)
            p.reset_parameters(seed=0)       # same initial weights for every candidate, like torch.manual_seed(0)
            out.append(p)
        return out

    together = run_fits(plans(), coords_all, target_all).cpu().numpy()            # default: one re-launched graph
    threaded = run_fits(plans(), coords_all, target_all, grouped=False).cpu().numpy()   # stream + host thread per fit
    alone = np.stack([p.fit_run(coords_all, target_all).cpu().numpy() for p in plans()])
    assert together.shape == (K, iters)
    np.testing.assert_allclose(threaded[:, :20], alone[:, :20], rtol=2e-3)
    np.testing.assert_allclose(threaded, alone, rtol=2e-2)
    # a second grouped run on the same plans continues their Adam steps (the first graph is retired)
    ps = plans()
    first = run_fits(ps, coords_all[:10], target_all[:10]).cpu().numpy()
    second = run_fits(ps, coords_all[10:20], target_all[10:20]).cpu().numpy()
    np.testing.assert_allclose(np.concatenate([first, second], 1), alone[:, :20], rtol=2e-3)
    assert all(p.adam_steps == 20 for p in ps)
    # same trajectory; late iterations only to 2 % (atomic-ordering noise in the bias gradients, amplified by Adam)
    np.testing.assert_allclose(together[:, :20], alone[:, :20], rtol=2e-3)
    np.testing.assert_allclose(together, alone, rtol=2e-2)
    # the true periodicity (candidate 0) is the one that fits best -- what the search ranks by
    final = together[:, -10:].mean(1)
    assert final.argmin() == 0, final
    assert (together[:, -1] < together[:, 0]).all()


@pytest.mark.parametrize("mode", ["1", "2"])
def test_fit_run_graph_modes_equal_plain_launches(monkeypatch, mode):
    """NPP_FIT_GRAPH: 1 = the whole run as one CUDA graph, 2 = one captured step re-launched with its per-step scalars
    in device memory, 0 = plain launches (default).  Same fit in every mode."""
    iters, n = 10, 1024
    coords, target = _fit_data(iters, n, seed=9)
    a, *_ = make(n, seed=21)
    b, *_ = make(n, seed=21)
    monkeypatch.setenv("NPP_FIT_GRAPH", "0")
    direct = a.fit_run(coords, target).cpu().numpy()
    assert a.launch_count() == 4 * iters
    monkeypatch.setenv("NPP_FIT_GRAPH", mode)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        graphed = b.fit_run(coords, target)
        again = b.fit_run(coords, target)          # second run: the first graph is retired, Adam steps continue
    torch.cuda.current_stream().wait_stream(s)
    np.testing.assert_allclose(graphed.cpu().numpy(), direct, rtol=1e-3)
    assert b.adam_steps == 2 * iters and again[-1].item() < graphed[0].item()


def test_search_mode_table_embedding_matches_coords_mode(monkeypatch):
    """NPP_B200_EMBED=table: the search-mode embedders materialise the reference's [N,42] / [N,20] encodings (checked
    against the oracle) and NPP_Net_light.forward(x, x_periodic) on them equals the coordinate path."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from models.helpers import create_npp_net
    args = types.SimpleNamespace(multires=10, i_embed=0, freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5],
                                 angle_offsets=[0], netdepth=4, netwidth=256, activation='snake', lrate=5e-4,
                                 netchunk=1024 * 64, normalize_type=1, p_topk=1)
    torch.manual_seed(4)
    monkeypatch.delenv("NPP_B200_EMBED", raising=False)
    kw, _, _, _, _, embedder, embedder_periodic = create_npp_net(args, torch.Tensor(ANGLES), torch.Tensor(PERIODS), RES,
                                                                 percep_net=None, is_search=True)
    model = kw["network_fn"]
    rng = np.random.default_rng(1)
    coords = torch.from_numpy(np.stack([rng.integers(0, RES[0], 500), rng.integers(0, RES[1], 500)], 1).astype(np.float32))
    with torch.no_grad():
        ref = model(embedder.embed(coords.clone()), embedder_periodic.embed(coords)).cpu().numpy()
    monkeypatch.setenv("NPP_B200_EMBED", "table")
    pos = embedder.embed(coords.clone().cuda())
    per = embedder_periodic.embed(coords.cuda())
    assert pos.shape == (500, 42) and per.shape == (500, 20)
    table = (embedder_periodic.cos_t, embedder_periodic.sin_t, embedder_periodic.period)
    pos_ref, per_ref = O.encode_search(coords.numpy(), table, embedder.freqs, RES)
    np.testing.assert_allclose(per.cpu().numpy(), per_ref, atol=2e-6)
    np.testing.assert_allclose(pos.cpu().numpy(), pos_ref, atol=2e-5)       # torch sin/cos on the GPU vs numpy
    with torch.no_grad():
        out = model(pos, per).cpu().numpy()
    assert rel(out, ref) < 1e-3
