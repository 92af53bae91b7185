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


def _report(tag, data):
    """Keep the measured errors (copied into profiles/ by hand after a GPU run)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"parity_{tag}.json"), "w") as fh:
            json.dump(data, fh, indent=1, sort_keys=True)


def _segments(plan, name, c):
    """Oracle-side input of a layer, split the way the kernels see it."""
    return c["a"][name]


@pytest.mark.parametrize("topk,n", [(3, 1000), (1, 777), (3, 128), (3, 1), (1, 129), (2, 300)])
def test_forward_backward_parity(topk, n):
    """Two checks per layer, forward and backward:
      per-layer : the layer applied to exactly the inputs the kernel consumed (our fp16 buffers) with the
                  fp32 master weights, in fp32 on the CPU  -> TOL = 1e-3 (north_star tolerance)
      cumulative: against the all-fp32 oracle run from the coordinates, i.e. the error accumulated over up
                  to 13 chained layers of fp16-operand GEMMs -> TOL_CUM = 2e-3, measured values reported."""
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, enc, topk_model=topk > 1)
    cd = torch.from_numpy(coords).cuda()
    logits = plan.forward(cd)
    torch.cuda.synchronize()
    W = 512
    buf = {}

    def ours(name):
        if name not in buf:
            buf[name] = plan.debug(name, n).cpu().numpy()
        return buf[name]

    # fp16 encoding that feeds the first GEMM (half rounding only)
    e1 = ours("enc1")
    assert rel(e1[:, :462], enc[:, :462]) < 5e-4
    assert np.all(e1[:, 462:] == 0)
    idx = {name: i for i, name in enumerate(plan.layer_names)}

    def our_input(name):
        """Concatenated input of `name` in reference column order, built from our own buffers."""
        if name == "periodic_linears.0":
            return e1[:, :462]
        if name.startswith("periodic_linears."):
            i = int(name.split(".")[1])
            h = ours(f"h{i - 1}")
            return np.concatenate([e1[:, :462], h], 1) if i - 1 == 4 else h
        if name == "feature_linear1":
            return ours(f"h{idx['periodic_linears.7']}")
        if name == "scale_linears.0":
            return np.concatenate([ours(f"h{idx['feature_linear1']}"), ours("enc_aux")[:, :462 * (topk - 1)]], 1)
        if name == "feature_linear2":
            return ours(f"h{idx['scale_linears.0']}")
        if name == "pos_linears.0":
            f1 = ours(f"h{idx['feature_linear1']}")
            return np.concatenate([f1, ours(f"h{idx['feature_linear2']}")], 1) if topk > 1 else f1
        raise KeyError(name)

    rep = {"fwd_layer": {}, "fwd_cum": {}, "bwd_layer": {}, "bwd_cum": {}, "grad_layer": {}, "grad_cum": {}}
    for name, i in idx.items():
        a_in = our_input(name)
        z = (a_in @ params[name + ".weight"].T + params[name + ".bias"]).astype(np.float32)
        snake_layer = name in c["z"]
        h_iso = O.snake(z) if snake_layer else z
        h = ours(f"h{i}")
        rep["fwd_layer"][name] = rel(h, h_iso)
        rep["fwd_cum"][name] = rel(h, c["h"][name])
        if snake_layer:
            d = ours(f"d{i}")
            rep["fwd_layer"][name + "/snake_grad"] = rel(d, O.snake_grad(z))
            rep["fwd_cum"][name + "/snake_grad"] = rel(d, O.snake_grad(c["z"][name]))
    hp = ours(f"h{idx['pos_linears.0']}")
    lg = logits.cpu().numpy()
    rep["fwd_layer"]["rgb_linear"] = rel(lg, hp @ params["rgb_linear.weight"].T + params["rgb_linear.bias"])
    rep["fwd_cum"]["rgb_linear"] = rel(lg, logits_ref)

    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.3).astype(np.float32)
    g_ref = O.mse_l2_grad_logits(logits_ref, target, mask)
    grads_ref, deltas_ref = O.backward(params, c, g_ref, topk_model=topk > 1)
    plan.backward(n, torch.from_numpy(g_ref).cuda())
    torch.cuda.synchronize()
    scale = plan.grad_scale()
    assert scale > 1.0
    dl = {name: ours(f"delta{i}") / scale for name, i in idx.items()}
    gv = {k: v.cpu().numpy() for k, v in plan.grad_views().items()}
    assert sorted(gv.keys()) == sorted(grads_ref.keys())

    # per-layer backward: dgrad from OUR consumer deltas, wgrad from OUR delta and OUR input
    def dact(name):
        i = idx[name]
        return ours(f"d{i}") if name in c["z"] else 1.0

    order = list(idx.keys())
    iso = {}
    iso["pos_linears.0"] = (g_ref @ params["rgb_linear.weight"]) * dact("pos_linears.0")
    if topk > 1:
        wp = params["pos_linears.0.weight"]
        iso["feature_linear2"] = dl["pos_linears.0"] @ wp[:, W:]
        iso["scale_linears.0"] = (dl["feature_linear2"] @ params["feature_linear2.weight"]) * dact("scale_linears.0")
        iso["feature_linear1"] = dl["scale_linears.0"] @ params["scale_linears.0.weight"][:, :W] + dl["pos_linears.0"] @ wp[:, :W]
    else:
        iso["feature_linear1"] = dl["pos_linears.0"] @ params["pos_linears.0.weight"]
    iso["periodic_linears.7"] = (dl["feature_linear1"] @ params["feature_linear1.weight"]) * dact("periodic_linears.7")
    for i in range(6, -1, -1):
        wn = params[f"periodic_linears.{i + 1}.weight"]
        if i == 4:
            wn = wn[:, 462:]
        iso[f"periodic_linears.{i}"] = (dl[f"periodic_linears.{i + 1}"] @ wn) * dact(f"periodic_linears.{i}")
    for name in order:
        rep["bwd_layer"][name] = rel(dl[name], iso[name])
        rep["bwd_cum"][name] = rel(dl[name], deltas_ref[name])
        rep["grad_layer"][name + ".weight"] = rel(gv[name + ".weight"], dl[name].T.astype(np.float64) @ our_input(name))
        rep["grad_layer"][name + ".bias"] = rel(gv[name + ".bias"], dl[name].sum(0, dtype=np.float64))
    rep["grad_layer"]["rgb_linear.weight"] = rel(gv["rgb_linear.weight"], g_ref.T.astype(np.float64) @ hp)
    rep["grad_layer"]["rgb_linear.bias"] = rel(gv["rgb_linear.bias"], g_ref.sum(0, dtype=np.float64))
    for k, ref in grads_ref.items():
        rep["grad_cum"][k] = rel(gv[k], ref)
    _report(f"top{topk}_n{n}", rep)

    TOL_CUM = 2e-3
    for sect in ("fwd_layer", "bwd_layer", "grad_layer"):
        for k, v in rep[sect].items():
            assert v < TOL, (sect, k, v)
    for sect in ("fwd_cum", "bwd_cum", "grad_cum"):
        for k, v in rep[sect].items():
            assert v < TOL_CUM, (sect, k, v)
    # the network output itself (what the PSNR is computed from) stays within the north_star bound
    assert rep["fwd_cum"]["rgb_linear"] < TOL


def test_shallow_network_depth4():
    """netdepth is a plan parameter: D=4 with the skip after layer 1 (forward + backward against the oracle)."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(2)
    n, topk, depth, skips = 500, 3, 4, (1,)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, ANGLES, PERIODS, freqs)
    plan = Plan(enc, depth=depth, skip_layer=skips[0], max_rows=n)
    params = O.init_params(rng, topk=topk, depth=depth, skips=skips)
    plan.load_state(params)
    coords = np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)
    tabs = [(enc.cos_t[j], enc.sin_t[j], enc.period[j]) for j in range(topk)]
    e = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, e, depth=depth, skips=skips)
    logits = plan.forward(torch.from_numpy(coords).cuda())
    assert rel(logits.cpu().numpy(), logits_ref) < TOL
    target = rng.random((n, 3), dtype=np.float32)
    g = O.mse_l2_grad_logits(logits_ref, target)
    grads_ref, _ = O.backward(params, c, g, depth=depth, skips=skips)
    plan.backward(n, torch.from_numpy(g).cuda())
    gv = plan.grad_views()
    assert sorted(gv) == sorted(grads_ref)
    for k, ref in grads_ref.items():
        assert rel(gv[k].cpu().numpy(), ref) < 2e-3, k


@pytest.mark.parametrize("topk", [3, 1])
def test_relu_activation_parity(topk):
    """activation != 'snake' selects F.relu (models/networks.py:51-54,66-69): forward, backward and fused train steps
    against the oracle.  relu'(z) flips between 0 and 1 where fp16-operand rounding moves z across zero (< 0.1 % of the
    units; each flip moves a gradient by a whole delta x input term, ~2 % of a weight gradient's norm in total), so the
    gradients are compared against the oracle backward evaluated with the kernel's own masks."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(7)
    n = 600
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, ANGLES[:topk], PERIODS[:topk], freqs)
    plan = Plan(enc, max_rows=n, activation="relu")
    params = O.init_params(rng, topk=topk)
    plan.load_state(params)
    coords = np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)
    tabs = [(enc.cos_t[j], enc.sin_t[j], enc.period[j]) for j in range(topk)]
    e = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, e, topk_model=topk > 1, activation="relu")
    cd = torch.from_numpy(coords).cuda()
    logits = plan.forward(cd)
    assert rel(logits.cpu().numpy(), logits_ref) < TOL
    h0 = plan.debug("h0", n).cpu().numpy()
    assert h0.min() == 0.0 and rel(h0, c["h"]["periodic_linears.0"]) < TOL
    d0 = plan.debug("d0", n).cpu().numpy()
    assert set(np.unique(d0)) <= {0.0, 1.0} and (d0 != O.relu_grad(c["z"]["periodic_linears.0"])).mean() < 1e-3
    target = rng.random((n, 3), dtype=np.float32)
    g = O.mse_l2_grad_logits(logits_ref, target)
    # the oracle backward with OUR 0/1 masks (relu' only looks at the sign of z): what remains is fp16-operand rounding
    c2 = dict(c)
    c2["z"] = {name: plan.debug(f"d{i}", n).cpu().numpy() - 0.5
               for i, name in enumerate(plan.layer_names) if name in c["z"]}
    flips = sum(((c2["z"][k] > 0) != (c["z"][k] > 0)).sum() for k in c2["z"]) / sum(v.size for v in c2["z"].values())
    assert flips < 1e-3, flips
    grads_ref, _ = O.backward(params, c2, g, topk_model=topk > 1, activation="relu")
    plan.backward(n, torch.from_numpy(g).cuda())
    gv = plan.grad_views()
    assert sorted(gv) == sorted(grads_ref)
    for k, ref in grads_ref.items():
        assert rel(gv[k].cpu().numpy(), ref) < 2e-3, (k, rel(gv[k].cpu().numpy(), ref))
    # fused train steps
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    td = torch.from_numpy(target).cuda()
    loss = torch.zeros((), device="cuda")
    for step in range(1, 5):
        plan.train_step(cd, td, None, O.lr_schedule(step), loss, step=step)
        l_ref, _ = O.train_step(p, m, v, step, e, target, None, O.lr_schedule(step), topk_model=topk > 1,
                                activation="relu")
        assert abs(loss.item() - l_ref) < 2e-3 * l_ref, (step, loss.item(), l_ref)


@pytest.mark.parametrize("topk", [3, 1])
def test_netwidth_256(topk):
    """W = 256 (the constructors' default, models/networks.py:9): one N tile per layer, pos_linears.0 on a zero-padded
    half tile, 128-wide RGB head.  Forward, backward and fused train steps against the oracle."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(11)
    n, W = 900, 256
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(RES, ANGLES[:topk], PERIODS[:topk], freqs)
    plan = Plan(enc, width=W, max_rows=n)
    params = O.init_params(rng, topk=topk, width=W)
    plan.load_state(params)
    assert {s.name: tuple(s.shape) for s in plan.slots} == {k: tuple(v.shape) for k, v in params.items()}
    coords = np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)
    tabs = [(enc.cos_t[j], enc.sin_t[j], enc.period[j]) for j in range(topk)]
    e = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, e, topk_model=topk > 1)
    cd = torch.from_numpy(coords).cuda()
    assert rel(plan.forward(cd).cpu().numpy(), logits_ref) < TOL
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.3).astype(np.float32)
    g = O.mse_l2_grad_logits(logits_ref, target, mask)
    grads_ref, _ = O.backward(params, c, g, topk_model=topk > 1)
    plan.backward(n, torch.from_numpy(g).cuda())
    gv = plan.grad_views()
    assert sorted(gv) == sorted(grads_ref)
    for k, ref in grads_ref.items():
        assert rel(gv[k].cpu().numpy(), ref) < 2e-3, (k, rel(gv[k].cpu().numpy(), ref))
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    td, md = torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda()
    loss = torch.zeros((), device="cuda")
    for step in range(1, 5):
        plan.train_step(cd, td, md, O.lr_schedule(step), loss, step=step)
        l_ref, _ = O.train_step(p, m, v, step, e, target, mask, O.lr_schedule(step), topk_model=topk > 1)
        assert abs(loss.item() - l_ref) < 1e-3 * l_ref, (step, loss.item(), l_ref)
    assert plan.launch_count() == 4      # encode, fused chain (forward + head + loss + backward), wgrad, update


def test_two_plans_stepping_concurrently():
    """Two fits on two streams from two host threads.  The fused head is a cooperative launch (grid barrier); two of
    them in flight at once can dead-lock, so a plan that finds another plan's head still running takes the two-kernel
    head for that step (coop_head_allowed in npp_api.cu).  Both fits must finish and match the same fits run alone."""
    import threading
    steps, n = 40, 4096
    made = [make(1, n, seed=s) for s in (31, 32)]
    data = []
    for plan, params, coords, tabs, freqs, rng in made:
        data.append((torch.from_numpy(coords).cuda(), torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()))

    def run(plan, cd, td, out, stream):
        loss, early = torch.zeros((), device="cuda"), torch.zeros((), device="cuda")
        with torch.cuda.stream(stream):
            for i in range(steps):
                plan.train_step(cd, td, None, 5e-4, early if i < 6 else loss, step=i + 1)
            out.extend([early, loss])

    alone = []
    for (plan, params, *_), (cd, td) in zip(made, data):
        out = []
        run(plan, cd, td, out, torch.cuda.current_stream())
        torch.cuda.synchronize()
        alone.append((out[0].item(), out[1].item()))
        plan.load_state(params)                      # back to the initial weights
        plan.exp_avg.zero_()
        plan.exp_avg_sq.zero_()
    torch.cuda.synchronize()
    outs, threads, streams = [[], []], [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for k in range(2):
        th = threading.Thread(target=run, args=(made[k][0], *data[k], outs[k], streams[k]))
        threads.append(th)
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    for k in range(2):
        early, last = outs[k][0].item(), outs[k][1].item()
        # step 6: same trajectory whichever head ran; step 40: the fits have moved far (loss halves) and fp16 / atomic
        # ordering noise has been amplified by Adam, so only the level is compared
        assert abs(early - alone[k][0]) < 2e-3 * alone[k][0], (k, early, alone[k][0])
        assert abs(last - alone[k][1]) < 0.15 * alone[k][1], (k, last, alone[k][1])


@pytest.mark.timeout(180, method="thread")
def test_three_full_size_fits_from_three_threads_finish():
    """Three 8192-row NPP_Net_top1 fits enqueued concurrently from three host threads on three streams, their kernels
    overlapping on the device.  This dead-locked in 10 of 12 runs while the CTA-pair weight-gradient kernel left part of
    its SMs' shared memory to other blocks (tests/diag_concurrent_big.py); it now asks for the
    whole SM (launch_wgrad in npp_api.cu).  All fits must finish and reproduce the same fits run one after the other."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    from npp_b200.search_fits import run_fits
    K, n, iters = 3, 8192, 60
    rng = np.random.default_rng(5)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    g = torch.Generator().manual_seed(5)
    coords = torch.stack([torch.randint(0, 512, (iters, n), generator=g),
                          torch.randint(0, 512, (iters, n), generator=g)], -1).float().cuda()
    target = torch.rand(iters, n, 3, generator=g).cuda()

    def plans():
        out = []
        for k in range(K):
            enc = EncoderSpec.from_proposals((512, 512), [[97.0, 187.0]], [[40.0 + 3 * k, 36.0 + 2 * k]], freqs)
            p = Plan(enc, max_rows=n)
            p.reset_parameters(seed=k)
            out.append(p)
        return out

    alone = torch.stack([run_fits([p], coords, target, grouped=False)[0] for p in plans()])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(K)]
    for rep in range(3):
        together = run_fits(plans(), coords, target, streams=streams, grouped=False)
        torch.cuda.synchronize()
        assert torch.isfinite(together).all()
        # same kernels, same order inside a fit: only atomic summation order differs
        assert (together[:, :5] - alone[:, :5]).abs().max().item() < 2e-3 * alone[:, :5].max().item()
        assert (together[:, -1] - alone[:, -1]).abs().max().item() < 0.15 * alone[:, -1].max().item()


def test_render_into_image():
    """npp_render_into: chunked forward + sigmoid / tanh scattered straight into an [H, W, 3] image
    (NPP_completion/train.py:277-309), against plan.forward + torch ops; ragged last chunk."""
    plan, params, _, tabs, freqs, rng = make(3, 128, max_rows=4096)
    n = 10000
    flat = rng.choice(RES[0] * RES[1], n, replace=False)
    coords = torch.from_numpy(np.stack([flat // RES[1], flat % RES[1]], 1).astype(np.float32)).cuda()
    logits = torch.cat([plan.forward(coords[i:i + 4096]) for i in range(0, n, 4096)])
    yy, xx = coords[:, 0].long(), coords[:, 1].long()
    for nt, fn in ((1, torch.sigmoid), (2, torch.tanh)):
        image = torch.full((1, RES[0], RES[1], 3), -7.0, device="cuda")
        out = plan.render_into(coords, image, normalize_type=nt)
        assert out is image
        ref = torch.full_like(image, -7.0)
        ref[0, yy, xx, :] = fn(logits)
        assert (image[0, yy, xx, :] - ref[0, yy, xx, :]).abs().max().item() < 2e-6
        untouched = torch.ones(RES, dtype=torch.bool, device="cuda")
        untouched[yy, xx] = False
        assert (image[0][untouched] == -7.0).all()
    with pytest.raises(Exception):
        plan.render_into(coords, torch.zeros(RES[0], RES[1], 4, device="cuda"))


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
    assert plan.launch_count() == 4          # encode, fused chain (forward + head + loss + backward), wgrad, update
    # after 8 Adam steps of size ~5e-4 the weights agree to a fraction of one step
    got = plan.state()
    for k in plan.grad_views():
        assert np.abs(got[k].cpu().numpy() - p[k]).max() < 1e-3, k


def test_prefetched_encoding_gives_the_same_steps():
    """npp_encode_prefetch (next batch encoded on the side stream while a step runs) must not change anything:
    same losses and same weights as encoding in line, over batches that rotate through both encoding sets and
    with a mismatching (unused) prefetch thrown in."""
    n = 3000
    rng = np.random.default_rng(11)
    batches = [torch.from_numpy(np.stack([rng.integers(0, RES[0], n), rng.integers(0, RES[1], n)], 1).astype(np.float32)).cuda()
               for _ in range(4)]
    targets = [torch.rand(n, 3, device="cuda") for _ in range(4)]
    mask = torch.ones(n, 1, device="cuda")

    def run(prefetch):
        plan, *_ = make(3, n)
        loss_d = torch.zeros((), device="cuda")
        losses = []
        for step in range(1, 8):
            b = (step - 1) % 4
            if prefetch:
                plan.prefetch_encode(batches[step % 4])            # the NEXT step's coordinates
                if step == 4:
                    plan.prefetch_encode(batches[(step + 2) % 4])  # a prefetch nobody picks up next (stale later)
            plan.train_step(batches[b], targets[b], mask, 5e-4, loss_d, step=step)
            losses.append(loss_d.item())
        assert plan.launch_count() == 4
        return losses, {k: v.clone() for k, v in plan.state().items()}

    l0, s0 = run(False)
    l1, s1 = run(True)
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-3 * abs(a), (l0, l1)   # run-to-run noise (float atomics) is ~1e-5; a wrong or stale encoding is O(1)
    for k in s0:
        # float atomics (bias-gradient sums, head) make two runs differ in the last bits, and Adam's normalised update
        # can turn that into whole 5e-4 steps on weights whose gradient is ~0: bound the worst weight by the 7 steps
        # taken and the average tightly
        d = (s0[k] - s1[k]).abs()
        mean_tol = 1e-5 if d.numel() >= 1024 else 7 * 5e-4   # a 3-element bias can take opposite steps on noise alone
        assert d.max().item() < 7 * 5e-4 + 1e-6 and d.mean().item() < mean_tol, (k, d.max().item(), d.mean().item())


@pytest.mark.parametrize("topk,n", [(1, 40000), (3, 19001)])
def test_many_stripes_per_cta_forward_matches_oracle(topk, n):
    """More than 148 x 128 rows: every CTA walks several stripes through the chain (forwarding barriers keep
    cycling, phantom stripes at the ragged end).  Logits of a strided sample of rows against the oracle."""
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    logits = plan.forward(torch.from_numpy(coords).cuda()).cpu().numpy()
    assert np.isfinite(logits).all()
    pick = np.r_[0:256, n // 2 - 64:n // 2 + 64, n - 300:n, rng.integers(0, n, 512)]
    enc = O.encode(coords[pick], tabs, freqs, RES)
    ref = O.forward(params, enc, topk_model=topk > 1)[0]
    err = np.linalg.norm(logits[pick] - ref) / np.linalg.norm(ref)
    assert err < 1e-3, err


def test_linearity_of_backward_in_grad():
    """Size-independent property: the weight gradient is linear in dL/dlogits, including the fp16 delta
    scaling (a power of two, so scaling g by 2^k must scale every gradient exactly by 2^k)."""
    n = 16384
    plan, params, coords, tabs, freqs, rng = make(3, n)
    cd = torch.from_numpy(coords).cuda()
    plan.forward(cd)
    g = torch.randn(n, 3, device="cuda") * 1e-5
    plan.backward(n, g)
    g1 = {k: v.clone() for k, v in plan.grad_views().items()}
    plan.backward(n, g * 1024.0)
    g2 = plan.grad_views()
    for k in g1:
        assert torch.isfinite(g1[k]).all() and g1[k].abs().max() > 0, k
        if k.endswith(".weight") and not k.startswith("rgb_linear"):
            # split-K slabs are reduced in a fixed order -> bit-exact linearity
            assert torch.equal(g1[k] * 1024.0, g2[k]), k
        else:
            # bias / head gradients are accumulated with atomics: linear up to fp32 summation order
            err = ((g1[k] * 1024.0 - g2[k]).abs().max() / g2[k].abs().max()).item()
            assert err < 1e-4, (k, err)


def test_rejects_bad_input():
    import npp_b200
    plan, *_ = make(1, 64, max_rows=256)
    with pytest.raises(ValueError):
        plan.forward(torch.zeros(300, 2, device="cuda"))
    with pytest.raises(npp_b200._native.NppError):
        npp_b200._native.check(plan.lib.npp_forward(plan.handle, None, 10, None, None))


def test_psnr_after_training_matches_oracle():
    """Fit a small synthetic near-periodic image for 120 steps with the CUDA path and with the fp32 CPU oracle from the
    same weights / coordinates / LR schedule; the final full-image reconstruction PSNR must agree within 0.1 dB
    (BASELINE.json north_star)."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(3)
    H, W = 48, 64
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    img = np.stack([0.5 + 0.4 * np.sin(2 * np.pi * xx / 16.0) * np.cos(2 * np.pi * yy / 12.0),
                    0.5 + 0.4 * np.cos(2 * np.pi * (xx + yy) / 16.0),
                    0.5 + 0.3 * np.sin(2 * np.pi * yy / 12.0)], -1).astype(np.float32)
    img = np.clip(img + rng.normal(0, 0.01, img.shape).astype(np.float32), 0, 1)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    angles, periods = [[90.0, 180.0]], [[16.0, 12.0]]
    enc = EncoderSpec.from_proposals((H, W), angles, periods, freqs)
    n = 1024
    plan = Plan(enc, max_rows=H * W)
    params = O.init_params(rng, topk=1)
    plan.load_state(params)
    tabs = [(enc.cos_t[0], enc.sin_t[0], enc.period[0])]
    all_coords = np.stack([yy.reshape(-1), xx.reshape(-1)], 1).astype(np.float32)
    table = O.encode(all_coords, tabs, freqs, (H, W))
    target_all = img.reshape(-1, 3)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    mask = np.ones((n, 1), np.float32)
    loss_d = torch.zeros((), device="cuda")
    md = torch.from_numpy(mask).cuda()
    for step in range(1, 121):
        sel = rng.choice(H * W, n, replace=False)
        lr = O.lr_schedule(step)
        plan.train_step(torch.from_numpy(all_coords[sel]).cuda(), torch.from_numpy(target_all[sel]).cuda(), md, lr,
                        loss_d, step=step)
        O.train_step(p, m, v, step, table[sel], target_all[sel], mask, lr, topk_model=False)

    def psnr(pred):
        return -10.0 * np.log10(np.mean((pred - target_all) ** 2))

    ours = torch.sigmoid(plan.forward(torch.from_numpy(all_coords).cuda())).cpu().numpy()
    ref = O.sigmoid(O.forward(p, table, topk_model=False)[0])
    a, b = psnr(ours), psnr(ref)
    _report("psnr", {"psnr_cuda_db": float(a), "psnr_oracle_db": float(b), "steps": 120, "rows_per_step": n})
    assert b > 15.0, b                     # the fit actually learned something
    assert abs(a - b) < 0.1, (a, b)


@pytest.mark.parametrize("topk,n,binary_mask", [(3, 26624, True), (1, 16384, False)])
def test_other_baseline_configs_follow_oracle(topk, n, binary_mask):
    """Shapes of BASELINE.json configs 3 and 5: segmentation 1024^2 (K=3, 8192 + 2*96^2 = 26624 rows, 0/1 loss mask)
    and one of the 192 independent K=1 fits (16384 rows).  Two train steps against the oracle."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(11)
    res = (1024, 1024) if topk == 3 else (512, 512)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(res, ANGLES[:topk], [[p * 2 for p in q] for q in PERIODS[:topk]], freqs)
    plan = Plan(enc, max_rows=n)
    params = O.init_params(rng, topk=topk)
    plan.load_state(params)
    coords = np.stack([rng.integers(0, res[0], n), rng.integers(0, res[1], n)], 1).astype(np.float32)
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.2).astype(np.float32) if binary_mask else np.ones((n, 1), np.float32)
    tabs = [(enc.cos_t[j], enc.sin_t[j], enc.period[j]) for j in range(topk)]
    e = O.encode(coords, tabs, freqs, res)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    cd, td, md = (torch.from_numpy(a).cuda() for a in (coords, target, mask))
    loss_d = torch.zeros((), device="cuda")
    for step in (1, 2):
        plan.train_step(cd, td, md, 5e-4, loss_d, step=step)
        ref_loss, _ = O.train_step(p, m, v, step, e, target, mask, 5e-4, topk_model=topk > 1)
        assert abs(loss_d.item() - ref_loss) < 1e-3 * ref_loss, (step, loss_d.item(), ref_loss)
    got = plan.state()
    for k in plan.grad_views():
        assert np.abs(got[k].cpu().numpy() - p[k]).max() < 1.1e-3, k      # two Adam steps of 5e-4


def test_full_size_batch_additivity():
    """Size-independent property at BASELINE config 4's full size (2^18 rows, 2048^2 image): the gradient of a batch
    equals the sum of the gradients of its two halves when both are normalised by the global row count -- exactly
    what the data-parallel path relies on.  fp32 summation order differs, hence the 1e-4 relative bound."""
    import npp_b200
    from npp_b200.plan import EncoderSpec, Plan
    rng = np.random.default_rng(5)
    n = 1 << 18
    res = (2048, 2048)
    freqs = (rng.standard_normal(10) * 10).astype(np.float32)
    enc = EncoderSpec.from_proposals(res, ANGLES, [[p * 4 for p in q] for q in PERIODS], freqs)
    plan = Plan(enc, max_rows=n)
    plan.load_state(O.init_params(rng, topk=3))
    coords = torch.stack([torch.randint(0, res[0], (n,)), torch.randint(0, res[1], (n,))], 1).float().cuda()
    target = torch.rand(n, 3, device="cuda")

    def grads(lo, hi):
        logits = plan.forward(coords[lo:hi])
        assert torch.isfinite(logits).all()
        _, g, _ = plan.mse(logits, target[lo:hi], None, n_norm=n)
        plan.backward(hi - lo, g)
        return plan.grads[: plan.trained_floats].clone()

    full = grads(0, n)
    halves = grads(0, n // 2) + grads(n // 2, n)
    rel = ((full - halves).norm() / full.norm()).item()
    assert full.abs().max() > 0 and rel < 1e-4, rel


# ------------------------------------------------------------------ fused step: gradients at full batch sizes
def _fused_step_grads(plan, cd, td, md, n, n_norm=None):
    """Gradient arena of ONE fused step (forward + head + loss + backward in one chain launch, then the grouped
    weight-gradient GEMM) through the three-phase API, without applying Adam."""
    n_norm = n if n_norm is None else n_norm
    plan.step_forward_backward(cd, td, md, n_norm)
    plan.step_wgrad(0, plan.layer_count(), n, n_norm)
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in plan.grad_views().items()}


@pytest.mark.parametrize("topk,n", [(3, 26624), (1, 40000), (3, 16384 + 77)])
def test_fused_step_gradients_match_oracle_at_batch_size(topk, n):
    """Backward with more than one stripe per CTA / several split-K row ranges (26 624 rows = 208 stripes on 148 SMs,
    40 000 rows = two split-K ranges) against the fp32 oracle: every weight and bias gradient of the fused step."""
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    e = O.encode(coords, tabs, freqs, RES)
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.3).astype(np.float32)
    logits_ref, c = O.forward(params, e, topk_model=topk > 1)
    g = O.mse_l2_grad_logits(logits_ref, target, mask)
    grads_ref, _ = O.backward(params, c, g, topk_model=topk > 1)
    cd, td, md = (torch.from_numpy(a).cuda() for a in (coords, target, mask))
    gv = _fused_step_grads(plan, cd, td, md, n)
    assert sorted(gv) == sorted(grads_ref)
    worst = max((rel(gv[k].cpu().numpy(), ref), k) for k, ref in grads_ref.items())
    # cumulative bound: 12 chained fp16-operand GEMMs (measured worst 1.2e-3 on the deepest layer)
    assert worst[0] < 1.5e-3, worst
    # and the per-step loss of the same launch
    loss = torch.zeros((), device="cuda")
    plan.step_finish(n, 0.0, loss, step=1)
    ref_loss = float(O.mse_l2(O.sigmoid(logits_ref), target, mask))
    assert abs(loss.item() - ref_loss) < 1e-3 * ref_loss, (loss.item(), ref_loss)


def test_three_phase_step_equals_fused_step():
    """npp_step_forward_backward + npp_step_wgrad (three layer groups, as the data-parallel step issues them) +
    npp_step_finish == npp_train_step on the same rows: same losses, same weights up to fp32 summation order."""
    from npp_b200.dp import DataParallelStep
    n = 5000
    a, params, coords, tabs, freqs, rng = make(3, n)
    b, *_ = make(3, n)
    target = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).cuda()
    mask = torch.from_numpy((rng.random((n, 1)) > 0.3).astype(np.float32)).cuda()
    cd = torch.from_numpy(coords).cuda()
    loss = torch.zeros((), device="cuda")
    dps = DataParallelStep(b, n_buckets=3)
    for step in range(1, 5):
        a.train_step(cd, target, mask, 5e-4, loss, step=step)
        lb = dps(cd, target, mask, 5e-4, n, step=step)
        assert abs(loss.item() - lb.item()) <= 2e-5 * abs(loss.item()), (step, loss.item(), lb.item())
    assert dps.launches == 1 + 1 + 3 * 2 + 2 + 3     # encode, chain, 3 x (wgrad, reduce), 2 rgb copies, adam + shadow + finish
    sa, sb = a.state(), b.state()
    ma = a.view(a.exp_avg, next(s for s in a.slots if s.name == "periodic_linears.3.weight"))
    mb = b.view(b.exp_avg, next(s for s in b.slots if s.name == "periodic_linears.3.weight"))
    assert rel(mb.cpu().numpy(), ma.cpu().numpy()) < 5e-4          # Adam's first moment = running mean of the gradients
    for k in a.grad_views():
        d = (sa[k] - sb[k]).abs()
        assert d.mean().item() < 2e-5, k                           # a fraction of one 5e-4 step on average


def test_adam_trajectory_is_not_vacuous():
    """Eight fused steps against the oracle, judged on quantities that separate a right update from a wrong one:
    exp_avg (the running mean of the gradients) within 2e-3 relative, and the mean |dw| between the two trajectories far
    below the learning rate (a sign error in the update would put it at ~lr)."""
    topk, n = 3, 2048
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    target = rng.random((n, 3), dtype=np.float32)
    mask = np.ones((n, 1), np.float32)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    cd, td, md = (torch.from_numpy(a).cuda() for a in (coords, target, mask))
    loss_d = torch.zeros((), device="cuda")
    lr = 5e-4
    for step in range(1, 9):
        plan.train_step(cd, td, md, lr, loss_d, step=step)
        O.train_step(p, m, v, step, enc, target, mask, lr, topk_model=True)
    got = plan.state()
    slots = {s.name: s for s in plan.slots}
    for k in plan.grad_views():
        ours_m = plan.view(plan.exp_avg, slots[k]).cpu().numpy()
        assert rel(ours_m, m[k]) < 2e-3, (k, rel(ours_m, m[k]))
        mean_dw = float(np.abs(got[k].cpu().numpy() - p[k]).mean())
        assert mean_dw < 0.05 * lr, (k, mean_dw)


def test_fused_step_survives_a_jump_of_the_gradient_maximum():
    """The fused step scales its fp16 deltas with the PREVIOUS step's max|dL/dlogit|.  A converged fit (tiny residuals)
    followed by an outlier batch (residuals of order one) is the worst case: the maximum jumps by three orders of
    magnitude between two steps.  Nothing may overflow, and the gradients of the jump step must still match the oracle."""
    topk, n = 3, 2048
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    logits_ref, c = O.forward(params, enc, topk_model=True)
    pred = O.sigmoid(logits_ref)
    cd = torch.from_numpy(coords).cuda()
    md = torch.ones(n, 1, device="cuda")
    easy = torch.from_numpy((pred + 1e-4 * rng.standard_normal(pred.shape)).astype(np.float32)).cuda()
    loss = torch.zeros((), device="cuda")
    for step in range(1, 4):                 # residuals ~1e-4: the ring now holds a tiny maximum
        plan.train_step(cd, easy, md, 0.0, loss, step=step)          # lr = 0: the weights stay put
        assert loss.item() < 1e-6
    hard = rng.random((n, 3), dtype=np.float32)
    hd = torch.from_numpy(hard).cuda()
    plan.step_forward_backward(cd, hd, md, n)                       # the jump step
    plan.step_wgrad(0, plan.layer_count(), n, n)
    torch.cuda.synchronize()
    g = O.mse_l2_grad_logits(logits_ref, hard, np.ones((n, 1), np.float32))
    grads_ref, _ = O.backward(params, c, g, topk_model=True)
    gv = plan.grad_views()
    for k, ref in grads_ref.items():
        got = gv[k].cpu().numpy()
        assert np.isfinite(got).all(), k
        assert rel(got, ref) < 2e-3, (k, rel(got, ref))
    plan.step_finish(n, 0.0, loss, step=4)
    assert abs(loss.item() - float(O.mse_l2(pred, hard, np.ones((n, 1), np.float32)))) < 1e-3 * loss.item()


@pytest.mark.parametrize("topk,n", [(3, 1), (1, 129), (3, 300)])
def test_fused_step_ragged_row_counts(topk, n):
    """Row counts that fill neither a 32-row TMEM quadrant nor a 128-row stripe: the fused step's loss and the direction
    of its first Adam update against the oracle, and an abandoned three-phase step must not leak into the next step."""
    plan, params, coords, tabs, freqs, rng = make(topk, n)
    enc = O.encode(coords, tabs, freqs, RES)
    target = rng.random((n, 3), dtype=np.float32)
    mask = (rng.random((n, 1)) > 0.3).astype(np.float32)
    mask[0] = 1.0
    cd, td, md = (torch.from_numpy(a).cuda() for a in (coords, target, mask))
    plan.step_forward_backward(cd, td, md, n)          # abandoned on purpose (no wgrad, no finish)
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    loss = torch.zeros((), device="cuda")
    plan.train_step(cd, td, md, 5e-4, loss, step=1)
    ref_loss, _ = O.train_step(p, m, v, 1, enc, target, mask, 5e-4, topk_model=topk > 1)
    assert abs(loss.item() - ref_loss) < 1e-3 * ref_loss, (loss.item(), ref_loss)
    slots = {s.name: s for s in plan.slots}
    for k in plan.grad_views():
        ours_m = plan.view(plan.exp_avg, slots[k]).cpu().numpy()
        assert rel(ours_m, m[k]) < 3e-3, (k, rel(ours_m, m[k]))     # exp_avg = 0.1 * gradient of exactly ONE step
