"""Pins oracle/npp_oracle.py against golden vectors produced by the live reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: the oracle is numpy fp32, the reference torch
fp32 -- both round every step to fp32 but use different libm / BLAS summation orders, so results agree
to a few ulp: encoding 2e-6 abs (values in [-1,1]), activations/gradients 2e-5 relative."""
import os

import numpy as np
import pytest

from oracle import npp_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def test_encoder_tables_and_encoding():
    g = np.load(os.path.join(G, "golden_encoding.npz"))
    res = tuple(g["res"])
    tabs = []
    for j in range(3):
        c, s, p = O.encoder_tables(g["angles"][j], g["periods"][j], [1], [0, -1, 1, 0.5, -0.5], [0])
        # numpy vs torch cos/sin may differ in the last ulp; periods are exact
        np.testing.assert_allclose(c, g["cos_t"][j], atol=1e-7)
        np.testing.assert_allclose(s, g["sin_t"][j], atol=1e-7)
        np.testing.assert_array_equal(p, g["period"][j])
        tabs.append((g["cos_t"][j], g["sin_t"][j], g["period"][j]))
    base = np.concatenate([O.encode_periodic(g["coords"], *t, res) for t in tabs], 1)
    assert base.shape == g["base"].shape == (40, 66)
    np.testing.assert_allclose(base, g["base"], atol=2e-6)
    full = O.encode(g["coords"], tabs, g["freqs"], res)
    assert full.shape == g["full"].shape == (40, 1386)
    np.testing.assert_allclose(full, g["full"], atol=5e-5)   # |f| <= ~22 amplifies the base ulps
    # column rule: out[:, b*22 + c]
    u = g["base"][:, :22]
    np.testing.assert_allclose(full[:, 22 * 3 + 5], np.sin(u[:, 5] * g["freqs"][1]), atol=5e-5)


@pytest.mark.parametrize("tag,topk", [("topk", 3), ("top1", 1), ("relu", 3)])
def test_forward_backward_adam(tag, topk):
    g = np.load(os.path.join(G, f"golden_{tag}.npz"))
    act = "relu" if tag == "relu" else "snake"
    p = {k[5:]: g[k].copy() for k in g.files if k.startswith("init/")}
    logits, c = O.forward(p, g["enc"], topk_model=topk > 1, activation=act)
    for k in [k for k in g.files if k.startswith("z/")]:
        name = k[2:]
        if name == "rgb_linear":
            ours = logits
        elif name in c["z"]:
            ours = c["z"][name]
        elif name in c["h"]:
            ours = c["h"][name]       # activation-free feature_linear1/2
        else:
            continue                  # alpha_linear / unused feature_linear2 never run
        assert rel(ours, g[k]) < 2e-5, name
    np.testing.assert_allclose(O.sigmoid(logits), g["pred"], atol=1e-6)
    loss = O.mse_l2(O.sigmoid(logits), g["target"], g["mask"])
    assert abs(loss - g["losses"][0]) < 1e-6
    grads, _ = O.backward(p, c, O.mse_l2_grad_logits(logits, g["target"], g["mask"]), topk_model=topk > 1,
                          activation=act)
    gkeys = [k[5:] for k in g.files if k.startswith("grad/")]
    assert sorted(gkeys) == sorted(grads.keys())          # same set of trained parameters
    for k in gkeys:
        assert rel(grads[k], g["grad/" + k]) < 2e-5, k

    # three Adam steps with the reference's LR rewrite
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    losses = []
    for it in range(1, 4):
        l, _ = O.train_step(p, m, v, it, g["enc"], g["target"], g["mask"], O.lr_schedule(it), topk_model=topk > 1,
                            activation=act)
        losses.append(l)
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    for k in p:
        # Adam's first steps move every weight by ~lr regardless of gradient size; fp32 noise in tiny
        # gradients can flip that, so compare against the step size rather than the weight magnitude.
        assert np.abs(p[k] - g["final/" + k]).max() < 2e-4, k
        assert rel(p[k], g["final/" + k]) < 5e-4, k


def test_lr_schedule():
    assert O.lr_schedule(1) == 5e-4 and O.lr_schedule(2) == 5e-4
    assert abs(O.lr_schedule(3) - 5e-4 * 0.1 ** (1 / 50000)) < 1e-12


def test_search_mode_encoders_and_light_network():
    """NPP_Net_light and the search-mode encoders (tests/golden/make_golden_light.py)."""
    g = np.load(os.path.join(G, "golden_light.npz"))
    res = tuple(g["res"])
    c_, s_, p_ = O.encoder_tables(g["angles"], g["periods"], [1], [0, -1, 1, 0.5, -0.5], [0])
    np.testing.assert_allclose(c_, g["cos_t"], atol=1e-7)
    np.testing.assert_allclose(s_, g["sin_t"], atol=1e-7)
    np.testing.assert_array_equal(p_, g["period"])
    pos, per = O.encode_search(g["coords"], (g["cos_t"], g["sin_t"], g["period"]), g["freqs"], res)
    assert pos.shape == g["pos"].shape == (48, 42) and per.shape == g["per"].shape == (48, 20)
    np.testing.assert_allclose(per, g["per"], atol=2e-6)
    np.testing.assert_allclose(pos, g["pos"], atol=5e-6)      # |f u| <= ~25: a few ulp of the argument

    p = {k[5:]: g[k].copy() for k in g.files if k.startswith("init/")}
    logits, c = O.forward_light(p, g["pos"], g["per"])
    for k in [k for k in g.files if k.startswith("z/")]:
        name = k[2:]
        ours = logits if name == "rgb_linear" else c["z"].get(name, c["h"].get(name))
        assert ours is not None, name
        assert rel(ours, g[k]) < 2e-5, name
    np.testing.assert_allclose(logits, g["logits"], atol=2e-6)
    grads, _ = O.backward_light(p, c, O.mse_l2_grad_logits(logits, g["target"], None))
    gkeys = [k[5:] for k in g.files if k.startswith("grad/")]
    assert sorted(gkeys) == sorted(grads.keys())     # scale_linears / feature_linear2 / alpha_linear never train
    for k in gkeys:
        assert rel(grads[k], g["grad/" + k]) < 2e-5, k
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(v_) for k, v_ in p.items()}
    losses = []
    for it in range(1, 4):
        l, _ = O.train_step_light(p, m, v, it, g["pos"], g["per"], g["target"], O.lr_schedule(it))
        losses.append(l)
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    for k in p:
        assert np.abs(p[k] - g["final/" + k]).max() < 2e-4, k
        assert rel(p[k], g["final/" + k]) < 5e-4, k
