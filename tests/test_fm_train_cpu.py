"""Pins the gradient math of oracle/fm_train.py (manual backward of the FM graph in training mode:
gathers, FM pairwise term, batch-norm with batch statistics, Dense(1, elu), mean sigmoid CE)
against torch autograd in float64, and the Adam / moving-average bookkeeping against a direct
restatement.  (TensorFlow-specific conventions are unpinned — see the oracle's header.)"""
import numpy as np
import pytest
import torch

from oracle import fm_train as ft
from oracle import tf_models as tm


def _case(seed, use_bn):
    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, 40, 60, [5, 9], [7, 4, 11], 1, 2)
    w = tm.make_fm_weights(rng, spec, 8, use_bn)
    R = 257
    users = rng.integers(0, 40, R)
    items = rng.integers(0, 60, R)
    sparse, dense = tm.row_features(spec, users, items)
    labels = (rng.random(R) < 0.4).astype(np.float32)
    return spec, w, users, items, sparse, dense, labels


def _torch_loss(p, use_bn, users, items, sparse, dense, labels):
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    u, i = torch.as_tensor(users), torch.as_tensor(items)
    sp = torch.as_tensor(sparse)
    x = torch.tensor(dense, dtype=torch.float64)
    P = torch.cat([t["user_embeds"][u][:, None], t["item_embeds"][i][:, None], t["sparse_embeds"][sp],
                   x[:, :, None] * t["dense_embeds"][None]], dim=1)
    L = torch.cat([t["user_linear"][u][:, None], t["item_linear"][i][:, None], t["sparse_linear"][sp],
                   x * t["dense_linear"][None]], dim=1)
    lin = L @ t["lin_kernel"] + t["lin_bias"][0]
    pw = 0.5 * (P.sum(1) ** 2 - (P ** 2).sum(1))
    if use_bn:
        mu, var = pw.mean(0), pw.var(0, unbiased=False)
        pw = (pw - mu) / torch.sqrt(var + 1e-3) * t["bn_gamma"] + t["bn_beta"]
    z = pw @ t["pw_kernel"] + t["pw_bias"][0]
    out = lin + torch.nn.functional.elu(z)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.tensor(labels, dtype=torch.float64))
    loss.backward()
    return float(loss), {k: v.grad.numpy() for k, v in t.items()}


@pytest.mark.parametrize("use_bn", [True, False])
def test_manual_backward_equals_autograd(use_bn):
    spec, w, users, items, sparse, dense, labels = _case(3, use_bn)
    st = ft.init_state(w, use_bn)
    loss, out, g, _ = ft.forward_backward(st["params"], use_bn, users, items, sparse, dense, labels)
    ref_loss, ref_g = _torch_loss(st["params"], use_bn, users, items, sparse, dense, labels)
    assert abs(loss - ref_loss) < 1e-12
    for k in ref_g:
        np.testing.assert_allclose(g[k], ref_g[k], rtol=1e-9, atol=1e-12, err_msg=k)
    # the training-mode forward with batch statistics differs from inference (moving statistics)
    ref_inf = tm.fm_forward(w, users, items, sparse, dense, dtype=np.float64)
    if use_bn:
        assert np.abs(out - ref_inf).max() > 1e-3
    else:
        np.testing.assert_allclose(out, ref_inf, rtol=1e-10, atol=1e-12)


def test_adam_and_moving_statistics_bookkeeping():
    spec, w, users, items, sparse, dense, labels = _case(5, True)
    st = ft.init_state(w, True)
    p0 = {k: v.copy() for k, v in st["params"].items()}
    mm0, mv0 = st["moving_mean"].copy(), st["moving_var"].copy()
    _, _, g, bn = ft.forward_backward(st["params"], True, users, items, sparse, dense, labels)
    lr, eps = 1e-2, 1e-5
    losses = [ft.train_step(st, users, items, sparse, dense, labels, lr, eps) for _ in range(3)]
    assert losses[2] < losses[0]                                   # same batch: the loss goes down
    # first step by hand: m = 0.1 g, v = 0.001 g^2, lr_1 = lr * sqrt(0.001) / 0.1
    st1 = ft.init_state(w, True)
    ft.train_step(st1, users, items, sparse, dense, labels, lr, eps)
    lr1 = lr * np.sqrt(1 - 0.999) / (1 - 0.9)
    for k in p0:
        want = p0[k] - lr1 * (0.1 * g[k]) / (np.sqrt(0.001 * g[k] ** 2) + eps)
        np.testing.assert_allclose(st1["params"][k], want, rtol=1e-12, atol=1e-15, err_msg=k)
    # rows without a gradient do not move on the first step (m = v = 0) ...
    untouched = np.setdiff1d(np.arange(41), users)
    np.testing.assert_array_equal(st1["params"]["user_embeds"][untouched], p0["user_embeds"][untouched])
    # ... but keep moving afterwards only if they once had one (dense decay of m over the variable)
    np.testing.assert_allclose(st1["moving_mean"], 0.99 * mm0 + 0.01 * bn[0])
    np.testing.assert_allclose(st1["moving_var"], 0.99 * mv0 + 0.01 * bn[1])
    w2 = ft.export_weights(st)
    assert set(w2) >= {"user_embeds", "lin_kernel", "pw_kernel", "fm_bn"}
