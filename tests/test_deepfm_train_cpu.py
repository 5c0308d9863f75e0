"""Pins the gradient math of oracle/deepfm_train.py (gathers, FM term, dense_nn with batch-norm in
training mode, final Dense(1), mean sigmoid CE) against torch autograd in float64."""
import numpy as np
import pytest
import torch

from oracle import deepfm_train as dt_
from oracle import tf_models as tm


def _case(seed, use_bn, hidden=(24, 16, 8)):
    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, 40, 60, [5, 9], [7, 4, 11], 1, 2)
    w = tm.make_deepfm_weights(rng, spec, 8, hidden, use_bn)
    R = 193
    users, items = rng.integers(0, 40, R), rng.integers(0, 60, R)
    sparse, dense = tm.row_features(spec, users, items)
    labels = (rng.random(R) < 0.4).astype(np.float32)
    return spec, w, users, items, sparse, dense, labels


def _torch_loss(st, users, items, sparse, dense, labels):
    p, n, use_bn = st["params"], st["n_layers"], st["use_bn"]
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    u, i, sp = torch.as_tensor(users), torch.as_tensor(items), torch.as_tensor(sparse)
    x = torch.tensor(dense, dtype=torch.float64)
    P = torch.cat([t["user_embeds"][u][:, None], t["item_embeds"][i][:, None], t["sparse_embeds"][sp],
                   x[:, :, None] * t["dense_embeds"][None]], dim=1)
    L = torch.cat([t["user_linear"][u][:, None], t["item_linear"][i][:, None], t["sparse_linear"][sp],
                   x * t["dense_linear"][None]], dim=1)
    lin = L @ t["lin_kernel"] + t["lin_bias"][0]
    pw = 0.5 * (P.sum(1) ** 2 - (P ** 2).sum(1))

    def bn(a, j):
        mu, var = a.mean(0), a.var(0, unbiased=False)
        return (a - mu) / torch.sqrt(var + 1e-3) * t[f"bn{j}_gamma"] + t[f"bn{j}_beta"]

    a = P.reshape(len(users), -1)
    if use_bn:
        a = bn(a, 0)
    for l in range(n):
        a = a @ t[f"W{l}"] + t[f"b{l}"]
        if l != n - 1:
            a = torch.relu(a)
            if use_bn:
                a = bn(a, l + 1)
    out = torch.cat([lin[:, None], pw, a], dim=1) @ t["out_kernel"] + t["out_bias"][0]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.tensor(labels, dtype=torch.float64))
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy() for k, v in t.items()}


@pytest.mark.parametrize("use_bn", [True, False])
@pytest.mark.parametrize("hidden", [(24, 16, 8), (16,)])
def test_manual_backward_equals_autograd(use_bn, hidden):
    spec, w, users, items, sparse, dense, labels = _case(3, use_bn, hidden)
    st = dt_.init_state(w, use_bn)
    loss, out, g, _ = dt_.forward_backward(st, users, items, sparse, dense, labels)
    ref_loss, ref_g = _torch_loss(st, users, items, sparse, dense, labels)
    assert abs(loss - ref_loss) < 1e-12
    assert set(g) == set(ref_g)
    for k in ref_g:
        np.testing.assert_allclose(g[k], ref_g[k], rtol=1e-8, atol=1e-12, err_msg=k)
    if not use_bn:       # without BN training == inference forward
        ref_inf = tm.deepfm_forward(w, users, items, sparse, dense, dtype=np.float64)
        np.testing.assert_allclose(out, ref_inf, rtol=1e-10, atol=1e-12)


def test_steps_reduce_loss_and_export_roundtrip():
    spec, w, users, items, sparse, dense, labels = _case(9, True)
    st = dt_.init_state(w, True)
    losses = [dt_.train_step(st, users, items, sparse, dense, labels, 1e-2) for _ in range(5)]
    assert losses[-1] < losses[0]
    w2 = dt_.export_weights(st)
    out = tm.deepfm_forward(w2, users, items, sparse, dense, dtype=np.float64)
    assert np.isfinite(out).all() and len(w2["mlp"]["bns"]) == 2
