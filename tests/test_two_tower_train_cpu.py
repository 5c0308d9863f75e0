"""oracle/two_tower_train.py self-checks: towers equal the inference restatement when BN is off, the loss is the
hand-written log-softmax of the adjusted logits, autograd gradients agree with central differences, steps
reduce the loss."""
import numpy as np
import pytest

from oracle import tf_models as tm
from oracle import two_tower_train as tt


def _case(seed, use_bn, B=96, K=8):
    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, 50, 70, [6, 9], [5, 4, 12], 1, 2)
    w = tm.make_two_tower_weights(rng, spec, K, (24, 12), use_bn)
    users, items = rng.integers(0, 50, B), rng.integers(0, 70, B)
    feats = (spec["user_sparse_unique"][users], spec["user_dense_unique"][users],
             spec["item_sparse_unique"][items], spec["item_dense_unique"][items])
    counts = np.bincount(items, minlength=70) / B
    return spec, w, users, items, feats, counts[items]


@pytest.mark.parametrize("norm", [False, True])
def test_forward_matches_inference_restatement_and_manual_loss(norm):
    spec, w, users, items, feats, corr = _case(1, False)
    st = tt.init_state(w, False)
    loss, _, _, U, V = tt.forward_backward(st, users, items, feats, norm=norm, temperature=0.7, correction=corr,
                                           remove_hits=True)
    ru = tm.tower_forward(w, users, feats[0], feats[1], "user", norm, dtype=np.float64)
    ri = tm.tower_forward(w, items, feats[2], feats[3], "item", norm, dtype=np.float64)
    np.testing.assert_allclose(U, ru, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(V, ri, rtol=1e-10, atol=1e-12)
    z = ru @ ri.T / 0.7 - np.log(np.clip(corr, 1e-8, 1.0))[None, :]
    B = len(users)
    hit = (items[None, :] == items[:, None]) & ~np.eye(B, dtype=bool)
    z = np.where(hit, tt.FLT_MIN, z)
    m = z.max(axis=1, keepdims=True)
    want = (np.log(np.exp(z - m).sum(axis=1)) + m[:, 0] - np.diag(z)).mean()
    assert abs(loss - want) < 1e-10


@pytest.mark.parametrize("use_bn", [True, False])
def test_gradients_match_central_differences(use_bn):
    spec, w, users, items, feats, corr = _case(2, use_bn, B=40)
    st = tt.init_state(w, use_bn)
    kw = dict(norm=True, temperature=0.5, correction=corr)
    _, g, _, _, _ = tt.forward_backward(st, users, items, feats, **kw)
    rng = np.random.default_rng(0)
    for k in ("user_W0", "item_b1", "sparse_embeds", "dense_embeds", "user_embeds"):
        p = st["params"][k]
        touched = np.argwhere(np.abs(g[k]) > 1e-9)
        for idx in touched[rng.choice(len(touched), size=min(3, len(touched)), replace=False)]:
            idx = tuple(idx)
            old, h = p[idx], 1e-6
            p[idx] = old + h
            lp = tt.forward_backward(st, users, items, feats, **kw)[0]
            p[idx] = old - h
            lm = tt.forward_backward(st, users, items, feats, **kw)[0]
            p[idx] = old
            assert abs((lp - lm) / (2 * h) - g[k][idx]) <= 1e-5 * max(1.0, abs(g[k][idx])), (k, idx)


def test_steps_reduce_loss():
    spec, w, users, items, feats, corr = _case(3, True)
    st = tt.init_state(w, True)
    losses = [tt.train_step(st, users, items, feats, 1e-2, norm=False, correction=corr) for _ in range(6)]
    assert losses[-1] < losses[0]
    assert st["t"] == 6 and all(np.isfinite(v).all() for v in st["params"].values())
