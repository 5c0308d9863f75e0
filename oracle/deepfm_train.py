"""CPU restatement of ONE DeepFM training step of the reference (TEST INFRASTRUCTURE ONLY).

Follows ``libreco/algorithms/deepfm.py:143-175`` (graph with ``is_training=True``),
``libreco/layers/dense.py:12-49`` (``dense_nn``: BN(input) -> [Dense -> ReLU -> BN] x (L-1) -> Dense,
``dropout_rate=None`` as in the reference's default), ``libreco/tfops/loss.py:14-18`` and
``libreco/training/tf_trainer.py:112-123``.  TensorFlow conventions (batch-norm, Adam) as stated in
``oracle/fm_train.py`` — **PARITY UNPINNED** for those; the gradient MATH is pinned against torch
autograd (``tests/test_deepfm_train_cpu.py``).
"""
from __future__ import annotations

import numpy as np

from .fm_train import B1, B2, BN_EPS, BN_MOMENTUM, TABLES


def _bn_train(x, gamma, beta):
    mu = x.mean(axis=0)
    var = x.var(axis=0)
    inv = 1.0 / np.sqrt(var + x.dtype.type(BN_EPS))
    xhat = (x - mu) * inv
    return xhat * gamma + beta, (xhat, inv, mu, var)


def _bn_back(dy, cache, gamma):
    xhat, inv, _, _ = cache
    R = len(dy)
    dxh = dy * gamma
    dx = inv / R * (R * dxh - dxh.sum(axis=0) - xhat * (dxh * xhat).sum(axis=0))
    return dx, (dy * xhat).sum(axis=0), dy.sum(axis=0)


def init_state(w, use_bn, dtype=np.float64):
    """Variables from a DeepFM weight dict (oracle.tf_models.make_deepfm_weights layout)."""
    p = {k: np.array(w[k], dtype=dtype) for k in TABLES if k in w}
    p["lin_kernel"] = np.array(w["lin_kernel"], dtype=dtype).reshape(-1)
    p["lin_bias"] = np.array(w["lin_bias"], dtype=dtype).reshape(1)
    p["out_kernel"] = np.array(w["out_kernel"], dtype=dtype).reshape(-1)
    p["out_bias"] = np.array(w["out_bias"], dtype=dtype).reshape(1)
    mlp = w["mlp"]
    n = len(mlp["kernels"])
    st = dict(use_bn=bool(use_bn), n_layers=n, t=0, moving={})
    for i in range(n):
        p[f"W{i}"] = np.array(mlp["kernels"][i], dtype=dtype)          # [din, dout]
        p[f"b{i}"] = np.array(mlp["biases"][i], dtype=dtype)
    if use_bn:
        bns = [mlp.get("bn_in")] + list(mlp.get("bns") or [])
        for j, bn in enumerate(bns):                                    # bn0 = input BN, bn{i+1} after layer i
            p[f"bn{j}_gamma"] = np.array(bn["gamma"], dtype=dtype)
            p[f"bn{j}_beta"] = np.array(bn["beta"], dtype=dtype)
            st["moving"][f"bn{j}"] = [np.array(bn["mean"], dtype=dtype), np.array(bn["var"], dtype=dtype)]
    st["params"] = p
    st["m"] = {k: np.zeros_like(v) for k, v in p.items()}
    st["v"] = {k: np.zeros_like(v) for k, v in p.items()}
    return st


def forward_backward(st, users, items, sparse, dense, labels):
    p, use_bn, n = st["params"], st["use_bn"], st["n_layers"]
    dt = p["user_embeds"].dtype
    R = len(users)
    Fs = sparse.shape[1] if sparse is not None else 0
    Fd = dense.shape[1] if dense is not None else 0
    K = p["user_embeds"].shape[1]
    P = [p["user_embeds"][users][:, None, :], p["item_embeds"][items][:, None, :]]
    L = [p["user_linear"][users][:, None], p["item_linear"][items][:, None]]
    if Fs:
        P.append(p["sparse_embeds"][sparse])
        L.append(p["sparse_linear"][sparse])
    if Fd:
        x = dense.astype(dt)
        P.append(x[:, :, None] * p["dense_embeds"][None, :, :])
        L.append(x * p["dense_linear"][None, :])
    P = np.concatenate(P, axis=1)
    L = np.concatenate(L, axis=1)
    F = P.shape[1]
    lin = L @ p["lin_kernel"] + p["lin_bias"][0]
    S = P.sum(axis=1)
    pw = 0.5 * (np.square(S) - np.square(P).sum(axis=1))
    # ---- dense_nn in training mode
    a = P.reshape(R, F * K)
    caches, stats = [], {}
    if use_bn:
        a, c = _bn_train(a, p["bn0_gamma"], p["bn0_beta"])
        caches.append(("bn", 0, c))
        stats["bn0"] = (c[2], c[3])
    for i in range(n):
        h = a @ p[f"W{i}"] + p[f"b{i}"]
        caches.append(("dense", i, a))
        a = h
        if i != n - 1:
            caches.append(("relu", i, h))
            a = np.maximum(h, 0)
            if use_bn:
                a, c = _bn_train(a, p[f"bn{i + 1}_gamma"], p[f"bn{i + 1}_beta"])
                caches.append(("bn", i + 1, c))
                stats[f"bn{i + 1}"] = (c[2], c[3])
    deep = a
    feat = np.concatenate([lin[:, None], pw, deep], axis=1)
    out = feat @ p["out_kernel"] + p["out_bias"][0]
    lab = labels.astype(dt)
    loss = (np.maximum(out, 0) - out * lab + np.log1p(np.exp(-np.abs(out)))).mean()

    # ---- backward
    sig = np.where(out >= 0, 1 / (1 + np.exp(-np.abs(out))), np.exp(-np.abs(out)) / (1 + np.exp(-np.abs(out))))
    dout = (sig - lab) / R
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g["out_kernel"] = feat.T @ dout
    g["out_bias"] = np.array([dout.sum()], dtype=dt)
    dfeat = dout[:, None] * p["out_kernel"][None, :]
    dlin, dpw, da = dfeat[:, 0], dfeat[:, 1:1 + K], dfeat[:, 1 + K:]
    for kind, i, c in reversed(caches):
        if kind == "bn":
            da, g[f"bn{i}_gamma"], g[f"bn{i}_beta"] = _bn_back(da, c, p[f"bn{i}_gamma"])
        elif kind == "relu":
            da = da * (c > 0)
        else:
            g[f"W{i}"] = c.T @ da
            g[f"b{i}"] = da.sum(axis=0)
            da = da @ p[f"W{i}"].T
    dconcat = da.reshape(R, F, K)
    g["lin_kernel"] = L.T @ dlin
    g["lin_bias"] = np.array([dlin.sum()], dtype=dt)
    dL = dlin[:, None] * p["lin_kernel"][None, :]
    dP = dpw[:, None, :] * (S[:, None, :] - P) + dconcat
    np.add.at(g["user_embeds"], users, dP[:, 0])
    np.add.at(g["item_embeds"], items, dP[:, 1])
    np.add.at(g["user_linear"], users, dL[:, 0])
    np.add.at(g["item_linear"], items, dL[:, 1])
    for f in range(Fs):
        np.add.at(g["sparse_embeds"], sparse[:, f], dP[:, 2 + f])
        np.add.at(g["sparse_linear"], sparse[:, f], dL[:, 2 + f])
    for f in range(Fd):
        g["dense_embeds"][f] = (x[:, f, None] * dP[:, 2 + Fs + f]).sum(axis=0)
        g["dense_linear"][f] = (x[:, f] * dL[:, 2 + Fs + f]).sum()
    return loss, out, g, stats


def train_step(st, users, items, sparse, dense, labels, lr, eps=1e-5, reg=0.0, decay_steps=0, decay_rate=0.96):
    """``reg``: tf.keras.regularizers.l2 on the embedding / linear tables (deepfm.py:186-259; tfops/configs.py:20-26):
    optimised loss = data loss + reg * sum w^2, the returned loss is the data loss (tf_trainer.py: sess.run(self.loss)).
    ``decay_steps`` > 0: tf.train.exponential_decay(lr, global_step, decay_steps, decay_rate, staircase=True)."""
    p = st["params"]
    dt = p["user_embeds"].dtype
    loss, _, g, stats = forward_backward(st, users, items, sparse, dense, labels)
    if reg:
        for k in TABLES:
            if k in p:
                g[k] = g[k] + dt.type(2.0 * reg) * p[k]
    if decay_steps:
        lr = lr * decay_rate ** (st["t"] // decay_steps)          # global_step = completed steps
    st["t"] += 1
    t = st["t"]
    lr_t = dt.type(lr) * np.sqrt(1 - dt.type(B2) ** t) / (1 - dt.type(B1) ** t)
    for k in p:
        st["m"][k] = dt.type(B1) * st["m"][k] + dt.type(1 - B1) * g[k]
        st["v"][k] = dt.type(B2) * st["v"][k] + dt.type(1 - B2) * np.square(g[k])
        p[k] -= lr_t * st["m"][k] / (np.sqrt(st["v"][k]) + dt.type(eps))
    for name, (mu, var) in stats.items():
        mm, mv = st["moving"][name]
        st["moving"][name] = [dt.type(BN_MOMENTUM) * mm + dt.type(1 - BN_MOMENTUM) * mu,
                              dt.type(BN_MOMENTUM) * mv + dt.type(1 - BN_MOMENTUM) * var]
    return float(loss)


def export_weights(st):
    """Back to the inference weight-dict layout (oracle.tf_models.make_deepfm_weights)."""
    p, n = st["params"], st["n_layers"]
    f32 = np.float32
    w = {k: p[k].astype(f32) for k in TABLES if k in p}
    w.update(lin_kernel=p["lin_kernel"].astype(f32), lin_bias=f32(p["lin_bias"][0]),
             out_kernel=p["out_kernel"].astype(f32), out_bias=f32(p["out_bias"][0]))
    mlp = dict(kernels=[p[f"W{i}"].astype(f32) for i in range(n)], biases=[p[f"b{i}"].astype(f32) for i in range(n)])
    if st["use_bn"]:
        def bn(j):
            mm, mv = st["moving"][f"bn{j}"]
            return dict(gamma=p[f"bn{j}_gamma"].astype(f32), beta=p[f"bn{j}_beta"].astype(f32),
                        mean=mm.astype(f32), var=mv.astype(f32))
        mlp["bn_in"] = bn(0)
        mlp["bns"] = [bn(i + 1) for i in range(n - 1)]
    w["mlp"] = mlp
    return w
