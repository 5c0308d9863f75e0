"""CPU restatement of ONE FM training step of the reference (TEST INFRASTRUCTURE ONLY).

Follows ``libreco/algorithms/fm.py:140-172`` (graph, ``is_training=True``),
``libreco/tfops/loss.py:14-18`` (mean sigmoid cross entropy) and
``libreco/training/tf_trainer.py:112-123`` (``tf.train.AdamOptimizer(lr, epsilon)`` grouped with the
batch-norm update ops).  **PARITY UNPINNED** for the TensorFlow-specific parts (TensorFlow is absent):

* ``tf.layers.batch_normalization(training=True)`` on a 2-D input: batch mean / BIASED batch variance
  (``tf.nn.moments``), epsilon 1e-3, moving statistics updated with momentum 0.99 using the same
  biased variance (non-fused implementation);
* ``tf.train.AdamOptimizer``: ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)``,
  ``var -= lr_t * m / (sqrt(v) + eps)``; for embedding variables (``IndexedSlices`` gradients,
  ``_apply_sparse_shared``) ``m`` and ``v`` are decayed over the WHOLE variable and the whole
  variable is updated — identical to a dense Adam step with a zero-filled gradient.

The gradient MATH is pinned: ``tests/test_fm_train_cpu.py`` compares the manual backward below with
torch autograd (float64) of the same forward.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
B1, B2 = 0.9, 0.999

TABLES = ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds",
          "user_linear", "item_linear", "sparse_linear", "dense_linear")
DENSE_VARS = ("lin_kernel", "lin_bias", "pw_kernel", "pw_bias", "bn_gamma", "bn_beta")


def init_state(w, use_bn, dtype=np.float64):
    """Trainable variables + Adam slots + BN moving statistics from an FM weight dict
    (oracle.tf_models.make_fm_weights layout)."""
    p = {k: np.array(w[k], dtype=dtype) for k in TABLES if k in w}
    p["lin_kernel"] = np.array(w["lin_kernel"], dtype=dtype).reshape(-1)
    p["lin_bias"] = np.array(w["lin_bias"], dtype=dtype).reshape(1)
    p["pw_kernel"] = np.array(w["pw_kernel"], dtype=dtype).reshape(-1)
    p["pw_bias"] = np.array(w["pw_bias"], dtype=dtype).reshape(1)
    K = p["pw_kernel"].shape[0]
    st = dict(params=p, use_bn=bool(use_bn), t=0)
    if use_bn:
        bn = w.get("fm_bn")
        p["bn_gamma"] = np.array(bn["gamma"] if bn else np.ones(K), dtype=dtype)
        p["bn_beta"] = np.array(bn["beta"] if bn else np.zeros(K), dtype=dtype)
        st["moving_mean"] = np.array(bn["mean"] if bn else np.zeros(K), dtype=dtype)
        st["moving_var"] = np.array(bn["var"] if bn else np.ones(K), dtype=dtype)
    st["m"] = {k: np.zeros_like(v) for k, v in p.items()}
    st["v"] = {k: np.zeros_like(v) for k, v in p.items()}
    return st


def forward_backward(p, use_bn, users, items, sparse, dense, labels):
    """Loss, logits and the gradient of every variable (dense-ified).  Returns (loss, out, grads, bn)."""
    dt = p["user_embeds"].dtype
    R = len(users)
    Fs = sparse.shape[1] if sparse is not None else 0
    Fd = dense.shape[1] if dense is not None else 0
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
    lin = L @ p["lin_kernel"] + p["lin_bias"][0]
    S = P.sum(axis=1)
    pw = 0.5 * (np.square(S) - np.square(P).sum(axis=1))
    bn = None
    if use_bn:
        mu = pw.mean(axis=0)
        var = pw.var(axis=0)                                   # biased
        inv = 1.0 / np.sqrt(var + dt.type(BN_EPS))
        xhat = (pw - mu) * inv
        y = xhat * p["bn_gamma"] + p["bn_beta"]
        bn = (mu, var)
    else:
        y = pw
    z = y @ p["pw_kernel"] + p["pw_bias"][0]
    out = lin + np.where(z > 0, z, np.expm1(z))
    lab = labels.astype(dt)
    loss = (np.maximum(out, 0) - out * lab + np.log1p(np.exp(-np.abs(out)))).mean()

    # ---------------- backward
    sig = np.where(out >= 0, 1 / (1 + np.exp(-np.abs(out))), np.exp(-np.abs(out)) / (1 + np.exp(-np.abs(out))))
    dout = (sig - lab) / R
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g["lin_kernel"] = L.T @ dout
    g["lin_bias"] = np.array([dout.sum()], dtype=dt)
    dL = dout[:, None] * p["lin_kernel"][None, :]
    dz = dout * np.where(z > 0, 1.0, np.exp(z))
    g["pw_kernel"] = y.T @ dz
    g["pw_bias"] = np.array([dz.sum()], dtype=dt)
    dy = dz[:, None] * p["pw_kernel"][None, :]
    if use_bn:
        g["bn_beta"] = dy.sum(axis=0)
        g["bn_gamma"] = (dy * xhat).sum(axis=0)
        dxh = dy * p["bn_gamma"]
        dpw = inv / R * (R * dxh - dxh.sum(axis=0) - xhat * (dxh * xhat).sum(axis=0))
    else:
        dpw = dy
    dP = dpw[:, None, :] * (S[:, None, :] - P)
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
    return loss, out, g, bn


def train_step(st, users, items, sparse, dense, labels, lr, eps=1e-5):
    """One optimisation step in place; returns the (pre-update) loss."""
    p = st["params"]
    dt = p["user_embeds"].dtype
    loss, _, g, bn = forward_backward(p, st["use_bn"], users, items, sparse, dense, labels)
    st["t"] += 1
    t = st["t"]
    lr_t = dt.type(lr) * np.sqrt(1 - dt.type(B2) ** t) / (1 - dt.type(B1) ** t)
    for k in p:
        st["m"][k] = dt.type(B1) * st["m"][k] + dt.type(1 - B1) * g[k]
        st["v"][k] = dt.type(B2) * st["v"][k] + dt.type(1 - B2) * np.square(g[k])
        p[k] -= lr_t * st["m"][k] / (np.sqrt(st["v"][k]) + dt.type(eps))
    if st["use_bn"]:
        mu, var = bn
        st["moving_mean"] = dt.type(BN_MOMENTUM) * st["moving_mean"] + dt.type(1 - BN_MOMENTUM) * mu
        st["moving_var"] = dt.type(BN_MOMENTUM) * st["moving_var"] + dt.type(1 - BN_MOMENTUM) * var
    return float(loss)


def export_weights(st):
    """Back to the inference weight-dict layout (oracle.tf_models / feat_models)."""
    p = st["params"]
    w = {k: p[k].astype(np.float32) for k in TABLES if k in p}
    w.update(lin_kernel=p["lin_kernel"].astype(np.float32), lin_bias=np.float32(p["lin_bias"][0]),
             pw_kernel=p["pw_kernel"].astype(np.float32), pw_bias=np.float32(p["pw_bias"][0]))
    if st["use_bn"]:
        w["fm_bn"] = dict(gamma=p["bn_gamma"].astype(np.float32), beta=p["bn_beta"].astype(np.float32),
                          mean=st["moving_mean"].astype(np.float32), var=st["moving_var"].astype(np.float32))
    return w
