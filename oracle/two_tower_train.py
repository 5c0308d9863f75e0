"""CPU restatement of ONE TwoTower training step of the reference, in-batch softmax loss
(TEST INFRASTRUCTURE ONLY — never imported by the product path).

Follows ``libreco/algorithms/two_tower.py:306-346,400-410`` (towers: concat(id embedding, sparse
embeddings, dense value x embedding) -> ``dense_nn`` in training mode -> optional
``tf.linalg.l2_normalize``), ``libreco/tfops/loss.py:71-75`` and ``two_tower.py:458-479``
(``logits = U V^T / temperature - log(clip(Q, 1e-8, 1))``, optional accidental-hit mask, mean sparse
softmax CE with the diagonal as labels) and ``libreco/training/tf_trainer.py:112-123`` (TF-Adam + BN
update ops).  The forward is written with torch float64 tensors and the gradients come from torch
autograd (the floating-point reference this tier allows); TensorFlow conventions (batch-norm momentum /
epsilon, Adam bias correction, dense Adam over embedding variables) are as stated in
``oracle/fm_train.py`` — **PARITY UNPINNED** for those (TensorFlow is not installable here).
"""
from __future__ import annotations

import numpy as np
import torch

from .fm_train import B1, B2, BN_EPS, BN_MOMENTUM

TABLES = ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds")
FLT_MIN = float(np.finfo(np.float32).min)      # tf.float32.min


def init_state(w, use_bn):
    p = {k: np.array(w[k], dtype=np.float64) for k in TABLES if w.get(k) is not None}
    st = dict(use_bn=bool(use_bn), t=0, moving={}, n_layers={})
    for which in ("user", "item"):
        mlp = w[f"{which}_tower"]
        n = st["n_layers"][which] = len(mlp["kernels"])
        for i in range(n):
            p[f"{which}_W{i}"] = np.array(mlp["kernels"][i], dtype=np.float64)      # [din, dout]
            p[f"{which}_b{i}"] = np.array(mlp["biases"][i], dtype=np.float64)
        if use_bn:
            for j, bn in enumerate([mlp.get("bn_in")] + list(mlp.get("bns") or [])):
                p[f"{which}_bn{j}_gamma"] = np.array(bn["gamma"], dtype=np.float64)
                p[f"{which}_bn{j}_beta"] = np.array(bn["beta"], dtype=np.float64)
                st["moving"][f"{which}_bn{j}"] = [np.array(bn["mean"], dtype=np.float64),
                                                  np.array(bn["var"], dtype=np.float64)]
    st["dense_cols"] = dict(user=list(w.get("user_dense_cols", [])), item=list(w.get("item_dense_cols", [])))
    st["params"] = p
    st["m"] = {k: np.zeros_like(v) for k, v in p.items()}
    st["v"] = {k: np.zeros_like(v) for k, v in p.items()}
    return st


def _tower(t, st, which, ids, sparse, dense, norm, stats):
    parts = [t[f"{which}_embeds"][torch.as_tensor(ids)]]
    n = len(ids)
    if sparse is not None and sparse.shape[1]:
        parts.append(t["sparse_embeds"][torch.as_tensor(sparse)].reshape(n, -1))
    if dense is not None and dense.shape[1]:
        x = torch.tensor(np.asarray(dense), dtype=torch.float64)
        parts.append((x[:, :, None] * t["dense_embeds"][st["dense_cols"][which]][None]).reshape(n, -1))
    a = torch.cat(parts, dim=1)

    def bn(a, j):
        mu, var = a.mean(0), a.var(0, unbiased=False)
        stats[f"{which}_bn{j}"] = (mu.detach().numpy(), var.detach().numpy())
        return (a - mu) / torch.sqrt(var + BN_EPS) * t[f"{which}_bn{j}_gamma"] + t[f"{which}_bn{j}_beta"]

    if st["use_bn"]:
        a = bn(a, 0)
    L = st["n_layers"][which]
    for i in range(L):
        a = a @ t[f"{which}_W{i}"] + t[f"{which}_b{i}"]
        if i != L - 1:
            a = torch.relu(a)
            if st["use_bn"]:
                a = bn(a, i + 1)
    if norm:
        a = a * torch.rsqrt(torch.clamp((a * a).sum(1, keepdim=True), min=1e-12))
    return a


def forward_backward(st, users, items, feats, norm=False, temperature=1.0, correction=None, remove_hits=False):
    """``feats`` = (user_sparse [B,Fus] | None, user_dense | None, item_sparse | None, item_dense | None).
    Returns (loss, grads dict, batch BN statistics, U, V)."""
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in st["params"].items()}
    stats = {}
    us, ud, is_, idn = feats
    U = _tower(t, st, "user", users, us, ud, norm, stats)
    V = _tower(t, st, "item", items, is_, idn, norm, stats)
    logits = U @ V.T / temperature
    if correction is not None:
        c = torch.clamp(torch.tensor(np.asarray(correction), dtype=torch.float64), 1e-8, 1.0)
        logits = logits - torch.log(c)[None, :]
    B = len(users)
    if remove_hits:
        it = torch.as_tensor(np.asarray(items))
        mask = (it[None, :] == it[:, None]) & ~torch.eye(B, dtype=torch.bool)
        logits = torch.where(mask, torch.full_like(logits, FLT_MIN), logits)
    loss = torch.nn.functional.cross_entropy(logits, torch.arange(B))
    loss.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(st["params"][k])) for k, v in t.items()}
    return float(loss.detach()), g, stats, U.detach().numpy(), V.detach().numpy()


def train_step(st, users, items, feats, lr, eps=1e-5, **kw):
    loss, g, stats, _, _ = forward_backward(st, users, items, feats, **kw)
    p = st["params"]
    st["t"] += 1
    t = st["t"]
    lr_t = lr * np.sqrt(1 - B2 ** t) / (1 - B1 ** t)
    for k in p:
        st["m"][k] = B1 * st["m"][k] + (1 - B1) * g[k]
        st["v"][k] = B2 * st["v"][k] + (1 - B2) * np.square(g[k])
        p[k] -= lr_t * st["m"][k] / (np.sqrt(st["v"][k]) + eps)
    for name, (mu, var) in stats.items():
        mm, mv = st["moving"][name]
        st["moving"][name] = [BN_MOMENTUM * mm + (1 - BN_MOMENTUM) * mu, BN_MOMENTUM * mv + (1 - BN_MOMENTUM) * var]
    return loss
