"""CPU restatement of ONE DIN training step of the reference (TEST INFRASTRUCTURE ONLY).

Follows ``libreco/algorithms/din.py:165-250`` with ``is_training=True`` and the paper attention
(``libreco/layers/attention.py:28-64``: Dense(1)(sigmoid(Dense(16)([q, k, q-k, q*k]))) * rsqrt(K'), masked softmax,
weighted key sum) over the item feature table of ``libreco/tfops/features.py:165-218`` (concat mode),
``libreco/tfops/loss.py:14-18`` (mean sigmoid CE), ``libreco/training/tf_trainer.py:112-123`` (TF-Adam + BN update
ops).  Forward in torch float64, gradients from torch autograd; TensorFlow conventions as in ``oracle/fm_train.py`` —
**PARITY UNPINNED** for those.
"""
from __future__ import annotations

import numpy as np
import torch

from .fm_train import B1, B2, BN_EPS, BN_MOMENTUM

TABLES = ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds")


def init_state(w, use_bn):
    p = {k: np.array(w[k], dtype=np.float64) for k in TABLES if w.get(k) is not None}
    mlp = w["mlp"]
    n = len(mlp["kernels"])
    st = dict(use_bn=bool(use_bn), t=0, moving={}, n_layers=n)
    for i in range(n):
        p[f"W{i}"] = np.array(mlp["kernels"][i], dtype=np.float64)
        p[f"b{i}"] = np.array(mlp["biases"][i], dtype=np.float64)
    if use_bn:
        for j, bn in enumerate([mlp.get("bn_in")] + list(mlp.get("bns") or [])):
            p[f"bn{j}_gamma"] = np.array(bn["gamma"], dtype=np.float64)
            p[f"bn{j}_beta"] = np.array(bn["beta"], dtype=np.float64)
            st["moving"][f"bn{j}"] = [np.array(bn["mean"], dtype=np.float64), np.array(bn["var"], dtype=np.float64)]
    p["out_kernel"] = np.array(w["out_kernel"], dtype=np.float64).reshape(-1)
    p["out_bias"] = np.array(w["out_bias"], dtype=np.float64).reshape(1)
    att = w["attention"]
    p["att_k1"] = np.array(att["k1"], dtype=np.float64)
    p["att_b1"] = np.array(att["b1"], dtype=np.float64)
    p["att_k2"] = np.array(att["k2"], dtype=np.float64).reshape(-1)
    p["att_b2"] = np.array(att["b2"], dtype=np.float64).reshape(1)
    st["params"] = p
    st["m"] = {k: np.zeros_like(v) for k, v in p.items()}
    st["v"] = {k: np.zeros_like(v) for k, v in p.items()}
    return st


def forward_backward(st, spec, users, items, seqs, lens, sparse, dense, labels):
    """Returns (loss, logits, grads, batch BN statistics, attention output)."""
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in st["params"].items()}
    R = len(users)
    parts = [t["item_embeds"]]
    n_rows = t["item_embeds"].shape[0]
    if spec.get("item_sparse_unique") is not None and len(spec["item_sparse_col_index"]):
        parts.append(t["sparse_embeds"][torch.as_tensor(np.asarray(spec["item_sparse_unique"], dtype=np.int64))].reshape(n_rows, -1))
    if spec.get("item_dense_unique") is not None and len(spec["item_dense_col_index"]):
        vals = torch.tensor(np.asarray(spec["item_dense_unique"]), dtype=torch.float64)
        parts.append((vals[:, :, None] * t["dense_embeds"][list(spec["item_dense_col_index"])][None]).reshape(n_rows, -1))
    G = torch.cat(parts, dim=1)
    Kp = G.shape[1]
    q = G[torch.as_tensor(items)]
    keys = G[torch.as_tensor(np.asarray(seqs, dtype=np.int64))]
    T = keys.shape[1]
    qq = q[:, None, :].expand(-1, T, -1)
    feat = torch.cat([qq, keys, qq - keys, qq * keys], dim=2)
    h = torch.sigmoid(feat @ t["att_k1"] + t["att_b1"])
    a = (h @ t["att_k2"] + t["att_b2"][0]) * (1.0 / np.sqrt(Kp))
    mask = torch.arange(T)[None, :] < torch.as_tensor(np.asarray(lens)).reshape(-1, 1)
    a = torch.where(mask, a, torch.full_like(a, -(2.0 ** 32) + 1))
    p_att = torch.softmax(a, dim=1)
    att_out = (p_att[:, :, None] * keys).sum(1)
    xs = [t["user_embeds"][torch.as_tensor(users)], t["item_embeds"][torch.as_tensor(items)]]
    if sparse is not None:
        xs.append(t["sparse_embeds"][torch.as_tensor(sparse)].reshape(R, -1))
    if dense is not None:
        x_d = torch.tensor(np.asarray(dense), dtype=torch.float64)
        xs.append((x_d[:, :, None] * t["dense_embeds"][None]).reshape(R, -1))
    xs.append(att_out)
    act = torch.cat(xs, dim=1)
    stats = {}

    def bn(z, j):
        mu, var = z.mean(0), z.var(0, unbiased=False)
        stats[f"bn{j}"] = (mu.detach().numpy(), var.detach().numpy())
        return (z - mu) / torch.sqrt(var + BN_EPS) * t[f"bn{j}_gamma"] + t[f"bn{j}_beta"]

    if st["use_bn"]:
        act = bn(act, 0)
    n = st["n_layers"]
    for i in range(n):
        act = act @ t[f"W{i}"] + t[f"b{i}"]
        if i != n - 1:
            act = torch.relu(act)
            if st["use_bn"]:
                act = bn(act, i + 1)
    out = act @ t["out_kernel"] + t["out_bias"][0]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.tensor(np.asarray(labels), dtype=torch.float64))
    loss.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(st["params"][k])) for k, v in t.items()}
    return float(loss.detach()), out.detach().numpy(), g, stats, att_out.detach().numpy()


def train_step(st, spec, users, items, seqs, lens, sparse, dense, labels, lr, eps=1e-5):
    loss, _, g, stats, _ = forward_backward(st, spec, users, items, seqs, lens, sparse, dense, labels)
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
