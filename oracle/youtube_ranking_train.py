"""CPU restatement of ONE YouTubeRanking training step of the reference (TEST INFRASTRUCTURE ONLY).

Follows ``libreco/algorithms/youtube_ranking.py:167-218`` (graph with ``is_training=True``: concat(user embedding,
item embedding, ``seq_embeds_pooling`` of the behaviour sequence — ``libreco/layers/embedding.py:54-85``: pad
row zeroed, sum over T, div_no_nan by sqrt(len) —, sparse embeddings, dense value x embedding) ->
``dense_nn`` -> Dense(1)), ``libreco/tfops/loss.py:14-18`` (mean sigmoid CE) and
``libreco/training/tf_trainer.py:112-123`` (TF-Adam + BN update ops).  Forward in torch float64, gradients from
torch autograd; TensorFlow conventions as in ``oracle/fm_train.py`` — **PARITY UNPINNED** for those.
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
    st["params"] = p
    st["m"] = {k: np.zeros_like(v) for k, v in p.items()}
    st["v"] = {k: np.zeros_like(v) for k, v in p.items()}
    return st


def forward_backward(st, users, items, seqs, lens, n_items, sparse, dense, labels):
    """Returns (loss, logits, grads, batch BN statistics)."""
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in st["params"].items()}
    R = len(users)
    E = t["item_embeds"]
    mask = torch.ones(E.shape[0], 1, dtype=torch.float64)
    mask[n_items] = 0.0                                        # the pad row reads as zero (embedding.py:66-72)
    pooled = (E * mask)[torch.as_tensor(np.asarray(seqs, dtype=np.int64))].sum(1)
    ln = torch.sqrt(torch.tensor(np.asarray(lens), dtype=torch.float64)).reshape(-1, 1)
    pooled = torch.where(ln > 0, pooled / torch.clamp(ln, min=1e-30), torch.zeros_like(pooled))
    parts = [t["user_embeds"][torch.as_tensor(users)], t["item_embeds"][torch.as_tensor(items)], pooled]
    if sparse is not None:
        parts.append(t["sparse_embeds"][torch.as_tensor(sparse)].reshape(R, -1))
    if dense is not None:
        x = torch.tensor(np.asarray(dense), dtype=torch.float64)
        parts.append((x[:, :, None] * t["dense_embeds"][None]).reshape(R, -1))
    a = torch.cat(parts, dim=1)
    stats = {}

    def bn(a, j):
        mu, var = a.mean(0), a.var(0, unbiased=False)
        stats[f"bn{j}"] = (mu.detach().numpy(), var.detach().numpy())
        return (a - mu) / torch.sqrt(var + BN_EPS) * t[f"bn{j}_gamma"] + t[f"bn{j}_beta"]

    if st["use_bn"]:
        a = bn(a, 0)
    n = st["n_layers"]
    for i in range(n):
        a = a @ t[f"W{i}"] + t[f"b{i}"]
        if i != n - 1:
            a = torch.relu(a)
            if st["use_bn"]:
                a = bn(a, i + 1)
    out = a @ t["out_kernel"] + t["out_bias"][0]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.tensor(np.asarray(labels), dtype=torch.float64))
    loss.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(st["params"][k])) for k, v in t.items()}
    return float(loss.detach()), out.detach().numpy(), g, stats


def train_step(st, users, items, seqs, lens, n_items, sparse, dense, labels, lr, eps=1e-5):
    loss, _, g, stats = forward_backward(st, users, items, seqs, lens, n_items, sparse, dense, labels)
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
