"""Numpy restatement of the reference's TensorFlow-graph models (inference forward).
TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  TensorFlow is not installed in the build container nor on the GPU box, so
the reference graphs cannot be executed and the reference's own tests hold no numeric golden
values for them (they assert ranges / invariants only, ``tests/utils_pred.py:6-27``).  These
functions follow the graph definitions line by line and are cross-checked in float64
(``dtype=np.float64``) — they are what the CUDA kernels are compared with, but they are not
themselves verified against a TensorFlow run.

Graphs restated (reference @ 7463d9d):
* FM            ``libreco/algorithms/fm.py:140-255``
* DeepFM        ``libreco/algorithms/deepfm.py:143-264``
* TwoTower      ``libreco/algorithms/two_tower.py:306-410`` (towers + optional L2 norm)
* YouTubeRanking``libreco/algorithms/youtube_ranking.py:167-248``
* DIN           ``libreco/algorithms/din.py:165-250`` (default ``din_attention``)
* shared pieces ``libreco/layers/dense.py:12-80`` (dense_nn / tf_dense),
  ``libreco/layers/embedding.py:4-85``, ``libreco/layers/attention.py:28-64``,
  ``libreco/tfops/features.py:6-236``; BN in inference mode = TF defaults
  (``tf.layers.batch_normalization``: epsilon 1e-3, moving statistics).

Weights are plain dicts of numpy arrays; the keys are this repo's own short names (see
``librecommender_b200/feat_models.py::WEIGHT_KEYS``) — the mapping from TF variable names of a
saved reference model lives in the product loader.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3


def _bn(x, bn):
    """tf.layers.batch_normalization(training=False)."""
    if bn is None:
        return x
    g, b, m, v = bn["gamma"], bn["beta"], bn["mean"], bn["var"]
    return (x - m) / np.sqrt(v + x.dtype.type(BN_EPS)) * g + b


def dense_nn(x, mlp):
    """dense.py:12-49: BN(x) -> [Dense -> ReLU -> BN] x (L-1) -> Dense (no activation)."""
    x = _bn(x, mlp.get("bn_in"))
    n = len(mlp["kernels"])
    for i in range(n):
        x = x @ mlp["kernels"][i] + mlp["biases"][i]
        if i != n - 1:
            x = np.maximum(x, 0)
            x = _bn(x, mlp["bns"][i] if mlp.get("bns") else None)
    return x


def field_index(spec, side_users, side_items, kind):
    """Which side / column feeds field f (libreco/prediction/preprocess.py:42-57 _extract_feats:
    columns keep the order of data_info.sparse_col / dense_col)."""
    ucol = spec[f"user_{kind}_col_index"]
    icol = spec[f"item_{kind}_col_index"]
    n = len(ucol) + len(icol)
    side = np.zeros(n, dtype=np.int32)
    col = np.zeros(n, dtype=np.int32)
    for f in range(n):
        if f in ucol:
            side[f], col[f] = 0, ucol.index(f)
        else:
            side[f], col[f] = 1, icol.index(f)
    return side, col


def row_features(spec, users, items):
    """get_original_feats (prediction/preprocess.py:15-57): [R, F_s] int indices, [R, F_d] values."""
    users = np.asarray(users)
    items = np.asarray(items)
    sparse = dense = None
    if spec["n_sparse"]:
        side, col = field_index(spec, users, items, "sparse")
        sparse = np.empty((len(users), spec["n_sparse"]), dtype=np.int64)
        for f in range(spec["n_sparse"]):
            src = spec["user_sparse_unique"][users, col[f]] if side[f] == 0 else spec["item_sparse_unique"][items, col[f]]
            sparse[:, f] = src
    if spec["n_dense"]:
        side, col = field_index(spec, users, items, "dense")
        dense = np.empty((len(users), spec["n_dense"]), dtype=np.float32)
        for f in range(spec["n_dense"]):
            src = spec["user_dense_unique"][users, col[f]] if side[f] == 0 else spec["item_dense_unique"][items, col[f]]
            dense[:, f] = src
    return sparse, dense


def multi_sparse_combine(table, sparse, info, combiner):
    """multi_sparse_combine_embedding + multi_sparse_alone (tfops/features.py:47-118).
    `table` [V, K] or [V]; `sparse` [R, F_raw] raw global indices; returns [R, F', K] / [R, F'] with
    F' = sparse_end + n_fields.  The field's OOV row counts as the zero vector (:94-100), mean / sqrtn
    divide by the number of non-OOV sub-features with div_no_nan (:105-116)."""
    one_d = table.ndim == 1
    T = table.reshape(len(table), -1)
    offs, lens, oovs = info["field_offset"], info["field_len"], info["feat_oov"]
    out = []
    if offs[0] > 0:
        out.append(T[sparse[:, :offs[0]]])
    for off, ln, oov in zip(offs, lens, oovs):
        idx = sparse[:, off:off + ln]
        e = T[idx] * (idx != oov)[:, :, None].astype(T.dtype)
        r = e.sum(axis=1, keepdims=True)
        if combiner in ("mean", "sqrtn"):
            cnt = (idx != oov).sum(axis=1).astype(T.dtype).reshape(-1, 1, 1)
            if combiner == "sqrtn":
                cnt = np.sqrt(cnt)
            r = np.divide(r, cnt, out=np.zeros_like(r), where=cnt != 0)
        out.append(r)
    res = np.concatenate(out, axis=1)
    return res[:, :, 0] if one_d else res


def _stacked_embeds(w, users, items, sparse, dense, dtype):
    """[R, F, K] field embeddings and [R, F] linear features (fm.py:174-255)."""
    P = [w["user_embeds"][users][:, None, :], w["item_embeds"][items][:, None, :]]
    lin = []
    if "user_linear" in w:
        lin = [w["user_linear"][users].reshape(-1, 1), w["item_linear"][items].reshape(-1, 1)]
    ms = w.get("multi_sparse")        # {"field_offset", "field_len", "feat_oov", "combiner"}
    if sparse is not None and ms is not None and ms["combiner"] in ("sum", "mean", "sqrtn"):
        P.append(multi_sparse_combine(w["sparse_embeds"], sparse, ms, ms["combiner"]))
        if "sparse_linear" in w:
            lin.append(multi_sparse_combine(w["sparse_linear"], sparse, ms, ms["combiner"]))
    elif sparse is not None:
        P.append(w["sparse_embeds"][sparse])
        if "sparse_linear" in w:
            lin.append(w["sparse_linear"][sparse])
    if dense is not None:
        P.append(dense[:, :, None].astype(dtype) * w["dense_embeds"][None, :, :])   # features.py:131-145
        if "dense_linear" in w:
            lin.append(dense.astype(dtype) * w["dense_linear"][None, :])
    P = np.concatenate(P, axis=1).astype(dtype)
    L = np.concatenate(lin, axis=1).astype(dtype) if lin else None
    return P, L


def _cast(w, dtype):
    def c(x):
        if isinstance(x, dict):
            return {k: c(v) for k, v in x.items()}
        if isinstance(x, list):
            return [c(v) for v in x]
        if isinstance(x, np.ndarray) and x.dtype.kind == "f":
            return x.astype(dtype)
        return x
    return c(w)


def fm_forward(w, users, items, sparse=None, dense=None, dtype=np.float32):
    """fm.py:152-171 — logits."""
    w = _cast(w, dtype)
    P, L = _stacked_embeds(w, users, items, sparse, dense, dtype)
    linear_term = L @ w["lin_kernel"].reshape(-1, 1) + w["lin_bias"]
    pw = 0.5 * (np.square(P.sum(axis=1)) - np.square(P).sum(axis=1))
    pw = _bn(pw, w.get("fm_bn"))
    z = pw @ w["pw_kernel"].reshape(-1, 1) + w["pw_bias"]
    z = np.where(z > 0, z, np.expm1(z))          # elu
    return (linear_term + z).reshape(-1)


def deepfm_forward(w, users, items, sparse=None, dense=None, dtype=np.float32):
    """deepfm.py:155-174 — logits."""
    w = _cast(w, dtype)
    P, L = _stacked_embeds(w, users, items, sparse, dense, dtype)
    linear_term = L @ w["lin_kernel"].reshape(-1, 1) + w["lin_bias"]
    pw = 0.5 * (np.square(P.sum(axis=1)) - np.square(P).sum(axis=1))
    deep = dense_nn(P.reshape(len(P), -1), w["mlp"])
    cat = np.concatenate([linear_term, pw, deep], axis=1)
    return (cat @ w["out_kernel"].reshape(-1, 1) + w["out_bias"]).reshape(-1)


def youtube_retrieval_user_vectors(w, spec, user_ids, seqs, lens, norm=False, dtype=np.float32):
    """youtube_retrieval.py:169-260 + dyn_embed_base.py:281-313 — user embeddings (without the pseudo bias):
    dense_nn(concat(sqrtn pooling of seq_embeds_var over the user's recent items, user sparse embeddings, user
    dense value x embedding))."""
    E = np.asarray(w["seq_embeds"], dtype=dtype)
    n_items = E.shape[0]
    Ez = np.concatenate([E, np.zeros((1, E.shape[1]), dtype=dtype)], axis=0)        # pad id n_items -> zero row
    pooled = Ez[np.asarray(seqs)[user_ids]].sum(axis=1)
    ln = np.sqrt(np.asarray(lens, dtype=dtype)[user_ids]).reshape(-1, 1)
    pooled = np.divide(pooled, ln, out=np.zeros_like(pooled), where=ln != 0)
    parts = [pooled]
    if len(spec["user_sparse_col_index"]):
        parts.append(np.asarray(w["sparse_embeds"], dtype=dtype)[spec["user_sparse_unique"][user_ids]].reshape(len(user_ids), -1))
    if len(spec["user_dense_col_index"]):
        cols = list(spec["user_dense_col_index"])
        parts.append((spec["user_dense_unique"][user_ids][:, :, None].astype(dtype)
                      * np.asarray(w["dense_embeds"], dtype=dtype)[cols][None]).reshape(len(user_ids), -1))
    v = dense_nn(np.concatenate(parts, axis=1).astype(dtype), _cast(w["mlp"], dtype))
    if norm:
        v = v / np.linalg.norm(v, axis=1, keepdims=True)
    return v


def wide_deep_forward(wd, users, items, sparse=None, dense=None, dtype=np.float32):
    """wide_deep.py:150-176 — logits.  ``wd``: the reference's variables (user_wide [n+1], ..., wide_kernel [F],
    wide_bias, user_deep [n+1, K], ..., mlp, deep_kernel [H], deep_bias)."""
    c = lambda x: np.asarray(x, dtype=dtype)      # noqa: E731
    wide = [c(wd["user_wide"])[users][:, None], c(wd["item_wide"])[items][:, None]]
    deep = [c(wd["user_deep"])[users][:, None, :], c(wd["item_deep"])[items][:, None, :]]
    if sparse is not None:
        wide.append(c(wd["sparse_wide"])[sparse])
        deep.append(c(wd["sparse_deep"])[sparse])
    if dense is not None:
        x = dense.astype(dtype)
        wide.append(x * c(wd["dense_wide"])[None, :])
        deep.append(x[:, :, None] * c(wd["dense_deep"])[None, :, :])
    wide = np.concatenate(wide, axis=1)
    deep = np.concatenate(deep, axis=1)
    wide_term = wide @ c(wd["wide_kernel"]).reshape(-1, 1) + dtype(wd["wide_bias"])
    h = dense_nn(deep.reshape(len(users), -1), _cast(wd["mlp"], dtype))
    deep_term = h @ c(wd["deep_kernel"]).reshape(-1, 1) + dtype(wd["deep_bias"])
    return (wide_term + deep_term).reshape(-1)


def tower_forward(w, ids, sparse, dense, which, norm, dtype=np.float32):
    """two_tower.py:306-346,400-410 — one tower over `ids` with that side's features."""
    w = _cast(w, dtype)
    parts = [w[f"{which}_embeds"][ids]]
    if sparse is not None:
        parts.append(w["sparse_embeds"][sparse].reshape(len(ids), -1))
    if dense is not None:
        cols = w[f"{which}_dense_cols"]
        parts.append((dense[:, :, None].astype(dtype) * w["dense_embeds"][cols][None, :, :]).reshape(len(ids), -1))
    x = np.concatenate(parts, axis=1).astype(dtype)
    v = dense_nn(x, w[f"{which}_tower"])
    if norm:
        v = v / np.sqrt(np.maximum(np.square(v).sum(axis=1, keepdims=True), dtype(1e-12)))
    return v


def seq_pool(item_embeds, seqs, lens, n_items, dtype=np.float32):
    """embedding.py:54-85 — pad row (index n_items) forced to zero, sum over T, / sqrt(len)."""
    E = item_embeds.astype(dtype).copy()
    E[n_items] = 0
    s = E[seqs].sum(axis=1)
    ln = np.sqrt(np.asarray(lens, dtype=dtype)).reshape(-1, 1)
    return np.divide(s, ln, out=np.zeros_like(s), where=ln != 0)


def youtube_ranking_forward(w, users, items, seqs, lens, n_items, sparse=None, dense=None, dtype=np.float32):
    """youtube_ranking.py:201-216 — logits."""
    w = _cast(w, dtype)
    parts = [w["user_embeds"][users], w["item_embeds"][items],
             seq_pool(w["item_embeds"], seqs, lens, n_items, dtype)]
    if sparse is not None:
        parts.append(w["sparse_embeds"][sparse].reshape(len(users), -1))
    if dense is not None:
        parts.append((dense[:, :, None].astype(dtype) * w["dense_embeds"][None]).reshape(len(users), -1))
    x = np.concatenate(parts, axis=1).astype(dtype)
    h = dense_nn(x, w["mlp"])
    return (h @ w["out_kernel"].reshape(-1, 1) + w["out_bias"]).reshape(-1)


def item_feature_table(w, spec, dtype=np.float32):
    """combine_seq_features (tfops/features.py:165-218, concat mode): for every item j in [0, N]
    the row [E_i[j] || flatten(E_s[item_sparse_unique[j]]) || flatten(item_dense_unique[j] * E_d[cols])]."""
    w = _cast(w, dtype)
    parts = [w["item_embeds"]]
    if spec.get("item_sparse_unique") is not None and len(spec["item_sparse_col_index"]):
        parts.append(w["sparse_embeds"][spec["item_sparse_unique"]].reshape(len(w["item_embeds"]), -1))
    if spec.get("item_dense_unique") is not None and len(spec["item_dense_col_index"]):
        cols = spec["item_dense_col_index"]
        parts.append((spec["item_dense_unique"][:, :, None].astype(dtype) * w["dense_embeds"][cols][None]).reshape(
            len(w["item_embeds"]), -1))
    return np.concatenate(parts, axis=1).astype(dtype)


def din_attention(q, keys, lens, att, dtype=np.float32):
    """attention.py:45-64: Dense1(sigmoid(Dense16([q,k,q-k,q*k]))) * rsqrt(K'), mask -2^32+1, softmax."""
    B, T, Kp = keys.shape
    qq = np.repeat(q[:, None, :], T, axis=1)
    feat = np.concatenate([qq, keys, qq - keys, qq * keys], axis=2).astype(dtype)
    h = feat @ att["k1"] + att["b1"]
    h = 1.0 / (1.0 + np.exp(-h))
    a = (h @ att["k2"].reshape(-1, 1) + att["b2"]).reshape(B, T)
    a = a * dtype(1.0 / np.sqrt(Kp))
    mask = np.arange(T)[None, :] < np.asarray(lens).reshape(-1, 1)
    a = np.where(mask, a, dtype(-(2 ** 32) + 1))
    a = a - a.max(axis=1, keepdims=True)
    p = np.exp(a)
    p = p / p.sum(axis=1, keepdims=True)
    return (p[:, :, None] * keys).sum(axis=1)


def tf_attention(q, keys, lens, dtype=np.float32):
    """attention.py:5-25 = tf.keras.layers.Attention(use_scale=False)([q[:, None], keys], mask=[None, m])
    (third-party: Keras dot-product attention): scores <q, k_t>, masked positions -= 1e9, softmax over t,
    output sum_t p_t k_t."""
    B, T, _ = keys.shape
    a = np.einsum("bk,btk->bt", q.astype(dtype), keys.astype(dtype))
    mask = np.arange(T)[None, :] < np.asarray(lens).reshape(-1, 1)
    a = a - dtype(1.0e9) * (~mask)
    a = a - a.max(axis=1, keepdims=True)
    p = np.exp(a)
    p = p / p.sum(axis=1, keepdims=True)
    return (p[:, :, None] * keys).sum(axis=1)


def din_forward(w, spec, users, items, seqs, lens, sparse=None, dense=None, dtype=np.float32):
    """din.py:182-250 — logits; ``w["use_tf_attention"]`` selects tf_attention (din.py:247-248)."""
    use_tf = bool(w.get("use_tf_attention", False))
    w = _cast({k: v for k, v in w.items() if k != "use_tf_attention"}, dtype)
    G = item_feature_table(w, spec, dtype)
    if use_tf:
        att_out = tf_attention(G[items], G[seqs], lens, dtype)
    else:
        att_out = din_attention(G[items], G[seqs], lens, w["attention"], dtype)
    parts = [w["user_embeds"][users], w["item_embeds"][items]]
    if sparse is not None:
        parts.append(w["sparse_embeds"][sparse].reshape(len(users), -1))
    if dense is not None:
        parts.append((dense[:, :, None].astype(dtype) * w["dense_embeds"][None]).reshape(len(users), -1))
    parts.append(att_out)
    x = np.concatenate(parts, axis=1).astype(dtype)
    h = dense_nn(x, w["mlp"])
    return (h @ w["out_kernel"].reshape(-1, 1) + w["out_bias"]).reshape(-1)


# ----------------------------------------------------------------------------------------------
# synthetic specs / weights for the tests: pure data generators, shared with the benches
# ----------------------------------------------------------------------------------------------
from librecommender_b200.synthetic import (  # noqa: E402,F401
    _glorot, _rand_bn, make_deepfm_weights, make_embeddings, make_fm_weights, make_mlp, make_multi_sparse_spec,
    make_seq_weights, make_spec, make_two_tower_weights)
