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
# synthetic specs / weights for the tests (glorot-uniform like the reference's initialisers)
# ----------------------------------------------------------------------------------------------
def make_spec(rng, n_users, n_items, user_sparse_sizes, item_sparse_sizes, n_user_dense, n_item_dense,
              interleave=True):
    """Feature layout in the reference's convention: one shared sparse table with per-field offsets
    and an OOV slot at the end of each field; unique tables carry an extra OOV row."""
    fs = len(user_sparse_sizes) + len(item_sparse_sizes)
    order = list(rng.permutation(fs)) if interleave else list(range(fs))
    ucol = sorted(order[: len(user_sparse_sizes)])
    icol = sorted(order[len(user_sparse_sizes):])
    sizes = {}
    for j, f in enumerate(ucol):
        sizes[f] = user_sparse_sizes[j]
    for j, f in enumerate(icol):
        sizes[f] = item_sparse_sizes[j]
    offsets, off = {}, 0
    for f in range(fs):
        offsets[f] = off
        off += sizes[f] + 1                       # + OOV slot
    def uniq(n_rows, cols):
        t = np.zeros((n_rows + 1, len(cols)), dtype=np.int32)
        for j, f in enumerate(cols):
            t[:n_rows, j] = offsets[f] + rng.integers(0, sizes[f], size=n_rows)
            t[n_rows, j] = offsets[f] + sizes[f]  # OOV row -> the field's oov index
        return t
    fd = n_user_dense + n_item_dense
    dorder = list(rng.permutation(fd)) if interleave else list(range(fd))
    udc = sorted(dorder[:n_user_dense])
    idc = sorted(dorder[n_user_dense:])
    spec = dict(
        n_users=n_users, n_items=n_items, n_sparse=fs, n_dense=fd, sparse_vocab=off,
        user_sparse_col_index=ucol, item_sparse_col_index=icol,
        user_dense_col_index=udc, item_dense_col_index=idc,
        user_sparse_unique=uniq(n_users, ucol) if ucol else None,
        item_sparse_unique=uniq(n_items, icol) if icol else None,
        user_dense_unique=rng.standard_normal((n_users + 1, len(udc))).astype(np.float32) if udc else None,
        item_dense_unique=rng.standard_normal((n_items + 1, len(idc))).astype(np.float32) if idc else None,
    )
    return spec


def make_multi_sparse_spec(rng, n_users, n_items, user_sparse_sizes, item_sparse_sizes, groups,
                           n_user_dense=1, n_item_dense=1, pad_frac=0.3):
    """Layout with multi-sparse fields in the reference's convention (feature/multi_sparse.py:73-95,
    feature/sparse.py:106-119): plain sparse columns first, then every multi-sparse field's
    sub-columns consecutively; the sub-columns of one field share one vocabulary range and one OOV
    slot (= the padding value of missing sub-features).  `groups` = [(side, vocab, length), ...]."""
    spec = make_spec(rng, n_users, n_items, user_sparse_sizes, item_sparse_sizes, n_user_dense, n_item_dense,
                     interleave=False)
    fs0 = spec["n_sparse"]
    off = spec["sparse_vocab"]
    ucol, icol = list(spec["user_sparse_col_index"]), list(spec["item_sparse_col_index"])
    uu = [spec["user_sparse_unique"]] if ucol else []
    iu = [spec["item_sparse_unique"]] if icol else []
    f_off, f_len, f_oov = [], [], []
    col = fs0
    for side, vocab, ln in groups:
        n_rows = n_users if side == "user" else n_items
        oov = off + vocab
        t = off + rng.integers(0, vocab, size=(n_rows + 1, ln))
        t[rng.random((n_rows + 1, ln)) < pad_frac] = oov          # padded (missing) sub-features
        t[n_rows, :] = oov                                          # OOV row
        t[: min(3, n_rows), :] = oov                                # rows with no feature at all -> div_no_nan
        (uu if side == "user" else iu).append(t.astype(np.int32))
        (ucol if side == "user" else icol).extend(range(col, col + ln))
        f_off.append(col); f_len.append(ln); f_oov.append(oov)
        col += ln
        off += vocab + 1
    spec.update(n_sparse=col, sparse_vocab=off, user_sparse_col_index=ucol, item_sparse_col_index=icol,
                user_sparse_unique=np.concatenate(uu, axis=1) if uu else None,
                item_sparse_unique=np.concatenate(iu, axis=1) if iu else None,
                multi_sparse_combine_info=dict(field_offset=f_off, field_len=f_len, feat_oov=np.array(f_oov)))
    return spec


def _glorot(rng, shape):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], 1)
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _rand_bn(rng, n):
    return dict(gamma=rng.uniform(0.5, 1.5, n).astype(np.float32), beta=rng.normal(0, 0.1, n).astype(np.float32),
                mean=rng.normal(0, 0.1, n).astype(np.float32), var=rng.uniform(0.5, 1.5, n).astype(np.float32))


def make_mlp(rng, din, hidden, use_bn):
    dims = [din] + list(hidden)
    mlp = dict(kernels=[_glorot(rng, (dims[i], dims[i + 1])) for i in range(len(hidden))],
               biases=[rng.normal(0, 0.05, dims[i + 1]).astype(np.float32) for i in range(len(hidden))])
    if use_bn:
        mlp["bn_in"] = _rand_bn(rng, din)
        mlp["bns"] = [_rand_bn(rng, dims[i + 1]) for i in range(len(hidden) - 1)]
    return mlp


def make_embeddings(rng, spec, K, linear):
    w = dict(user_embeds=_glorot(rng, (spec["n_users"] + 1, K)), item_embeds=_glorot(rng, (spec["n_items"] + 1, K)))
    if spec["n_sparse"]:
        w["sparse_embeds"] = _glorot(rng, (spec["sparse_vocab"], K))
    if spec["n_dense"]:
        w["dense_embeds"] = _glorot(rng, (spec["n_dense"], K))
    if linear:
        w["user_linear"] = _glorot(rng, (spec["n_users"] + 1, 1)).reshape(-1)
        w["item_linear"] = _glorot(rng, (spec["n_items"] + 1, 1)).reshape(-1)
        if spec["n_sparse"]:
            w["sparse_linear"] = rng.uniform(-0.05, 0.05, spec["sparse_vocab"]).astype(np.float32)
        if spec["n_dense"]:
            w["dense_linear"] = rng.uniform(-0.5, 0.5, spec["n_dense"]).astype(np.float32)
    return w


def make_fm_weights(rng, spec, K, use_bn=True):
    w = make_embeddings(rng, spec, K, linear=True)
    F = 2 + spec["n_sparse"] + spec["n_dense"]
    w.update(lin_kernel=_glorot(rng, (F, 1)).reshape(-1), lin_bias=np.float32(0.03),
             pw_kernel=_glorot(rng, (K, 1)).reshape(-1), pw_bias=np.float32(-0.02))
    if use_bn:
        w["fm_bn"] = _rand_bn(rng, K)
    return w


def make_deepfm_weights(rng, spec, K, hidden=(128, 64, 32), use_bn=True):
    w = make_embeddings(rng, spec, K, linear=True)
    F = 2 + spec["n_sparse"] + spec["n_dense"]
    w.update(lin_kernel=_glorot(rng, (F, 1)).reshape(-1), lin_bias=np.float32(0.01),
             mlp=make_mlp(rng, F * K, hidden, use_bn),
             out_kernel=_glorot(rng, (1 + K + hidden[-1], 1)).reshape(-1), out_bias=np.float32(0.05))
    return w


def make_seq_weights(rng, spec, K, hidden=(64, 32), use_bn=True, din=True):
    w = make_embeddings(rng, spec, K, linear=False)
    F = 2 + spec["n_sparse"] + spec["n_dense"]
    if din:
        Kp = K * (1 + len(spec["item_sparse_col_index"]) + len(spec["item_dense_col_index"]))
        w["attention"] = dict(k1=_glorot(rng, (4 * Kp, 16)), b1=rng.normal(0, 0.05, 16).astype(np.float32),
                              k2=_glorot(rng, (16, 1)).reshape(-1), b2=np.float32(0.02))
        din_w = F * K + Kp
    else:
        din_w = (F + 1) * K
    w["mlp"] = make_mlp(rng, din_w, hidden, use_bn)
    w["out_kernel"] = _glorot(rng, (hidden[-1], 1)).reshape(-1)
    w["out_bias"] = np.float32(-0.01)
    return w


def make_two_tower_weights(rng, spec, K, hidden=(64, 32), use_bn=True):
    w = make_embeddings(rng, spec, K, linear=False)
    w["item_embeds"] = w["item_embeds"][: spec["n_items"]]          # two_tower.py:266-271: no OOV row
    nu = 1 + len(spec["user_sparse_col_index"]) + len(spec["user_dense_col_index"])
    ni = 1 + len(spec["item_sparse_col_index"]) + len(spec["item_dense_col_index"])
    w["user_tower"] = make_mlp(rng, nu * K, hidden, use_bn)
    w["item_tower"] = make_mlp(rng, ni * K, hidden, use_bn)
    w["user_dense_cols"] = list(spec["user_dense_col_index"])
    w["item_dense_cols"] = list(spec["item_dense_col_index"])
    return w
