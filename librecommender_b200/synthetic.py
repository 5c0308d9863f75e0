"""Synthetic feature layouts and glorot-style random weights in the reference's conventions (one shared
sparse table with per-field offsets and an OOV slot per field, unique tables with an extra OOV row;
``libreco/feature/sparse.py:106-119``, ``data/data_info.py:399-413``).  Pure data generators — used by the
tests, by ``bench.py --config ...`` and by the profiling drivers; no model arithmetic lives here."""
from __future__ import annotations

import numpy as np


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
