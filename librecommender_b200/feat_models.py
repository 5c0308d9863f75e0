"""Inference engines for the reference's feature ("TfBase") models on the GPU: FM, DeepFM
(``libreco/algorithms/fm.py``, ``deepfm.py``) — forward (``predict``) and all-items scoring
(``recommend_user`` = ``recommend_tf_feat``, ``libreco/recommendation/recommend.py:81-105``).

* The per-row feed the reference materialises on the host (``process_tf_feat`` /
  ``get_original_feats``) is never built: ``b200_feat_forward`` reads the per-user / per-item
  unique feature tables in the kernel, and for all-items scoring rows are the implicit grid
  (user, item 0..N-1).
* BatchNorm layers are folded into scale/shift (inference = moving statistics, TF defaults
  epsilon 1e-3) and, inside ``dense_nn``, into the following Dense layer.
* Weights arrive as a dict with the keys of ``WEIGHT_KEYS`` (numpy arrays) — see
  ``from_tf_variables`` for the mapping from a reference ``*_tf_variables.npz``.
"""
from __future__ import annotations

import ctypes
from ctypes import Structure, c_int32, c_int64, c_void_p

import numpy as np

from . import _lib
from .consumed import ConsumedCSR, as_csr

MAX_FIELDS = 128
BN_EPS = 1e-3

WEIGHT_KEYS = (
    "user_embeds", "item_embeds", "sparse_embeds", "dense_embeds",
    "user_linear", "item_linear", "sparse_linear", "dense_linear",
    "lin_kernel", "lin_bias", "pw_kernel", "pw_bias", "fm_bn", "mlp", "out_kernel", "out_bias",
)


class FeatLayoutStruct(Structure):
    _fields_ = [
        ("embed_size", c_int32), ("n_sparse", c_int32), ("n_dense", c_int32),
        ("id_mask", c_int32), ("dense_embed_row", c_int32 * MAX_FIELDS),
        ("sparse_side", c_int32 * MAX_FIELDS), ("sparse_col", c_int32 * MAX_FIELDS),
        ("dense_side", c_int32 * MAX_FIELDS), ("dense_col", c_int32 * MAX_FIELDS),
        ("user_sparse_unique", c_void_p), ("ld_us", c_int64),
        ("item_sparse_unique", c_void_p), ("ld_is", c_int64),
        ("user_dense_unique", c_void_p), ("ld_ud", c_int64),
        ("item_dense_unique", c_void_p), ("ld_id", c_int64),
        ("sparse_rows", c_void_p), ("ld_sparse_rows", c_int64),
        ("dense_rows", c_void_p), ("ld_dense_rows", c_int64),
    ]


class FeatTablesStruct(Structure):
    _fields_ = [(n, c_void_p) for n in ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds",
                                        "user_linear", "item_linear", "sparse_linear", "dense_linear")]


def _dev(x, device, dtype):
    import torch

    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype).contiguous()


# Dense-layer kernel selection: "auto" sends layers that are compute-bound on the SIMT kernel
# (din >= TC_MIN_DIN, enough rows to fill the SMs) to the tcgen05 3xTF32 kernel; both are this
# library's own CUDA kernels and both meet the 1e-5 bar.  "f32" / "tf32x3" force one (tests).
LINEAR_IMPL = "auto"
TC_MIN_DIN = 64
TC_MIN_ROWS = 4096
TC_MIN_MACS = 1 << 27
DIN_FUSED_ATTENTION = False   # DIN all-items: sigmoid / Dense(1) fused into the attention GEMM's epilogue.  Measured
# (profiles/r02_launches_din_*.csv): no gain — the fused GEMM 524 us + softmax kernel 242 us per user vs 359 + 395 us
# for the [N, 16 len] round trip; the per-tile cost of the K = 64 product dominates either way.  Kept as a tested variant.
TC_LONG_K = 1024        # long reductions (weight gradients over the batch) starve the SIMT kernel's few CTAs


_SPLIT_CACHE = {}


def split_weights(Wt):
    """hi / lo tf32 copies of a layer's weights for b200_linear_tf32x3, made once per weight tensor
    (keyed by storage pointer + shape + version counter, so in-place updates re-split)."""
    import torch

    key = (Wt.data_ptr(), tuple(Wt.shape), Wt.stride(0), Wt._version)
    hit = _SPLIT_CACHE.get(key)
    if hit is None:
        dout, din = Wt.shape
        ld = int(_lib.lib.b200_linear_tf32x3_split_ld(din))
        buf = torch.empty(2 * dout * ld, dtype=torch.float32, device=Wt.device)
        _lib.check(_lib.lib.b200_linear_tf32x3_split_weights(_lib.ptr(Wt), Wt.stride(0), din, dout, _lib.ptr(buf),
                                                             _lib.current_stream()))
        if len(_SPLIT_CACHE) > 256:
            _SPLIT_CACHE.clear()
        hit = _SPLIT_CACHE[key] = (buf, Wt)      # keep Wt alive so the pointer key stays unique
    return hit[0]


def linear(x, Wt, b, relu, cache_split=True):
    """tf_dense (libreco/layers/dense.py:52-80) with BN folded: act(x Wt^T + b), fp32 device tensors."""
    import torch

    R, din, dout = x.shape[0], Wt.shape[1], Wt.shape[0]
    y = torch.empty((R, dout), dtype=torch.float32, device=x.device)
    bp = _lib.ptr(b) if b is not None else None
    aligned = x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.stride(1) == 1 and Wt.stride(1) == 1
    # few output rows but a long reduction (the weight gradients dWt = dY^T X of the training steps: 128 x 1792
    # outputs over 8192 rows) would run on a handful of SIMT CTAs: send those to the tensor-core kernel too
    use_tc = LINEAR_IMPL == "tf32x3" or (LINEAR_IMPL == "auto" and din >= TC_MIN_DIN and
                                         (R >= TC_MIN_ROWS or din >= TC_LONG_K or R * din * dout >= TC_MIN_MACS))
    w_ok = cache_split or (Wt.stride(0) % 4 == 0 and Wt.data_ptr() % 16 == 0)
    if use_tc and aligned and w_ok and not cache_split and din >= TC_LONG_K:
        # few output tiles, long reduction: split the reduction over enough CTAs to fill the SMs
        tiles = -(-R // 128) * -(-dout // 128)
        splits = max(1, min(16, 148 // tiles, din // 256))
        if splits > 1:
            part = torch.empty(splits * R * dout, dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib.b200_linear_tf32x3_splitk(_lib.ptr(x), x.stride(0), R, _lib.ptr(Wt), Wt.stride(0), bp, din,
                                                          dout, 1 if relu else 0, splits, _lib.ptr(part),
                                                          part.numel() * 4, _lib.ptr(y), y.stride(0),
                                                          _lib.current_stream()))
            return y
    if use_tc and aligned and w_ok:
        ws = split_weights(Wt) if cache_split else None
        _lib.check(_lib.lib.b200_linear_tf32x3(_lib.ptr(x), x.stride(0), R, _lib.ptr(Wt), Wt.stride(0), _lib.ptr(ws), bp,
                                               din, dout, 1 if relu else 0, _lib.ptr(y), y.stride(0),
                                               _lib.current_stream()))
    else:
        _lib.check(_lib.lib.b200_linear_f32(_lib.ptr(x), x.stride(0), R, _lib.ptr(Wt), Wt.stride(0), bp, din, dout,
                                            1 if relu else 0, _lib.ptr(y), y.stride(0), _lib.current_stream()))
    return y


def _side_cols(user_cols, item_cols):
    n = len(user_cols) + len(item_cols)
    side, col = [0] * n, [0] * n
    for f in range(n):
        if f in user_cols:
            side[f], col[f] = 0, list(user_cols).index(f)
        else:
            side[f], col[f] = 1, list(item_cols).index(f)
    return side, col


def fold_bn(bn):
    """BN(x) = x * scale + shift at inference."""
    if bn is None:
        return None, None
    scale = (bn["gamma"] / np.sqrt(bn["var"] + np.float32(BN_EPS))).astype(np.float32)
    shift = (bn["beta"] - bn["mean"] * scale).astype(np.float32)
    return scale, shift


def fold_mlp(mlp):
    """dense_nn (libreco/layers/dense.py:12-49) -> [(Wt [dout, din], bias, relu)], BN folded into
    the Dense that FOLLOWS it: Dense(BN(a)) = a (diag(s) W) + (t W + b)."""
    layers = []
    scale, shift = fold_bn(mlp.get("bn_in"))
    n = len(mlp["kernels"])
    for i in range(n):
        W = np.asarray(mlp["kernels"][i], dtype=np.float32)
        b = np.asarray(mlp["biases"][i], dtype=np.float32)
        if scale is not None:
            b = (shift @ W + b).astype(np.float32)
            W = (scale[:, None] * W).astype(np.float32)
        layers.append((np.ascontiguousarray(W.T), b, i != n - 1))
        scale, shift = (None, None)
        if i != n - 1 and mlp.get("bns"):
            scale, shift = fold_bn(mlp["bns"][i])
    return layers


_COMBINERS = {"sum": 0, "mean": 1, "sqrtn": 2}


def _spec_get(spec):
    """Uniform read access to a layout description: a plain dict (tests, fixtures) or the reference's
    ``DataInfo`` object (libreco/data/data_info.py:107-290), whose column indices live in
    ``data_info.user_sparse_col.index`` etc. (`Feature(name, index)` tuples, :209-247)."""
    if isinstance(spec, dict):
        return spec.get

    def g(k, d=None):
        if k.endswith("_col_index") and not hasattr(spec, k):
            feat = getattr(spec, k[: -len("_index")], None)
            return list(feat.index) if feat is not None and getattr(feat, "index", None) is not None else d
        v = getattr(spec, k, d)
        return d if v is None else v

    return g


def combine_multi_sparse(spec, weights, combiner="sqrtn", device=None):
    """multi_sparse_combine_embedding (libreco/tfops/features.py:47-118) hoisted out of the row path.

    A multi-sparse field (e.g. genre1..genre3 sharing one vocabulary and one OOV slot) is a function
    of the user or of the item only, so its pooled embedding / pooled linear weight is computed ONCE
    per unique-table row (``b200_multi_sparse_combine``) and appended to the shared sparse tables;
    the field then is an ordinary single-index field for every downstream kernel.  Returns
    ``(spec_dict, weights_dict)`` in the reduced field layout the reference's graph uses after
    combining (``true_sparse_field_size``, feature/multi_sparse.py:146-157): plain sparse fields
    first, then one field per multi-sparse group.  With ``combiner="normal"`` or no multi-sparse info
    the inputs are returned unchanged."""
    import torch

    g = _spec_get(spec)
    info = g("multi_sparse_combine_info")
    if info is None or combiner not in _COMBINERS:
        return spec, weights
    ig = info.get if isinstance(info, dict) else (lambda k, d=None: getattr(info, k, d))
    offs, lens, oovs = list(ig("field_offset")), list(ig("field_len")), [int(x) for x in ig("feat_oov")]
    device = device if device is not None else _lib.require_cuda()
    ucol, icol = list(g("user_sparse_col_index") or []), list(g("item_sparse_col_index") or [])
    sparse_end = offs[0]
    E = _dev(weights["sparse_embeds"], device, torch.float32)
    K = E.shape[1]
    lin = _dev(weights.get("sparse_linear"), device, torch.float32)
    uniq = {"user": _dev(g("user_sparse_unique"), device, torch.int32) if ucol else None,
            "item": _dev(g("item_sparse_unique"), device, torch.int32) if icol else None}
    cols = {"user": ucol, "item": icol}
    new_cols = {"user": [c for c in ucol if c < sparse_end], "item": [c for c in icol if c < sparse_end]}
    new_uniq = {w: [uniq[w][:, [cols[w].index(c) for c in new_cols[w]]]] if new_cols[w] else [] for w in ("user", "item")}
    add_e, add_l, base = [], [], E.shape[0]
    for gi, (off, ln, oov) in enumerate(zip(offs, lens, oovs)):
        which = "user" if off in ucol else "item"
        members = list(range(off, off + ln))
        if any(c not in cols[which] for c in members):
            raise ValueError(f"multi-sparse field at offset {off} straddles the user / item sides")
        idx = uniq[which][:, [cols[which].index(c) for c in members]].contiguous()
        n = idx.shape[0]
        ce = torch.empty((n, K), dtype=torch.float32, device=device)
        _lib.check(_lib.lib.b200_multi_sparse_combine(_lib.ptr(E), E.stride(0), K, _lib.ptr(idx), idx.stride(0), ln, n,
                                                      oov, _COMBINERS[combiner], _lib.ptr(ce), ce.stride(0),
                                                      _lib.current_stream()))
        add_e.append(ce)
        if lin is not None:
            cl = torch.empty(n, dtype=torch.float32, device=device)
            _lib.check(_lib.lib.b200_multi_sparse_combine(_lib.ptr(lin), 1, 1, _lib.ptr(idx), idx.stride(0), ln, n, oov,
                                                          _COMBINERS[combiner], _lib.ptr(cl), 1, _lib.current_stream()))
            add_l.append(cl)
        new_uniq[which].append((base + torch.arange(n, device=device, dtype=torch.int32)).view(-1, 1))
        new_cols[which].append(sparse_end + gi)
        base += n
    out_spec = {k: g(k) for k in ("n_users", "n_items", "user_dense_col_index", "item_dense_col_index",
                                  "user_dense_unique", "item_dense_unique")}
    out_spec.update(user_sparse_col_index=new_cols["user"], item_sparse_col_index=new_cols["item"],
                    user_sparse_unique=torch.cat(new_uniq["user"], dim=1).contiguous() if new_uniq["user"] else None,
                    item_sparse_unique=torch.cat(new_uniq["item"], dim=1).contiguous() if new_uniq["item"] else None,
                    n_sparse=sparse_end + len(offs), n_dense=g("n_dense"), multi_sparse_combine_info=None)
    out_w = dict(weights)
    out_w["sparse_embeds"] = torch.cat([E] + add_e, dim=0)
    if lin is not None:
        out_w["sparse_linear"] = torch.cat([lin] + add_l, dim=0)
    return out_spec, out_w


class FeatSpec:
    """Device-resident feature layout (from the reference's DataInfo or an equivalent dict)."""

    def __init__(self, spec, embed_size, device=None):
        import torch

        self.device = device if device is not None else _lib.require_cuda()
        g = _spec_get(spec)
        self.n_users, self.n_items = int(g("n_users")), int(g("n_items"))
        ucol, icol = list(g("user_sparse_col_index") or []), list(g("item_sparse_col_index") or [])
        udc, idc = list(g("user_dense_col_index") or []), list(g("item_dense_col_index") or [])
        self.n_sparse, self.n_dense = len(ucol) + len(icol), len(udc) + len(idc)
        if self.n_sparse > MAX_FIELDS or self.n_dense > MAX_FIELDS:
            raise ValueError("too many feature fields")
        self.us = _dev(g("user_sparse_unique"), self.device, torch.int32) if ucol else None
        self.is_ = _dev(g("item_sparse_unique"), self.device, torch.int32) if icol else None
        self.ud = _dev(g("user_dense_unique"), self.device, torch.float32) if udc else None
        self.id_ = _dev(g("item_dense_unique"), self.device, torch.float32) if idc else None
        L = FeatLayoutStruct()
        L.embed_size, L.n_sparse, L.n_dense = int(embed_size), self.n_sparse, self.n_dense
        L.id_mask = 3
        for f in range(self.n_dense):
            L.dense_embed_row[f] = f
        side, col = _side_cols(ucol, icol)
        for f in range(self.n_sparse):
            L.sparse_side[f], L.sparse_col[f] = side[f], col[f]
        side, col = _side_cols(udc, idc)
        for f in range(self.n_dense):
            L.dense_side[f], L.dense_col[f] = side[f], col[f]
        for name, t in (("user_sparse_unique", self.us), ("item_sparse_unique", self.is_),
                        ("user_dense_unique", self.ud), ("item_dense_unique", self.id_)):
            setattr(L, name, t.data_ptr() if t is not None else None)
        L.ld_us = self.us.stride(0) if self.us is not None else 0
        L.ld_is = self.is_.stride(0) if self.is_ is not None else 0
        L.ld_ud = self.ud.stride(0) if self.ud is not None else 0
        L.ld_id = self.id_.stride(0) if self.id_ is not None else 0
        L.sparse_rows, L.dense_rows = None, None
        self.layout = L
        self.item_sparse_cols, self.item_dense_cols = icol, idc
        self.user_sparse_cols, self.user_dense_cols = ucol, udc

    def with_rows(self, sparse_rows, dense_rows):
        """Layout copy that reads explicit per-row features (predict with given feature rows)."""
        L = FeatLayoutStruct.from_buffer_copy(self.layout)
        if sparse_rows is not None:
            L.sparse_rows, L.ld_sparse_rows = sparse_rows.data_ptr(), sparse_rows.stride(0)
        if dense_rows is not None:
            L.dense_rows, L.ld_dense_rows = dense_rows.data_ptr(), dense_rows.stride(0)
        return L


class _FeatModelBase:
    """Shared machinery: device tables, row forward, all-items scoring + masked top-K."""

    needs_linear = True

    def __init__(self, spec, weights, user_consumed=None, task="ranking", device=None):
        import torch

        self._torch = torch
        K = int(weights["user_embeds"].shape[1])
        if not isinstance(spec, FeatSpec):   # multi-sparse fields pooled once (default combiner as the reference's)
            spec, weights = combine_multi_sparse(spec, weights, weights.get("multi_sparse_combiner", "sqrtn"), device)
        self.spec = spec if isinstance(spec, FeatSpec) else FeatSpec(spec, K, device)
        self.device = self.spec.device
        self.K = K
        self.task = task
        self.n_users, self.n_items = self.spec.n_users, self.spec.n_items
        self.F = 2 + self.spec.n_sparse + self.spec.n_dense
        f32 = torch.float32
        self.t = {k: _dev(weights.get(k), self.device, f32) for k in
                  ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds",
                   "user_linear", "item_linear", "sparse_linear", "dense_linear")}
        T = FeatTablesStruct()
        for k, v in self.t.items():
            setattr(T, k, v.data_ptr() if v is not None else None)
        self.tables = T
        if self.needs_linear:
            self.lin_kernel = _dev(np.asarray(weights["lin_kernel"]).reshape(-1), self.device, f32)
            self.lin_bias = float(np.asarray(weights["lin_bias"]).reshape(-1)[0])
        csr = user_consumed if user_consumed is not None else ConsumedCSR(
            np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32))
        self.csr = as_csr(csr, self.n_users)
        self.indptr_d, self.idx_d = self.csr.device(self.device)

    # -- to be provided by subclasses ---------------------------------------------------------
    def _forward(self, layout, users_d, items_d, R, grid_items):
        raise NotImplementedError

    # -- hoisted all-items scoring: one-sided partial sums -----------------------
    def _side(self, which):
        """(layout restricted to the fields of one side, their positions in the global field order)."""
        cache = self.__dict__.setdefault("_side_cache", {})
        if which not in cache:
            sp = self.spec
            L = FeatLayoutStruct.from_buffer_copy(sp.layout)
            L.id_mask = 1 if which == "user" else 2
            scols = sp.user_sparse_cols if which == "user" else sp.item_sparse_cols
            dcols = sp.user_dense_cols if which == "user" else sp.item_dense_cols
            L.n_sparse, L.n_dense = len(scols), len(dcols)
            for f in range(len(scols)):
                L.sparse_side[f], L.sparse_col[f] = (0 if which == "user" else 1), f
            for f in range(len(dcols)):
                L.dense_side[f], L.dense_col[f] = (0 if which == "user" else 1), f
                L.dense_embed_row[f] = dcols[f]
            pos = [0 if which == "user" else 1] + [2 + c for c in scols] + [2 + sp.n_sparse + c for c in dcols]
            cache[which] = (L, pos)
        return cache[which]

    def _side_partials(self, which, ids_d, want_concat):
        """S = sum_f e, Q = sum_f e^2, linear partial (no bias) and optionally the concatenated
        embeddings of ONE side for the given ids."""
        torch = self._torch
        L, pos = self._side(which)
        n = int(ids_d.numel())
        S = torch.empty((n, self.K), dtype=torch.float32, device=self.device)
        Q = torch.empty((n, self.K), dtype=torch.float32, device=self.device)
        lin = torch.empty(n, dtype=torch.float32, device=self.device)
        concat = torch.empty((n, len(pos) * self.K), dtype=torch.float32, device=self.device) if want_concat else None
        lk = self.__dict__.setdefault("_side_lin", {})
        if which not in lk:
            lk[which] = self.lin_kernel[torch.as_tensor(pos, device=self.device)].contiguous()
        self._feat_forward(L, ids_d, ids_d, n, 0, concat=concat, lin=lin, ssum=S, sqsum=Q,
                           lin_kernel=lk[which], lin_bias=0.0)
        return S, Q, lin, concat

    # -- public ------------------------------------------------------------------------------------
    def logits(self, users, items, sparse_rows=None, dense_rows=None):
        torch = self._torch
        u = torch.as_tensor(np.asarray(users, dtype=np.int64)).to(self.device)
        i = torch.as_tensor(np.asarray(items, dtype=np.int64)).to(self.device)
        layout = self.spec.layout
        if sparse_rows is not None or dense_rows is not None:
            sr = _dev(sparse_rows, self.device, torch.int32)
            dr = _dev(dense_rows, self.device, torch.float32)
            layout = self.spec.with_rows(sr, dr)
            self._keep = (sr, dr)
        return self._forward(layout, u, i, u.numel(), 0)

    def predict(self, users, items):
        """predict_tf_feat + normalize_prediction (prediction/predict.py:18-33,43-92), known ids."""
        z = self.logits(users, items)
        if self.task == "ranking":
            z = self._torch.sigmoid(z)
        return z.cpu().numpy()

    def score_all_items(self, user_ids_d):
        """[b, n_items] logits of every (user, item) pair — rows are the implicit grid."""
        b = int(user_ids_d.numel())
        out = self._forward(self.spec.layout, user_ids_d, None, b * self.n_items, self.n_items)
        return out.view(b, self.n_items)

    def recommend(self, user_ids, n_rec, filter_consumed=True, return_scores=False, rows_per_chunk=None):
        """recommend_tf_feat (recommend.py:81-105): all-items scoring, consumed filter, top-K."""
        torch = self._torch
        if n_rec > self.n_items:
            raise ValueError(f"`n_rec` {n_rec} exceeds num of items {self.n_items}")
        uid = torch.as_tensor(np.asarray(user_ids, dtype=np.int64)).to(self.device)
        B = uid.numel()
        if rows_per_chunk is None:   # users per chunk: bound the [users, n_items] score matrix to ~1 GiB
            rows_per_chunk = max(1, min(B, (1 << 28) // max(self.n_items, 1)))
        out_ids = torch.empty((B, n_rec), dtype=torch.int64, device=self.device)
        out_sc = torch.empty((B, n_rec), dtype=torch.float32, device=self.device)
        lib, stream = _lib.lib, _lib.current_stream()
        for r0 in range(0, B, rows_per_chunk):
            u = uid[r0:r0 + rows_per_chunk]
            b = u.numel()
            scores = self.score_all_items(u).contiguous()
            if filter_consumed and self.csr.nnz > 0:
                _lib.check(lib.b200_mask_consumed(_lib.ptr(scores), scores.stride(0), _lib.ptr(u), b,
                                                  self.n_items, n_rec, _lib.ptr(self.indptr_d),
                                                  _lib.ptr(self.idx_d), self.csr.n_users, stream))
            nbytes = ctypes.c_size_t(0)
            _lib.check(lib.b200_topk_rows_workspace_bytes(b, self.n_items, n_rec, ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            _lib.check(lib.b200_topk_rows(_lib.ptr(scores), scores.stride(0), b, self.n_items, n_rec,
                                          _lib.ptr(out_ids[r0:r0 + b]), _lib.ptr(out_sc[r0:r0 + b]),
                                          _lib.ptr(ws), nbytes.value, stream))
        ids = out_ids.cpu().numpy()
        if return_scores:
            sc = out_sc.cpu().numpy()
            return ids, (1.0 / (1.0 + np.exp(-sc)) if self.task == "ranking" else sc)
        return ids

    def max_grid_rows(self):
        return 1 << 22

    def default_recs(self, n_rec=2000):
        """Top-``min(n_rec, n_items)`` for the OOV user without the consumed filter — what
        ``TfBase.fit`` stores as ``default_recs`` for cold-start users (``bases/tf_base.py:145-153``)."""
        return self.recommend([self.n_users], min(int(n_rec), self.n_items), filter_consumed=False).flatten()

    def recommend_dynamic(self, user_id, n_rec, data_info, user_feats=None, seq=None, filter_consumed=True,
                          inner_id=False, return_scores=False):
        """``recommend_tf_feat`` for ONE user with features / behaviour sequence supplied for this call
        (``recommendation/recommend.py:39-54,81-105``, ``recommendation/preprocess.py:104-159``): the
        user's feature columns are overridden in a per-row feature matrix of the N (user, item) rows
        (``dynamic_feats.dynamic_feature_rows``; the device tables are not touched), a sequence model
        reads the supplied sequence instead of the cached one."""
        torch = self._torch
        from .dynamic_feats import build_rec_seq, dynamic_feature_rows

        if getattr(self, "has_multi_sparse", False) and user_feats:
            raise NotImplementedError("feature overrides on layouts with multi-sparse fields")
        if n_rec > self.n_items:
            raise ValueError(f"`n_rec` {n_rec} exceeds num of items {self.n_items}")
        u = int(user_id)
        uid = torch.tensor([u], dtype=torch.int64, device=self.device)
        N = self.n_items
        layout = self.spec.layout
        keep = None
        if user_feats:
            sp, de = dynamic_feature_rows(data_info, u, user_feats)
            sr = _dev(sp, self.device, torch.int32)
            dr = _dev(de, self.device, torch.float32)
            layout = self.spec.with_rows(sr, dr)
            keep = (sr, dr)
        restore = None
        if seq is not None and len(seq) > 0 and hasattr(self, "seqs"):
            row, ln = build_rec_seq(seq, N, self.T, getattr(data_info, "item2id", None), inner_id)
            restore = (self.seqs[u].clone(), self.lens[u].clone())
            self.seqs[u] = torch.from_numpy(row[0]).to(self.device)
            self.lens[u] = int(ln[0])
        try:
            if user_feats:          # explicit per-row features: the flat (user, item) grid
                scores = self._forward(layout, uid, None, N, N).view(1, N).contiguous()
            else:
                scores = self.score_all_items(uid).contiguous()
        finally:
            if restore is not None:
                self.seqs[u], self.lens[u] = restore
        del keep
        lib, stream = _lib.lib, _lib.current_stream()
        if filter_consumed and self.csr.nnz > 0:
            _lib.check(lib.b200_mask_consumed(_lib.ptr(scores), scores.stride(0), _lib.ptr(uid), 1, N, n_rec,
                                              _lib.ptr(self.indptr_d), _lib.ptr(self.idx_d), self.csr.n_users, stream))
        out_ids = torch.empty((1, n_rec), dtype=torch.int64, device=self.device)
        out_sc = torch.empty((1, n_rec), dtype=torch.float32, device=self.device)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.b200_topk_rows_workspace_bytes(1, N, n_rec, ctypes.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        _lib.check(lib.b200_topk_rows(_lib.ptr(scores), scores.stride(0), 1, N, n_rec, _lib.ptr(out_ids),
                                      _lib.ptr(out_sc), _lib.ptr(ws), nbytes.value, stream))
        ids = out_ids.cpu().numpy()
        if return_scores:
            sc = out_sc.cpu().numpy()
            return ids, (1.0 / (1.0 + np.exp(-sc)) if self.task == "ranking" else sc)
        return ids

    def assign_oov(self, sparse_oov=None):
        """``assign_tf_variables_oov`` (``bases/tf_base.py:310-353``) on the device tables, in place."""
        torch = self._torch
        with torch.no_grad():
            for name, n in (("user_embeds", self.n_users), ("user_linear", self.n_users),
                            ("item_embeds", self.n_items), ("item_linear", self.n_items)):
                v = self.t.get(name)
                if v is not None and v.shape[0] > n:
                    v[n] = v[:n].mean(dim=0)
            if sparse_oov is not None:
                for name in ("sparse_embeds", "sparse_linear"):
                    v = self.t.get(name)
                    if v is None:
                        continue
                    start = 0
                    for oov in [int(o) for o in sparse_oov]:
                        if start >= oov:
                            continue
                        v[oov] = v[start:oov].mean(dim=0)
                        start = oov + 1
        for k in ("_item_part", "_item_side"):               # cached item-side partials depend on the tables
            self.__dict__.pop(k, None)
        if hasattr(self, "_rebuild_item_features"):
            self._rebuild_item_features()

    # -- helpers -----------------------------------------------------------------------------------
    def _feat_forward(self, layout, users_d, items_d, R, grid_items, concat=None, pw=None, lin=None,
                      fm_out=None, head=None, row_offset=0, ssum=None, sqsum=None, lin_kernel=None, lin_bias=None):
        head = head or {}
        lk = lin_kernel if lin_kernel is not None else (self.lin_kernel if self.needs_linear else None)
        lb = lin_bias if lin_bias is not None else (self.lin_bias if self.needs_linear else 0.0)
        _lib.check(_lib.lib.b200_feat_forward(
            ctypes.byref(layout), ctypes.byref(self.tables), _lib.ptr(users_d), _lib.ptr(items_d), R,
            grid_items, row_offset, _lib.ptr(concat), concat.stride(0) if concat is not None else 0,
            _lib.ptr(pw), pw.stride(0) if pw is not None else 0, _lib.ptr(lin), _lib.ptr(fm_out),
            _lib.ptr(lk), float(lb),
            _lib.ptr(head.get("bn_scale")), _lib.ptr(head.get("bn_shift")), _lib.ptr(head.get("pw_kernel")),
            float(head.get("pw_bias", 0.0)), _lib.ptr(ssum), _lib.ptr(sqsum),
            ssum.stride(0) if ssum is not None else 0, _lib.current_stream()))

    def _mlp(self, x, layers):
        torch = self._torch
        for Wt, b, relu in layers:
            x = linear(x, Wt, b, relu)
        return x

    def _upload_mlp(self, mlp):
        torch = self._torch
        return [(_dev(Wt, self.device, torch.float32), _dev(b, self.device, torch.float32), relu)
                for Wt, b, relu in fold_mlp(mlp)]


class FM(_FeatModelBase):
    """libreco/algorithms/fm.py:140-172 (inference)."""

    def __init__(self, spec, weights, user_consumed=None, task="ranking", device=None):
        super().__init__(spec, weights, user_consumed, task, device)
        torch = self._torch
        scale, shift = fold_bn(weights.get("fm_bn"))
        self.head = dict(bn_scale=_dev(scale, self.device, torch.float32),
                         bn_shift=_dev(shift, self.device, torch.float32),
                         pw_kernel=_dev(np.asarray(weights["pw_kernel"]).reshape(-1), self.device, torch.float32),
                         pw_bias=float(np.asarray(weights["pw_bias"]).reshape(-1)[0]))

    def _forward(self, layout, users_d, items_d, R, grid_items):
        out = self._torch.empty(R, dtype=self._torch.float32, device=self.device)
        self._feat_forward(layout, users_d, items_d, R, grid_items, fm_out=out, head=self.head)
        return out

    def score_all_items(self, user_ids_d):
        """Hoisted: item-side sums once per model, user-side sums once per call, K adds per pair."""
        torch = self._torch
        if self.K > 64:
            return super().score_all_items(user_ids_d)
        if "_item_side" not in self.__dict__:
            ids = torch.arange(self.n_items, device=self.device)
            self._item_side = self._side_partials("item", ids, False)[:3]
        Si, Qi, li = self._item_side
        Su, Qu, lu, _ = self._side_partials("user", user_ids_d, False)
        b = int(user_ids_d.numel())
        scores = torch.empty((b, self.n_items), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.b200_fm_pair_scores(
            _lib.ptr(Su), _lib.ptr(Qu), _lib.ptr(lu), b, _lib.ptr(Si), _lib.ptr(Qi), _lib.ptr(li), self.n_items,
            self.K, self.lin_bias, _lib.ptr(self.head["bn_scale"]), _lib.ptr(self.head["bn_shift"]),
            _lib.ptr(self.head["pw_kernel"]), self.head["pw_bias"], _lib.ptr(scores), scores.stride(0),
            _lib.current_stream()))
        return scores

    def max_grid_rows(self):
        return 1 << 28


class DeepFM(_FeatModelBase):
    """libreco/algorithms/deepfm.py:143-175 (inference)."""

    def __init__(self, spec, weights, user_consumed=None, task="ranking", device=None):
        super().__init__(spec, weights, user_consumed, task, device)
        torch = self._torch
        self.mlp = self._upload_mlp(weights["mlp"])
        self.hidden_last = self.mlp[-1][0].shape[0]
        self.out_kernel = _dev(np.asarray(weights["out_kernel"]).reshape(-1), self.device, torch.float32)
        self.out_bias = float(np.asarray(weights["out_bias"]).reshape(-1)[0])

    def _forward(self, layout, users_d, items_d, R, grid_items):
        torch = self._torch
        out = torch.empty(R, dtype=torch.float32, device=self.device)
        step = self.max_grid_rows()
        for r0 in range(0, R, step):                  # bound the [rows, F*K] deep input
            r1 = min(R, r0 + step)
            n = r1 - r0
            concat = torch.empty((n, self.F * self.K), dtype=torch.float32, device=self.device)
            pw = torch.empty((n, self.K), dtype=torch.float32, device=self.device)
            lin = torch.empty(n, dtype=torch.float32, device=self.device)
            if grid_items > 0:
                self._feat_forward(layout, users_d, None, n, grid_items, concat=concat, pw=pw, lin=lin,
                                   row_offset=r0)
            else:
                self._feat_forward(layout, users_d[r0:r1], items_d[r0:r1], n, 0, concat=concat, pw=pw, lin=lin)
            deep = self._mlp(concat, self.mlp)
            lin2 = lin.view(n, 1)
            _lib.check(_lib.lib.b200_concat_dense(
                _lib.ptr(lin2), 1, 1, _lib.ptr(pw), pw.stride(0), self.K, _lib.ptr(deep), deep.stride(0),
                self.hidden_last, _lib.ptr(self.out_kernel), self.out_bias, n, _lib.ptr(out[r0:r1]),
                _lib.current_stream()))
        return out

    def _hoistable(self):
        n = len(self.mlp)
        dims = [w.shape[0] for w, _, _ in self.mlp]
        return self.K <= 64 and n in (2, 3) and dims[0] <= 256 and dims[1] <= 64 and (n == 2 or dims[2] <= 32)

    def _first_layer_partial(self, which, concat, with_bias):
        torch = self._torch
        cache = self.__dict__.setdefault("_w1_side", {})
        if which not in cache:
            _, pos = self._side(which)
            cols = torch.cat([torch.arange(g * self.K, (g + 1) * self.K, device=self.device) for g in pos])
            cache[which] = self.mlp[0][0][:, cols].contiguous()          # Wt [H1, F_side*K], BN already folded
        Wt = cache[which]
        return linear(concat, Wt, self.mlp[0][1] if with_bias else None, False)

    def score_all_items(self, user_ids_d):
        """Hoisted: first-layer partial products per side, only the small layers per (user, item)."""
        torch = self._torch
        if not self._hoistable():
            return super().score_all_items(user_ids_d)
        if "_item_side" not in self.__dict__:
            ids = torch.arange(self.n_items, device=self.device)
            Si, Qi, li, ci = self._side_partials("item", ids, True)
            self._item_side = (Si, Qi, li, self._first_layer_partial("item", ci, False))
        Si, Qi, li, Pi = self._item_side
        Su, Qu, lu, cu = self._side_partials("user", user_ids_d, True)
        Pu = self._first_layer_partial("user", cu, True)
        W2t, b2, _ = self.mlp[1]
        three = len(self.mlp) == 3
        if "_tail" not in self.__dict__:
            self._tail = (W2t.t().contiguous(), self.mlp[2][0].t().contiguous() if three else None)
        W2, W3 = self._tail
        b = int(user_ids_d.numel())
        scores = torch.empty((b, self.n_items), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.b200_deepfm_pair_scores(
            _lib.ptr(Su), _lib.ptr(Qu), _lib.ptr(lu), _lib.ptr(Pu), b, _lib.ptr(Si), _lib.ptr(Qi), _lib.ptr(li),
            _lib.ptr(Pi), self.n_items, self.K, Pu.shape[1], W2.shape[1], W3.shape[1] if three else 0,
            self.lin_bias, _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(W3), _lib.ptr(self.mlp[2][1]) if three else None,
            _lib.ptr(self.out_kernel), self.out_bias, _lib.ptr(scores), scores.stride(0), _lib.current_stream()))
        return scores

    def max_grid_rows(self):
        # deep input bytes per row = F*K*4; keep a chunk under ~1 GiB
        return max(1, (1 << 30) // (self.F * self.K * 4))


def wide_deep_weights(user_wide, item_wide, sparse_wide, dense_wide, wide_kernel, wide_bias, user_deep, item_deep,
                      sparse_deep, dense_deep, mlp, deep_kernel, deep_bias):
    """WideDeep (``libreco/algorithms/wide_deep.py:150-262``) on the DeepFM engine: the wide term
    ``Dense1(concat of the 1-d wide embeddings)`` IS DeepFM's linear term, the deep tower IS DeepFM's, and
    ``output = wide_term + Dense1(deep)`` is DeepFM's head ``<[lin, pw, deep], w> + b`` with weight 1 on ``lin``, 0 on
    the pairwise block and the ``deep_term`` kernel / bias on the rest.  Returns the weight dict :class:`DeepFM` takes
    (variables named as in the reference: ``user_wide_var`` ... ``dense_deep_var``, ``wide_term`` / ``deep_term``)."""
    K = int(np.asarray(user_deep).shape[1])
    w = dict(user_embeds=user_deep, item_embeds=item_deep, user_linear=np.asarray(user_wide).reshape(-1),
             item_linear=np.asarray(item_wide).reshape(-1), lin_kernel=np.asarray(wide_kernel).reshape(-1),
             lin_bias=np.float32(np.asarray(wide_bias).reshape(-1)[0]), mlp=mlp,
             out_kernel=np.concatenate([np.ones(1, np.float32), np.zeros(K, np.float32),
                                        np.asarray(deep_kernel, dtype=np.float32).reshape(-1)]),
             out_bias=np.float32(np.asarray(deep_bias).reshape(-1)[0]))
    if sparse_deep is not None:
        w["sparse_embeds"], w["sparse_linear"] = sparse_deep, np.asarray(sparse_wide).reshape(-1)
    if dense_deep is not None:
        w["dense_embeds"], w["dense_linear"] = dense_deep, np.asarray(dense_wide).reshape(-1)
    return w


def from_tf_variables(npz, names=None):
    """Map a reference ``<name>_tf_variables.npz`` (utils/save_load.py:70-80) to WEIGHT_KEYS.
    Embedding names are fixed by the reference code; the un-named
    ``tf_dense`` heads get TF-version-dependent auto names, so `names` may override the defaults."""
    names = names or {}
    g = lambda k, d: npz[names.get(k, d)] if names.get(k, d) in npz else None
    w = {}
    for k in ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds"):
        v = g(k, f"embedding/{k}_var:0")
        if v is not None:
            w[k] = v
    for k in ("user_linear", "item_linear", "sparse_linear", "dense_linear"):
        v = g(k, f"embedding/{k}_var:0")
        if v is not None:
            w[k] = np.asarray(v).reshape(-1)
    return w


# ==============================================================================================
# sequence models (DIN, YouTubeRanking) and TwoTower
# ==============================================================================================
def recent_sequences(user_consumed, n_users, n_items, max_seq_len):
    """get_recent_seqs (libreco/batch/sequence.py:75-91): last `max_seq_len` consumed items per
    user, padded with n_items; an extra all-pad OOV row with length 1."""
    seqs = np.full((n_users + 1, max_seq_len), n_items, dtype=np.int32)
    lens = np.ones(n_users + 1, dtype=np.int32)
    for u in range(n_users):
        items = user_consumed[u] if u in user_consumed else []
        n = min(len(items), max_seq_len)
        if n:
            seqs[u, :n] = items[-n:] if len(items) >= max_seq_len else items
        lens[u] = n if len(items) < max_seq_len else max_seq_len
    return seqs, lens


def recent_sequences_csr(consumed, n_items, max_seq_len):
    """Vectorised get_recent_seqs (libreco/batch/sequence.py:75-91) over a ConsumedCSR (arrival
    order): no per-user Python loop — 10 M users in seconds instead of minutes."""
    indptr, idx = consumed.indptr, consumed.idx
    n_users = len(indptr) - 1
    clen = np.diff(indptr)
    lens = np.minimum(clen, max_seq_len).astype(np.int32)
    seqs = np.full((n_users + 1, max_seq_len), n_items, dtype=np.int32)
    t = np.arange(max_seq_len, dtype=np.int64)[None, :]
    src = (indptr[1:] - lens)[:, None] + t
    valid = t < lens[:, None]
    seqs[:n_users][valid] = idx[src[valid]]
    return seqs, np.append(lens, np.int32(1)).astype(np.int32)


def permute_mlp_input(mlp, perm):
    """Re-order the input features of a dense_nn (first kernel rows + input BN) by `perm`."""
    out = dict(mlp)
    out["kernels"] = [np.asarray(mlp["kernels"][0])[perm]] + list(mlp["kernels"][1:])
    if mlp.get("bn_in") is not None:
        out["bn_in"] = {k: np.asarray(v)[perm] for k, v in mlp["bn_in"].items()}
    return out


class _SeqModelBase(_FeatModelBase):
    needs_linear = False

    def __init__(self, spec, weights, recent_seqs, recent_seq_lens, user_consumed=None, task="ranking",
                 device=None):
        super().__init__(spec, weights, user_consumed, task, device)
        torch = self._torch
        self.seqs = _dev(recent_seqs, self.device, torch.int32)
        self.lens = _dev(recent_seq_lens, self.device, torch.int32)
        self.T = int(self.seqs.shape[1])
        self.out_kernel = _dev(np.asarray(weights["out_kernel"]).reshape(-1), self.device, torch.float32)
        self.out_bias = float(np.asarray(weights["out_bias"]).reshape(-1)[0])
        self.extra = 0          # width of the sequence block appended to the concatenated row

    def _seq_block(self, layout, users_d, items_d, n, grid_items, row_offset, out_view):
        raise NotImplementedError

    def _forward(self, layout, users_d, items_d, R, grid_items):
        torch = self._torch
        out = torch.empty(R, dtype=torch.float32, device=self.device)
        width = self.F * self.K + self.extra
        step = max(1, (1 << 30) // (width * 4))
        for r0 in range(0, R, step):
            r1 = min(R, r0 + step)
            n = r1 - r0
            x = torch.empty((n, width), dtype=torch.float32, device=self.device)
            if grid_items > 0:
                self._feat_forward(layout, users_d, None, n, grid_items, concat=x, row_offset=r0)
                self._seq_block(users_d, None, n, grid_items, r0, x[:, self.F * self.K:])
            else:
                self._feat_forward(layout, users_d[r0:r1], items_d[r0:r1], n, 0, concat=x)
                self._seq_block(users_d[r0:r1], items_d[r0:r1], n, 0, 0, x[:, self.F * self.K:])
            h = self._mlp(x, self.mlp)
            _lib.check(_lib.lib.b200_concat_dense(
                _lib.ptr(h), h.stride(0), h.shape[1], None, 0, 0, None, 0, 0, _lib.ptr(self.out_kernel),
                self.out_bias, n, _lib.ptr(out[r0:r1]), _lib.current_stream()))
        return out

    def max_grid_rows(self):
        return max(1, (1 << 30) // ((self.F * self.K + self.extra) * 4))


class YouTubeRanking(_SeqModelBase):
    """libreco/algorithms/youtube_ranking.py:167-218 (inference)."""

    def __init__(self, spec, weights, recent_seqs, recent_seq_lens, user_consumed=None, task="ranking",
                 device=None):
        super().__init__(spec, weights, recent_seqs, recent_seq_lens, user_consumed, task, device)
        K, F = self.K, self.F
        self.extra = K
        # reference order [user, item, pooled, sparse.., dense..] -> ours [user, item, sparse.., dense.., pooled]
        perm = np.concatenate([np.arange(0, 2 * K), np.arange(3 * K, (F + 1) * K), np.arange(2 * K, 3 * K)])
        self.mlp = self._upload_mlp(permute_mlp_input(weights["mlp"], perm))

    # ---- hoisted all-items scoring: everything user-only or item-only once ----
    def _hoistable(self):
        dims = [w.shape[0] for w, _, _ in self.mlp]
        n = len(dims)
        return self.K <= 64 and n in (2, 3) and dims[0] <= 256 and dims[1] <= 64 and (n == 2 or dims[2] <= 32)

    def _side_input(self, which, ids_d):
        """Concatenated field embeddings of ONE side for the given ids (+ the pooled history for users)
        and the matching columns of the first MLP layer."""
        torch = self._torch
        L, pos = self._side(which)
        n = int(ids_d.numel())
        width = len(pos) * self.K + (self.K if which == "user" else 0)
        x = torch.empty((n, width), dtype=torch.float32, device=self.device)
        self._feat_forward(L, ids_d, ids_d, n, 0, concat=x)
        if which == "user":
            self._seq_block(ids_d, ids_d, n, 0, 0, x[:, len(pos) * self.K:])
        cache = self.__dict__.setdefault("_w1_side", {})
        if which not in cache:
            groups = list(pos) + ([self.F] if which == "user" else [])     # pooled block sits after the F fields
            cols = torch.cat([torch.arange(g * self.K, (g + 1) * self.K, device=self.device) for g in groups])
            cache[which] = self.mlp[0][0][:, cols].contiguous()
        return x, cache[which]

    def score_all_items(self, user_ids_d):
        """youtube_ranking.py:199-218 over the implicit (user, item) grid: the first Dense layer splits
        into a user part (id, user features, pooled history) and an item part, computed once per user /
        once per item; a pair then costs H1 adds + the small layers (the DeepFM pair kernel with an
        empty FM part)."""
        torch = self._torch
        if not self._hoistable():
            return super().score_all_items(user_ids_d)
        if "_item_part" not in self.__dict__:
            xi, Wi = self._side_input("item", torch.arange(self.n_items, device=self.device))
            self._item_part = linear(xi, Wi, None, False)
            three = len(self.mlp) == 3
            self._tail = (self.mlp[1][0].t().contiguous(), self.mlp[2][0].t().contiguous() if three else None)
            self._w_out = torch.cat([torch.zeros(1 + self.K, dtype=torch.float32, device=self.device),
                                     self.out_kernel]).contiguous()
            self._zeros_i = torch.zeros((self.n_items, self.K + 1), dtype=torch.float32, device=self.device)
        xu, Wu = self._side_input("user", user_ids_d)
        Pu = linear(xu, Wu, self.mlp[0][1], False)
        Pi = self._item_part
        W2, W3 = self._tail
        three = W3 is not None
        b = int(user_ids_d.numel())
        zu = torch.zeros((b, self.K + 1), dtype=torch.float32, device=self.device)
        zi = self._zeros_i
        scores = torch.empty((b, self.n_items), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.b200_deepfm_pair_scores(
            _lib.ptr(zu), _lib.ptr(zu), _lib.ptr(zu), _lib.ptr(Pu), b, _lib.ptr(zi), _lib.ptr(zi), _lib.ptr(zi),
            _lib.ptr(Pi), self.n_items, self.K, Pu.shape[1], W2.shape[1], W3.shape[1] if three else 0, 0.0,
            _lib.ptr(W2), _lib.ptr(self.mlp[1][1]), _lib.ptr(W3), _lib.ptr(self.mlp[2][1]) if three else None,
            _lib.ptr(self._w_out), self.out_bias, _lib.ptr(scores), scores.stride(0), _lib.current_stream()))
        return scores

    def _seq_block(self, users_d, items_d, n, grid_items, row_offset, out_view):
        E = self.t["item_embeds"]
        _lib.check(_lib.lib.b200_seq_pool(
            _lib.ptr(E), E.stride(0), self.K, self.n_items, _lib.ptr(self.seqs), self.seqs.stride(0),
            _lib.ptr(self.lens), self.T, _lib.ptr(users_d), n, grid_items, row_offset,
            _lib.ptr(out_view), out_view.stride(0), _lib.current_stream()))


class DIN(_SeqModelBase):
    """libreco/algorithms/din.py:165-250 (inference; ``use_tf_attention`` either way: with
    ``weights["use_tf_attention"]`` the attention is the weight-free dot-product form of
    ``layers/attention.py:5-25`` and all-items scoring runs on the flat (user, item) grid)."""

    def __init__(self, spec, weights, recent_seqs, recent_seq_lens, user_consumed=None, task="ranking",
                 device=None):
        super().__init__(spec, weights, recent_seqs, recent_seq_lens, user_consumed, task, device)
        torch = self._torch
        self._rebuild_item_features()
        self.Kp = int(self.G.shape[1])
        self.extra = self.Kp
        # use_tf_attention=True (din.py:247-248, layers/attention.py:5-25): dot-product attention, no weights
        self.use_tf_attention = bool(weights.get("use_tf_attention", False)) or weights.get("attention") is None
        if self.use_tf_attention:
            self.att = dict(k1=None, b1=None, k2=None, b2=0.0)
        else:
            att = weights["attention"]
            self.att = dict(k1=_dev(att["k1"], self.device, torch.float32), b1=_dev(att["b1"], self.device, torch.float32),
                            k2=_dev(np.asarray(att["k2"]).reshape(-1), self.device, torch.float32),
                            b2=float(np.asarray(att["b2"]).reshape(-1)[0]))
        self.mlp = self._upload_mlp(weights["mlp"])

    def _rebuild_item_features(self):
        """item feature table G (combine_seq_features, concat mode; tfops/features.py:165-218): built once
        per set of tables (again after ``assign_oov``)."""
        torch = self._torch
        parts = [self.t["item_embeds"]]
        if self.spec.is_ is not None:
            parts.append(self.t["sparse_embeds"][self.spec.is_.long()].reshape(self.n_items + 1, -1))
        if self.spec.id_ is not None:
            cols = torch.as_tensor(self.spec.item_dense_cols, device=self.device)
            parts.append((self.spec.id_[:, :, None] * self.t["dense_embeds"][cols][None]).reshape(self.n_items + 1, -1))
        self.G = torch.cat(parts, dim=1).contiguous()

    # ---- hoisted all-items scoring (SURVEY.md 8d "a7 DIN all-items") ---------------------------------
    def _hoistable(self):
        dims = [w.shape[0] for w, _, _ in self.mlp]
        n = len(dims)
        return (self.K <= 64 and n in (2, 3) and dims[0] <= 256 and dims[1] <= 64 and (n == 2 or dims[2] <= 32)
                and self.Kp % 4 == 0 and not self.use_tf_attention)

    def _side_concat(self, which, ids_d):
        torch = self._torch
        L, pos = self._side(which)
        n = int(ids_d.numel())
        x = torch.empty((n, len(pos) * self.K), dtype=torch.float32, device=self.device)
        self._feat_forward(L, ids_d, ids_d, n, 0, concat=x)
        cache = self.__dict__.setdefault("_w1_side", {})
        if which not in cache:
            cols = torch.cat([torch.arange(g * self.K, (g + 1) * self.K, device=self.device) for g in pos])
            cache[which] = self.mlp[0][0][:, cols].contiguous()
        return x, cache[which]

    def score_all_items(self, user_ids_d):
        """din.py:165-250 over (this user) x (every item).  Per user: the attention's Dense(16) becomes
        one GEMM [N, K'] x [K', 16 len] on the library GEMM kernel (tcgen05 3xTF32 for large N), a warp
        per item finishes sigmoid / Dense(1) / softmax / weighted key sum, the first MLP layer splits
        into user / item / attention parts and the pair kernel runs the small layers."""
        torch = self._torch
        if not self._hoistable():
            return super().score_all_items(user_ids_d)
        N, Kp, FK = self.n_items, self.Kp, self.F * self.K
        if "_item_part" not in self.__dict__:
            xi, Wi = self._side_concat("item", torch.arange(N, device=self.device))
            self._item_part = linear(xi, Wi, None, False)
            self._w_att = self.mlp[0][0][:, FK:FK + Kp].contiguous()           # [H1, K']
            three = len(self.mlp) == 3
            self._tail = (self.mlp[1][0].t().contiguous(), self.mlp[2][0].t().contiguous() if three else None)
            self._w_out = torch.cat([torch.zeros(1 + self.K, dtype=torch.float32, device=self.device),
                                     self.out_kernel]).contiguous()
            self._zeros_i = torch.zeros((N, self.K + 1), dtype=torch.float32, device=self.device)
        W2, W3 = self._tail
        three = W3 is not None
        b = int(user_ids_d.numel())
        xu, Wu = self._side_concat("user", user_ids_d)
        Pu_all = linear(xu, Wu, self.mlp[0][1], False)                           # [b, H1] incl. bias
        scores = torch.empty((b, N), dtype=torch.float32, device=self.device)
        zu = torch.zeros((1, self.K + 1), dtype=torch.float32, device=self.device)
        lens_h = self.lens[user_ids_d].clamp(0, self.T).cpu().numpy()            # one small D2H per call
        Gn = self.G[:N]
        att = torch.empty((N, Kp), dtype=torch.float32, device=self.device)
        lib, st = _lib.lib, _lib.current_stream()
        for r in range(b):
            ln = int(lens_h[r])
            seq = self.seqs[user_ids_d[r]]                                       # int32 [T] view (device)
            Z = None
            fused = False
            if ln > 0:
                Wt = torch.empty((16 * ln, Kp), dtype=torch.float32, device=self.device)
                bias = torch.empty(16 * ln, dtype=torch.float32, device=self.device)
                _lib.check(lib.b200_din_user_weights(_lib.ptr(self.G), self.G.stride(0), Kp, _lib.ptr(seq), ln,
                                                     _lib.ptr(self.att["k1"]), _lib.ptr(self.att["b1"]), _lib.ptr(Wt),
                                                     Wt.stride(0), _lib.ptr(bias), st))
                fused = (DIN_FUSED_ATTENTION and ln <= 64 and ln * Kp * 4 <= 48 * 1024 and Kp % 4 == 0
                         and Gn.stride(0) % 4 == 0 and N >= TC_MIN_ROWS)
                if fused:
                    # sigmoid + Dense(1) in the GEMM's epilogue: [N, ln] logits instead of [N, 16 ln] pre-activations
                    A = torch.empty((N, ln), dtype=torch.float32, device=self.device)
                    _lib.check(lib.b200_linear_tf32x3_sigmoid_dot(
                        _lib.ptr(Gn), Gn.stride(0), N, _lib.ptr(Wt), Wt.stride(0), None, _lib.ptr(bias), Kp, 16 * ln,
                        _lib.ptr(self.att["k2"]), _lib.ptr(A), A.stride(0), st))
                    _lib.check(lib.b200_din_attention_from_logits(
                        _lib.ptr(A), A.stride(0), N, _lib.ptr(self.G), self.G.stride(0), Kp, _lib.ptr(seq), ln,
                        self.att["b2"], _lib.ptr(att), att.stride(0), st))
                else:
                    Z = linear(Gn, Wt, bias, False, cache_split=False)           # [N, 16 ln]
            if not fused:
                _lib.check(lib.b200_din_attention_hoisted(
                    _lib.ptr(Z), Z.stride(0) if Z is not None else 0, N, _lib.ptr(self.G), self.G.stride(0), Kp,
                    _lib.ptr(seq), ln, _lib.ptr(self.att["k2"]), self.att["b2"], _lib.ptr(att), att.stride(0), st))
            Pi = self._item_part + linear(att, self._w_att, None, False)         # [N, H1]
            Pu = Pu_all[r:r + 1]
            _lib.check(lib.b200_deepfm_pair_scores(
                _lib.ptr(zu), _lib.ptr(zu), _lib.ptr(zu), _lib.ptr(Pu), 1, _lib.ptr(self._zeros_i),
                _lib.ptr(self._zeros_i), _lib.ptr(self._zeros_i), _lib.ptr(Pi), N, self.K, Pu.shape[1], W2.shape[1],
                W3.shape[1] if three else 0, 0.0, _lib.ptr(W2), _lib.ptr(self.mlp[1][1]), _lib.ptr(W3),
                _lib.ptr(self.mlp[2][1]) if three else None, _lib.ptr(self._w_out), self.out_bias,
                _lib.ptr(scores[r]), scores.stride(0), st))
        return scores

    def _seq_block(self, users_d, items_d, n, grid_items, row_offset, out_view):
        _lib.check(_lib.lib.b200_din_attention(
            _lib.ptr(self.G), self.G.stride(0), self.Kp, _lib.ptr(items_d), _lib.ptr(self.seqs),
            self.seqs.stride(0), _lib.ptr(self.lens), self.T, _lib.ptr(users_d), n, grid_items, row_offset,
            _lib.ptr(self.att["k1"]), _lib.ptr(self.att["b1"]), _lib.ptr(self.att["k2"]), self.att["b2"],
            _lib.ptr(out_view), out_view.stride(0), _lib.current_stream()))


class TwoTower:
    """libreco/algorithms/two_tower.py:306-346,400-410 + DynEmbedBase.set_embeddings
    (libreco/bases/dyn_embed_base.py:240-269): both towers over ALL users / items on the GPU; the
    results (plus the mean OOV rows of embed_base.py:257-265) feed the embed scorer directly."""

    def __init__(self, spec, weights, norm_embed=False, device=None):
        import torch

        self._torch = torch
        K = int(weights["user_embeds"].shape[1])
        spec, weights = combine_multi_sparse(spec, weights, weights.get("multi_sparse_combiner", "sqrtn"), device)
        self.base = FeatSpec(spec, K, device)
        self.device = self.base.device
        self.K = K
        self.norm_embed = norm_embed
        self.n_users, self.n_items = self.base.n_users, self.base.n_items
        f32 = torch.float32
        self.t = {k: _dev(weights.get(k), self.device, f32) for k in
                  ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds")}
        T = FeatTablesStruct()
        for k, v in self.t.items():
            setattr(T, k, v.data_ptr() if v is not None else None)
        self.tables = T
        self.layouts, self.mlps, self.widths = {}, {}, {}
        for which, mask in (("user", 1), ("item", 2)):
            L = FeatLayoutStruct.from_buffer_copy(self.base.layout)
            L.id_mask = mask
            scols = self.base.user_sparse_cols if which == "user" else self.base.item_sparse_cols
            dcols = self.base.user_dense_cols if which == "user" else self.base.item_dense_cols
            L.n_sparse, L.n_dense = len(scols), len(dcols)
            for f in range(len(scols)):
                L.sparse_side[f], L.sparse_col[f] = (0 if which == "user" else 1), f
            for f in range(len(dcols)):
                L.dense_side[f], L.dense_col[f] = (0 if which == "user" else 1), f
                L.dense_embed_row[f] = dcols[f]
            self.layouts[which] = L
            self.widths[which] = (1 + len(scols) + len(dcols)) * K
            self.mlps[which] = [(_dev(Wt, self.device, f32), _dev(b, self.device, f32), relu)
                                for Wt, b, relu in fold_mlp(weights[f"{which}_tower"])]

    def tower(self, which, ids):
        torch = self._torch
        ids_d = torch.as_tensor(np.asarray(ids, dtype=np.int64)).to(self.device)
        n = ids_d.numel()
        x = torch.empty((n, self.widths[which]), dtype=torch.float32, device=self.device)
        L = self.layouts[which]
        _lib.check(_lib.lib.b200_feat_forward(
            ctypes.byref(L), ctypes.byref(self.tables), _lib.ptr(ids_d), _lib.ptr(ids_d), n, 0, 0,
            _lib.ptr(x), x.stride(0), None, 0, None, None, None, 0.0, None, None, None, 0.0,
            None, None, 0, _lib.current_stream()))
        for Wt, b, relu in self.mlps[which]:
            x = linear(x, Wt, b, relu)
        if self.norm_embed:
            _lib.check(_lib.lib.b200_l2_normalize_rows(_lib.ptr(x), x.stride(0), n, x.shape[1],
                                                       _lib.current_stream()))
        return x

    def set_embeddings(self, chunk=1 << 20):
        """User / item vectors of every id + the mean OOV row; returns device tensors
        [n_users+1, d], [n_items+1, d]."""
        torch = self._torch
        outs = []
        for which, n in (("user", self.n_users), ("item", self.n_items)):
            rows = [self.tower(which, np.arange(i, min(n, i + chunk))) for i in range(0, n, chunk)]
            E = torch.cat(rows, dim=0)
            outs.append(torch.cat([E, E.mean(dim=0, keepdim=True)], dim=0))
        return outs[0], outs[1]


class YouTubeRetrieval:
    """libreco/algorithms/youtube_retrieval.py:169-260 (inference; SURVEY 8f-4 adjacent model) + the serving step of
    ``DynEmbedBase.set_embeddings`` / ``dyn_user_embedding`` (``bases/dyn_embed_base.py:166-269``): the user vector is
    ``dense_nn(concat(sqrtn-pooled behaviour sequence over seq_embeds_var, user sparse embeddings, user dense
    value x embedding))`` (optionally L2-normalised) with a pseudo bias 1 appended, the item vector is
    ``[item_embeds_var | item_bias_var]`` — the reference's own trick for folding the softmax bias into the dot
    product — so all-items retrieval IS the embed scorer (K4) on d = H + 1.  The sequence of a user = its last
    ``T`` consumed items (``_set_recent_seqs``; ``feat_models.recent_sequences``), pooled by ``b200_seq_pool``
    (sum / sqrt(count) = ``safe_embedding_lookup_sparse(combiner="sqrtn")``, empty history -> zero vector)."""

    def __init__(self, spec, weights, recent_seqs, recent_seq_lens, norm_embed=False, device=None):
        import torch

        self._torch = torch
        K = int(np.asarray(weights["seq_embeds"]).shape[1])
        self.base = FeatSpec(spec, K, device)
        self.device, self.K, self.norm_embed = self.base.device, K, bool(norm_embed)
        self.n_users, self.n_items = self.base.n_users, self.base.n_items
        f32 = torch.float32
        self.seq_embeds = _dev(weights["seq_embeds"], self.device, f32)              # [n_items, K]
        self.item_embeds = _dev(weights["item_embeds"], self.device, f32)            # [n_items, H]
        self.item_biases = _dev(np.asarray(weights["item_biases"]).reshape(-1), self.device, f32)
        self.t = {k: _dev(weights.get(k), self.device, f32) for k in ("sparse_embeds", "dense_embeds")}
        T = FeatTablesStruct()
        for k, v in self.t.items():
            setattr(T, k, v.data_ptr() if v is not None else None)
        self.tables = T
        L = FeatLayoutStruct.from_buffer_copy(self.base.layout)
        L.id_mask = 0                                      # no id-embedding field: the pooled sequence takes its place
        scols, dcols = self.base.user_sparse_cols, self.base.user_dense_cols
        L.n_sparse, L.n_dense = len(scols), len(dcols)
        for f in range(len(scols)):
            L.sparse_side[f], L.sparse_col[f] = 0, f
        for f in range(len(dcols)):
            L.dense_side[f], L.dense_col[f] = 0, f
            L.dense_embed_row[f] = dcols[f]
        self.layout, self.n_feat = L, len(scols) + len(dcols)
        self.seqs = _dev(recent_seqs, self.device, torch.int32)
        self.lens = _dev(recent_seq_lens, self.device, torch.int32)
        self.mlp = [(_dev(Wt, self.device, f32), _dev(b, self.device, f32), relu) for Wt, b, relu in fold_mlp(weights["mlp"])]

    def user_vectors(self, ids):
        """[n, H] user embeddings of the given (inner) user ids, before the pseudo bias."""
        torch = self._torch
        ids_d = torch.as_tensor(np.asarray(ids, dtype=np.int64)).to(self.device)
        n, K = int(ids_d.numel()), self.K
        x = torch.empty((n, (1 + self.n_feat) * K), dtype=torch.float32, device=self.device)
        pooled = x[:, :K]
        _lib.check(_lib.lib.b200_seq_pool(
            _lib.ptr(self.seq_embeds), self.seq_embeds.stride(0), K, self.n_items, _lib.ptr(self.seqs),
            self.seqs.stride(0), _lib.ptr(self.lens), self.seqs.shape[1], _lib.ptr(ids_d), n, 0, 0, _lib.ptr(pooled),
            x.stride(0), _lib.current_stream()))
        if self.n_feat:
            feat = x[:, K:]
            _lib.check(_lib.lib.b200_feat_forward(
                ctypes.byref(self.layout), ctypes.byref(self.tables), _lib.ptr(ids_d), _lib.ptr(ids_d), n, 0, 0,
                _lib.ptr(feat), x.stride(0), None, 0, None, None, None, 0.0, None, None, None, 0.0,
                None, None, 0, _lib.current_stream()))
        for Wt, b, relu in self.mlp:
            x = linear(x, Wt, b, relu)
        if self.norm_embed:
            _lib.check(_lib.lib.b200_l2_normalize_rows(_lib.ptr(x), x.stride(0), n, x.shape[1], _lib.current_stream()))
        return x

    def set_embeddings(self, chunk=1 << 20):
        """Device tensors ``U [n_users + 1, H + 1]`` (pseudo bias 1 in the last column, last row = mean OOV row) and
        ``I [n_items + 1, H + 1]`` (``[item_embeds | item_biases]`` + the mean row) for :class:`engine.EmbedScorer`."""
        torch = self._torch
        rows = [self.user_vectors(np.arange(i, min(self.n_users, i + chunk))) for i in range(0, self.n_users, chunk)]
        U = torch.cat(rows, dim=0)
        if self.norm_embed:          # dyn_embed_base.py:264-265: the item side is normalised too (before the bias column)
            I = self.item_embeds / self.item_embeds.norm(dim=1, keepdim=True)
        else:
            I = self.item_embeds
        U = torch.cat([U, torch.ones((U.shape[0], 1), dtype=torch.float32, device=self.device)], dim=1)
        I = torch.cat([I, self.item_biases[:, None]], dim=1)
        return (torch.cat([U, U.mean(dim=0, keepdim=True)], dim=0).contiguous(),
                torch.cat([I, I.mean(dim=0, keepdim=True)], dim=0).contiguous())
