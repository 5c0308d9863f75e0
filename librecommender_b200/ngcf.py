"""NGCF embedding propagation on the GPU (an adjacent model that reuses the LightGCN
kernels) — inference forward of ``NGCFModel`` in
``libreco/algorithms/torch_modules/ngcf_module.py:8-130``:

* ``build_ngcf_laplacian_csr``: ``L = D^-1 (A + I)`` (row-normalised bipartite adjacency with self
  loops, :61-85), vectorised on the device like the LightGCN builder;
* ``NGCFPropagator.forward()``: per layer ``side = L E`` (``b200_spmm_csr``),
  ``self = side W_self + b_self``, ``pair = (side ⊙ E) W_pair + b_pair`` (``b200_linear_*``),
  ``E' = normalize(leaky_relu(self + pair, 0.2))`` (``b200_ngcf_combine``); the output is the
  concatenation of all layers (:96-124).  Dropout is a training-time feature and is not applied.

The resulting ``(user_embeds, item_embeds)`` feed ``recommend_from_embedding`` like every EmbedBase
model (``libreco/algorithms/ngcf.py`` -> ``set_embeddings``)."""
from __future__ import annotations

import numpy as np

from . import _lib
from .consumed import as_csr
from .feat_models import _dev, linear
from .lightgcn import SpmmGraph


def build_ngcf_laplacian_csr(user_consumed, n_users, n_items, device=None):
    """(indptr int64[n+1], col int32[nnz], val float32[nnz]) of D^-1 (A + I), sorted by (row, col)."""
    import torch

    device = device if device is not None else _lib.require_cuda()
    csr = as_csr(user_consumed, n_users)
    indptr_d, idx_d = csr.device(device)
    indptr_d = indptr_d[: n_users + 1]
    counts = indptr_d[1:] - indptr_d[:-1]
    users = torch.repeat_interleave(torch.arange(n_users, device=device), counts)
    items = idx_d[: int(indptr_d[-1])].to(torch.int64)
    n = n_users + n_items
    shift = max(1, (n - 1).bit_length())
    und = torch.unique((users << shift) | (items + n_users))
    r, c = und >> shift, und & ((1 << shift) - 1)
    diag = torch.arange(n, device=device)
    key = torch.cat([und, (c << shift) | r, (diag << shift) | diag])      # both directions + self loops
    key = torch.sort(key).values
    rows = key >> shift
    cols = (key & ((1 << shift) - 1)).to(torch.int32)
    deg = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(deg, 0)
    # the reference computes np.power(row_sum, -1) on the float matrix row sums (float64 after
    # adding ssp.eye) and stores the products as float32 (ngcf_module.py:76-84)
    dinv = (1.0 / deg.to(torch.float64)).to(torch.float32)
    val = dinv[rows]
    return indptr, cols.contiguous(), val.contiguous()


class NGCFPropagator:
    """Inference forward of the reference module from its parameters.

    ``weights``: ``user_embed`` [n_users, d], ``item_embed`` [n_items, d] and per layer k
    ``W_self_k``, ``b_self_k``, ``W_pair_k``, ``b_pair_k`` (the reference's ParameterDict names,
    ngcf_module.py:33-59; kernels are [d_in, d_out])."""

    def __init__(self, n_users, n_items, user_consumed, weights, device=None):
        import torch

        self._torch = torch
        self.device = device if device is not None else _lib.require_cuda()
        self.n_users, self.n_items = n_users, n_items
        self.graph = SpmmGraph(*build_ngcf_laplacian_csr(user_consumed, n_users, n_items, self.device))
        f32 = torch.float32
        self.E0 = torch.cat([_dev(weights["user_embed"], self.device, f32), _dev(weights["item_embed"], self.device, f32)])
        self.layers = []
        k = 0
        while f"W_self_{k}" in weights:
            g = lambda name: _dev(np.asarray(weights[f"{name}_{k}"]), self.device, f32)   # noqa: E731
            self.layers.append((g("W_self").t().contiguous(), g("b_self").reshape(-1).contiguous(),
                                g("W_pair").t().contiguous(), g("b_pair").reshape(-1).contiguous()))
            k += 1

    def forward(self):
        """(user_embeds [n_users, d + sum(layers)], item_embeds [n_items, ...]) device tensors."""
        torch = self._torch
        lib, st = _lib.lib, _lib.current_stream()
        outs, E = [self.E0], self.E0
        n = E.shape[0]
        for Wst, bs, Wpt, bp in self.layers:
            side = torch.empty_like(E)
            self.graph.spmm(E, out=side)
            prod = torch.empty_like(E)
            _lib.check(lib.b200_mul_elementwise(_lib.ptr(side), _lib.ptr(E), side.numel(), _lib.ptr(prod), st))
            s_part = linear(side, Wst, bs, False)
            p_part = linear(prod, Wpt, bp, False)
            nxt = torch.empty((n, Wst.shape[0]), dtype=torch.float32, device=self.device)
            _lib.check(lib.b200_ngcf_combine(_lib.ptr(s_part), s_part.stride(0), _lib.ptr(p_part), p_part.stride(0), n,
                                             Wst.shape[0], 0.2, _lib.ptr(nxt), nxt.stride(0), st))
            outs.append(nxt)
            E = nxt
        full = torch.cat(outs, dim=1)
        return full[: self.n_users], full[self.n_users:]
