"""LightGCN propagation on the GPU — drop-in for ``LightGCNModel`` of
``libreco/algorithms/torch_modules/lightgcn_module.py:7-96``.

* ``build_laplacian_csr``: vectorised replacement of ``_build_laplacian_matrix`` (:36-61, a Python
  loop into a DOK matrix in the reference): binary bipartite adjacency from ``user_consumed``
  (duplicates collapse), ``L = D^-1/2 A D^-1/2`` in float32, CSR sorted by (row, col) — the same
  order ``scipy``'s ``tocoo()`` gives the reference's COO tensor.
* ``propagate``: ``E^{l+1} = L E^l`` with the CUDA CSR SpMM (``b200_spmm_csr``), layer mean fused.
* ``LightGCNModel``: ``nn.Module`` with the reference's attribute names
  (``user_init_embeds`` / ``item_init_embeds``) whose ``forward(use_dropout)`` returns
  ``(user_embeds, item_embeds)``; differentiable (backward = the same SpMM on L^T, L symmetric;
  under edge dropout the transposed values are reached through a precomputed permutation).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .consumed import as_csr


def build_laplacian_csr(user_consumed, n_users, n_items, device=None):
    """Returns (indptr int64[n+1], col int32[nnz], val float32[nnz]) on ``device``."""
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
    und = torch.unique((users << shift) | (items + n_users))       # one edge per (u, i) pair
    r = und >> shift
    c = und & ((1 << shift) - 1)
    key = torch.cat([und, (c << shift) | r])                       # both directions
    key = torch.sort(key).values                                   # (row, col) order
    rows = key >> shift
    cols = (key & ((1 << shift) - 1)).to(torch.int32)
    deg = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(deg, 0)
    # D^-1/2 with the reference's own arithmetic (numpy float32 power, inf -> 0; :50-52)
    with np.errstate(divide="ignore"):
        dinv_h = np.power(deg.cpu().numpy().astype(np.float32), np.float32(-0.5)).astype(np.float32)
    dinv_h[np.isinf(dinv_h)] = 0.0
    dinv = torch.from_numpy(dinv_h).to(device)
    val = (dinv[rows] * 1.0) * dinv[cols.to(torch.int64)]
    return indptr, cols.contiguous(), val.contiguous()


class SpmmGraph:
    """CSR matrix + the long-row chunk plan ``b200_spmm_csr`` needs."""

    def __init__(self, indptr, col, val):
        import torch

        self.indptr, self.col, self.val = indptr, col, val
        self.device = indptr.device
        self.n = indptr.numel() - 1
        thr = _lib.lib.b200_spmm_long_row_threshold()
        chunk = _lib.lib.b200_spmm_chunk()
        deg = indptr[1:] - indptr[:-1]
        long_rows = torch.nonzero(deg > thr).flatten()
        self.n_long = int(long_rows.numel())
        if self.n_long:
            nch = (deg[long_rows] + chunk - 1) // chunk
            self.long_rows = long_rows.to(torch.int32)
            self.long_chunk_ptr = torch.zeros(self.n_long + 1, dtype=torch.int64, device=self.device)
            self.long_chunk_ptr[1:] = torch.cumsum(nch, 0)
            self.n_chunks = int(self.long_chunk_ptr[-1])
            owner = torch.repeat_interleave(torch.arange(self.n_long, device=self.device), nch)
            self.chunk_row = long_rows[owner].to(torch.int32)
            self.chunk_k = (torch.arange(self.n_chunks, device=self.device)
                            - self.long_chunk_ptr[owner]).to(torch.int32)
        else:
            self.long_rows = self.long_chunk_ptr = self.chunk_row = self.chunk_k = None
            self.n_chunks = 0
        self._partials = None
        self._tperm = None

    @property
    def nnz(self):
        return int(self.col.numel())

    def partials(self, d):
        import torch

        if self.n_chunks == 0:
            return None
        if self._partials is None or self._partials.numel() < self.n_chunks * d:
            self._partials = torch.empty(self.n_chunks * d, dtype=torch.float32, device=self.device)
        return self._partials

    def transpose_perm(self):
        """perm such that val[perm] are the values of L^T in CSR order (structure is symmetric)."""
        import torch

        if self._tperm is None:
            deg = self.indptr[1:] - self.indptr[:-1]
            rows = torch.repeat_interleave(torch.arange(self.n, device=self.device), deg)
            key_t = self.col.to(torch.int64) * self.n + rows      # entry (c, r) of the transpose
            self._tperm = torch.argsort(key_t)                    # position k of L^T <- entry perm[k]
        return self._tperm

    def spmm(self, E, out=None, acc=None, acc_init=False, final_div=0.0, val=None):
        import torch

        E = E.contiguous()
        d = int(E.shape[1])
        val = self.val if val is None else val
        if out is None and acc is None:
            out = torch.empty((self.n, d), dtype=torch.float32, device=self.device)
        part = self.partials(d)
        _lib.check(_lib.lib.b200_spmm_csr(
            _lib.ptr(self.indptr), _lib.ptr(self.col), _lib.ptr(val), self.n,
            _lib.ptr(E), E.stride(0), d,
            _lib.ptr(out), out.stride(0) if out is not None else 0,
            _lib.ptr(acc), acc.stride(0) if acc is not None else 0,
            1 if acc_init else 0, float(final_div),
            _lib.ptr(self.long_rows), _lib.ptr(self.long_chunk_ptr), self.n_long,
            _lib.ptr(self.chunk_row), _lib.ptr(self.chunk_k), self.n_chunks,
            _lib.ptr(part), _lib.current_stream()))
        return out


def propagate(graph: SpmmGraph, E0, n_layers: int, val=None):
    """mean over l = 0..n_layers of L^l E0 (lightgcn_module.py:74-88)."""
    import torch

    E0 = E0.contiguous().float()
    if n_layers == 0:
        return E0.clone()
    acc = torch.empty_like(E0)
    bufs = [torch.empty_like(E0), torch.empty_like(E0)]
    cur = E0
    for layer in range(n_layers):
        out = bufs[layer & 1]
        last = layer == n_layers - 1
        graph.spmm(cur, out=None if last else out, acc=acc, acc_init=(layer == 0),
                   final_div=float(n_layers + 1) if last else 0.0, val=val)
        cur = out
    return acc


def _make_function():
    import torch

    class _Propagate(torch.autograd.Function):
        @staticmethod
        def forward(ctx, E0, graph, n_layers, val):
            ctx.graph, ctx.n_layers, ctx.val = graph, n_layers, val
            return propagate(graph, E0.detach(), n_layers, val)

        @staticmethod
        def backward(ctx, grad):
            g = ctx.graph
            val_t = None
            if ctx.val is not None:                     # dropout made L non-symmetric
                val_t = ctx.val[g.transpose_perm()].contiguous()
            return propagate(g, grad.contiguous(), ctx.n_layers, val_t), None, None, None

    return _Propagate


_PropagateFn = None


def propagate_autograd(graph, E0, n_layers, val=None):
    global _PropagateFn
    if _PropagateFn is None:
        _PropagateFn = _make_function()
    return _PropagateFn.apply(E0, graph, n_layers, val)


def make_lightgcn_model_class():
    import torch
    from torch import nn

    class LightGCNModel(nn.Module):
        """Same constructor and ``forward(use_dropout)`` contract as the reference module."""

        def __init__(self, n_users, n_items, embed_size, n_layers, dropout_rate, user_consumed, device):
            super().__init__()
            self.n_users, self.n_items = n_users, n_items
            self.embed_size, self.n_layers = embed_size, n_layers
            self.dropout_rate = dropout_rate
            self.user_consumed = user_consumed
            self.device = torch.device(device)
            self.user_init_embeds = nn.Embedding(n_users, embed_size)
            self.item_init_embeds = nn.Embedding(n_items, embed_size)
            nn.init.normal_(self.user_init_embeds.weight, 0.0, 0.1)     # lightgcn_module.py:32-33
            nn.init.normal_(self.item_init_embeds.weight, 0.0, 0.1)
            self.to(self.device)
            self.graph = SpmmGraph(*build_laplacian_csr(user_consumed, n_users, n_items, self.device))

        def forward(self, use_dropout):
            return self.embedding_propagation(use_dropout)

        def embedding_propagation(self, use_dropout):
            val = None
            if use_dropout and self.dropout_rate > 0:
                # lightgcn_module.py:90-96: CPU torch.rand stream, floor(rand + keep), rescale
                keep = 1 - self.dropout_rate
                mask = torch.floor(torch.rand(self.graph.nnz) + keep).bool().to(self.device)
                val = torch.where(mask, self.graph.val / keep, torch.zeros_like(self.graph.val))
            E0 = torch.cat([self.user_init_embeds.weight, self.item_init_embeds.weight], dim=0)
            out = propagate_autograd(self.graph, E0, self.n_layers, val)
            return torch.split(out, [self.n_users, self.n_items])

    return LightGCNModel
