"""``user_consumed`` in the layout the device filter reads: CSR.

The reference keeps ``user_consumed: dict[int, list[int]]`` built by
``libreco/data/consumed.py:7-19`` → ``recfarm.build_consumed_unique``
(``rust/src/utils.rs:8-35``: per-user lists in arrival order with CONSECUTIVE
repeats removed — lists may still contain duplicates, and ``len()`` including
those duplicates enters the "can we filter" test of ``ranking.py:38``).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class ConsumedCSR:
    """indptr int64[n_users+1], idx int32[nnz]; order inside a row is preserved."""

    def __init__(self, indptr: np.ndarray, idx: np.ndarray):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        assert self.indptr.ndim == 1 and len(self.indptr) >= 1
        assert int(self.indptr[-1]) == len(self.idx)
        self._dev = {}

    @property
    def n_users(self) -> int:
        if self.indptr is None:
            return self._n_users
        return len(self.indptr) - 1

    @property
    def nnz(self) -> int:
        if self.idx is None:
            return int(next(iter(self._dev.values()))[1].numel())
        return len(self.idx)

    # ---- constructors ------------------------------------------------------------------
    @classmethod
    def from_interactions(cls, user_indices, item_indices, n_users: int) -> "ConsumedCSR":
        """Native (C++) equivalent of ``build_consumed_unique`` for the user side."""
        u = np.ascontiguousarray(user_indices, dtype=np.int64)
        i = np.ascontiguousarray(item_indices, dtype=np.int64)
        if len(u) != len(i):
            raise ValueError("user_indices and item_indices differ in length")
        indptr = np.empty(n_users + 1, dtype=np.int64)
        idx = np.empty(max(len(u), 1), dtype=np.int32)
        nnz = ctypes.c_int64(0)
        _lib.check(_lib.lib.b200_build_consumed_csr_host(
            _lib.ptr(u), _lib.ptr(i), len(u), n_users, _lib.ptr(indptr), _lib.ptr(idx),
            ctypes.byref(nnz)))
        return cls(indptr, idx[: nnz.value].copy())

    @classmethod
    def from_dict(cls, user_consumed, n_users: int | None = None) -> "ConsumedCSR":
        """Convert the reference's ``dict[int, list[int]]`` (lists kept verbatim)."""
        if n_users is None:
            n_users = (max(user_consumed) + 1) if len(user_consumed) else 0
        counts = np.zeros(n_users, dtype=np.int64)
        for u, items in user_consumed.items():
            if 0 <= u < n_users:
                counts[u] = len(items)
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        idx = np.empty(int(indptr[-1]), dtype=np.int32)
        for u, items in user_consumed.items():
            if 0 <= u < n_users and len(items):
                idx[indptr[u]: indptr[u + 1]] = items
        return cls(indptr, idx)

    def to_dict(self) -> dict:
        out = {}
        for u in range(self.n_users):
            b, e = int(self.indptr[u]), int(self.indptr[u + 1])
            if e > b:
                out[u] = self.idx[b:e].tolist()
        return out

    def row(self, u: int) -> np.ndarray:
        return self.idx[int(self.indptr[u]): int(self.indptr[u + 1])]

    # ---- device residency --------------------------------------------------------------
    def device(self, device):
        import torch

        dv = torch.device(device)
        if dv.type == "cuda" and dv.index is None:       # "cuda" and "cuda:<current>" are the same residency
            dv = torch.device("cuda", torch.cuda.current_device())
        key = str(dv)
        if key not in self._dev:
            if self.indptr is None:
                raise ValueError(f"device-only CSR lives on {list(self._dev)}, asked for {key}")
            idx = self.idx if len(self.idx) else np.zeros(1, dtype=np.int32)
            self._dev[key] = (
                torch.from_numpy(self.indptr).to(device),
                torch.from_numpy(idx).to(device),
            )
        return self._dev[key]

    @classmethod
    def from_device_tensors(cls, indptr, idx):
        """Wrap CSR tensors that already live on the GPU (large synthetic catalogues)."""
        self = cls.__new__(cls)
        self.indptr = None
        self.idx = None
        self._dev = {str(indptr.device): (indptr.contiguous(), idx.contiguous())}
        self._n_users = indptr.numel() - 1
        return self


_dict_cache: dict = {}


def as_csr(user_consumed, n_users: int) -> ConsumedCSR:
    """Accept a :class:`ConsumedCSR` or the reference's dict (converted once per dict
    object and cached; the reference never mutates ``user_consumed`` after the data build)."""
    if isinstance(user_consumed, ConsumedCSR):
        return user_consumed
    key = (id(user_consumed), len(user_consumed), n_users)
    hit = _dict_cache.get(key)
    if hit is not None and hit[0] is user_consumed:
        return hit[1]
    csr = ConsumedCSR.from_dict(user_consumed, n_users)
    if len(_dict_cache) > 8:
        _dict_cache.clear()
    _dict_cache[key] = (user_consumed, csr)
    return csr
