"""Negative sampling behind the reference's call seams.

Two modes:

* **parity** — ``negatives_from_random`` / ``negatives_from_popular`` / ``negatives_from_unconsumed``
  with the reference's signatures (``libreco/sampling/negatives.py:17,34,55``).  They consume the
  SAME host random streams as the reference (numpy ``Generator`` seeded ``seed % 3407 * 11``,
  ``libreco/batch/collators.py:180-187``; Python ``random`` for the unconsumed sampler), so the
  sampled indices are bit-identical under a fixed seed.  Host work by construction: those
  generators are sequential.
* **fast** — :class:`DeviceNegativeSampler`: the Philox-based CUDA kernel
  (``b200_sample_negatives``) with the same rejection rules; output is a pure function of
  (seed, step, index).  Statistically equivalent, not stream-equal.
"""
from __future__ import annotations

import math
import random as _py_random

import numpy as np

from . import _lib
from .consumed import as_csr

MODES = {"random": 0, "unconsumed": 1, "popular": 2}


# ----------------------------------------------------------------------------------------------
# parity mode (host streams identical to the reference)
# ----------------------------------------------------------------------------------------------
def _clashes(negatives, positives, extra):
    hit = negatives == positives
    if extra is not None and len(extra) > 0:
        hit = hit | (negatives == extra)
    return np.flatnonzero(hit)


def negatives_from_random(np_rng, n_items, items_pos, num_neg, items=None, tolerance=10):
    """negatives.py:17-31: one vectorised ``choice`` then up to ``tolerance`` re-draw rounds for
    entries equal to their own positive (or to ``items`` when given)."""
    positives = np.repeat(items_pos, num_neg) if num_neg > 1 else items_pos
    extra = np.repeat(items, num_neg) if (num_neg > 1 and items is not None) else items
    total = len(positives)
    negatives = np_rng.choice(n_items, size=total, replace=not (total < n_items))
    for _ in range(tolerance):
        bad = _clashes(negatives, positives, extra)
        if len(bad) == 0:
            break
        negatives[bad] = np_rng.choice(n_items, size=len(bad), replace=True)
    return negatives


def negatives_from_popular(np_rng, n_items, items_pos, num_neg, items=None, probs=None):
    """negatives.py:34-43: weighted ``choice`` and ONE re-draw round."""
    positives = np.repeat(items_pos, num_neg) if num_neg > 1 else items_pos
    extra = np.repeat(items, num_neg) if (num_neg > 1 and items is not None) else items
    negatives = np_rng.choice(n_items, size=len(positives), replace=True, p=probs)
    bad = _clashes(negatives, positives, extra)
    if len(bad):
        negatives[bad] = np_rng.choice(n_items, size=len(bad), replace=True, p=probs)
    return negatives


def negatives_from_unconsumed(user_consumed_set, users, items, n_items, num_neg, tolerance=10):
    """negatives.py:55-82: per (user, positive) rejection sampling on Python's ``random`` stream."""
    rnd, fl = _py_random.random, math.floor
    out = []
    for u, pos in zip(users, items):
        mine = []
        for _ in range(num_neg):
            cand = fl(n_items * rnd())
            accepted = False
            for _ in range(tolerance):
                if cand != pos and cand not in mine and cand not in user_consumed_set[u]:
                    accepted = True
                    break
                cand = fl(n_items * rnd())
            if not accepted:
                for _ in range(tolerance):
                    if cand != pos and cand not in mine:
                        break
                    cand = fl(n_items * rnd())
            mine.append(cand)
        out.extend(mine)
    return np.array(out)


def neg_probs_from_frequency(item_consumed, n_items, temperature):
    """negatives.py:85-93."""
    freqs = np.array([len(set(item_consumed[i])) for i in range(n_items)], dtype=np.float64)
    if temperature != 1.0:
        freqs = np.power(freqs, temperature)
    return freqs / np.sum(freqs)


def collator_seed(seed: int) -> int:
    """collators.py:180-187 — the per-collator / per-worker seed derivation."""
    return seed % 3407 * 11


# ----------------------------------------------------------------------------------------------
# fast mode (device)
# ----------------------------------------------------------------------------------------------
def rank_stream_seed(seed: int, rank: int) -> int:
    """Philox key of rank `rank` (stream id = (seed, rank, step)): rank 0 keeps the
    single-process stream, every other rank gets a key a splitmix64 step away, so data-parallel ranks
    never draw the same negatives for the same step."""
    if rank == 0:
        return int(seed)
    z = (int(seed) + (rank * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return int((z ^ (z >> 31)) & 0x7FFFFFFFFFFFFFFF)


class DeviceNegativeSampler:
    def __init__(self, n_items, user_consumed=None, n_users=None, neg_probs=None, seed=42,
                 tolerance=10, device=None, rank=None):
        import torch

        self.device = device if device is not None else _lib.require_cuda()
        self.n_items = int(n_items)
        if rank is None:   # one process per GPU: the data-parallel rank selects the stream
            rank = torch.distributed.get_rank() if (torch.distributed.is_available()
                                                    and torch.distributed.is_initialized()) else 0
        self.rank = int(rank)
        self.seed = rank_stream_seed(int(collator_seed(seed)), self.rank)
        self.tolerance = int(tolerance)
        self.step = 0
        self.indptr = self.idx_sorted = None
        self.n_users = 0
        if user_consumed is not None:
            csr = as_csr(user_consumed, n_users)   # n_users None: derived from the dict's largest key
            indptr, idx = csr.device(self.device)
            self.n_users = csr.n_users
            # per-user sorted copy for the binary-search rejection test
            deg = indptr[1:] - indptr[:-1]
            owner = torch.repeat_interleave(torch.arange(self.n_users, device=self.device), deg)
            nnz = int(indptr[-1])
            key = (owner << 32) | idx[:nnz].to(torch.int64)
            self.idx_sorted = (torch.sort(key).values & 0xFFFFFFFF).to(torch.int32).contiguous()
            self.indptr = indptr.contiguous()
        self.cdf = None
        if neg_probs is not None:
            p = torch.as_tensor(np.asarray(neg_probs, dtype=np.float64), device=self.device)
            self.cdf = torch.cumsum(p, 0).to(torch.float32).contiguous()

    def sample(self, users, items_pos, num_neg, sampler="random", step=None):
        """users / items_pos: device int64 tensors; returns device int64[len * num_neg]."""
        import torch

        mode = MODES[sampler]
        if step is None:
            step = self.step
            self.step += 1
        items_pos = items_pos.contiguous()
        users = users.contiguous() if users is not None else None
        out = torch.empty(items_pos.numel() * num_neg, dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib.b200_sample_negatives(
            _lib.ptr(users), _lib.ptr(items_pos), items_pos.numel(), int(num_neg), self.n_items, mode,
            self.tolerance, self.seed, int(step), _lib.ptr(self.indptr), _lib.ptr(self.idx_sorted),
            self.n_users, _lib.ptr(self.cdf), _lib.ptr(out), _lib.current_stream()))
        return out
