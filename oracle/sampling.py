"""Oracle for negative sampling.  TEST INFRASTRUCTURE ONLY.

Two things live here:

1. ``device_sampler_reference`` — a numpy restatement of the DEVICE sampler
   (``librecommender_b200/csrc/sampler.cu``: Philox4x32-10 counter RNG + the rejection rules of
   ``libreco/sampling/negatives.py:17-82``).  The kernel must match it bit-for-bit.
2. ``check_reference_invariants`` — the properties the reference's own tests pin for its samplers
   (``tests/test_collators.py:399-414``: a sampled negative is never the row's positive and, for
   the ``unconsumed`` sampler, never a consumed item when avoidable), used on both the host parity
   mode and the device mode.

The reference's samplers themselves (numpy ``Generator.choice`` on PCG64, Python ``random``) are
third-party streams: the host parity mode (``librecommender_b200/sampling.py``) is validated
against the unmodified reference functions directly (``tests/test_sampling_cpu.py``, run when
``/root/reference`` is present) and against golden vectors generated from them.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint32 arrays (or scalars)."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & MASK32).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _draw(mode, cdf, n_items, seed, step, index, attempt):
    """One candidate for flat sample `index` at `attempt` (scalars)."""
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF
    r0, r1, _, _ = philox4x32_10(index & 0xFFFFFFFF, (index >> 32) & 0xFFFFFFFF, attempt,
                                 step & 0xFFFFFFFF, k0, k1)
    r0, r1 = int(r0), int(r1)
    if mode != 2:
        return (((r0 << 32) | r1) * int(n_items)) >> 64
    u = np.float32(r0 >> 8) * np.float32(1.0 / 16777216.0)
    lo, hi = 0, int(n_items) - 1
    while lo < hi:
        mid = (lo + hi) >> 1
        if cdf[mid] > u:
            hi = mid
        else:
            lo = mid + 1
    return lo


def device_sampler_reference(users, items_pos, num_neg, n_items, mode, tolerance, seed, step,
                             consumed_sorted=None, cdf=None):
    """Bit-exact restatement of sample_negatives_kernel (scalar loops: small inputs only)."""
    out = np.empty(len(items_pos) * num_neg, dtype=np.int64)
    for j, pos in enumerate(np.asarray(items_pos).tolist()):
        cons = ()
        if mode == 1:
            cons = consumed_sorted.get(int(users[j]), ())
        negs = []
        for t in range(num_neg):
            index, attempt = j * num_neg + t, 0
            n = _draw(mode, cdf, n_items, seed, step, index, attempt)
            attempt += 1
            if mode == 0:
                a = 0
                while a < tolerance and n == pos:
                    n = _draw(mode, cdf, n_items, seed, step, index, attempt)
                    attempt += 1
                    a += 1
            elif mode == 2:
                if n == pos:
                    n = _draw(mode, cdf, n_items, seed, step, index, attempt)
            else:
                ok = False
                for _ in range(tolerance):
                    if n != pos and n not in negs and n not in cons:
                        ok = True
                        break
                    n = _draw(mode, cdf, n_items, seed, step, index, attempt)
                    attempt += 1
                if not ok:
                    for _ in range(tolerance):
                        if n != pos and n not in negs:
                            break
                        n = _draw(mode, cdf, n_items, seed, step, index, attempt)
                        attempt += 1
            negs.append(n)
        out[j * num_neg:(j + 1) * num_neg] = negs
    return out


def check_reference_invariants(negatives, users, items_pos, num_neg, n_items, user_consumed=None,
                               allow_consumed_fraction=0.0):
    """tests/test_collators.py:399-414 style invariants."""
    negatives = np.asarray(negatives)
    assert negatives.shape == (len(items_pos) * num_neg,)
    assert negatives.min() >= 0 and negatives.max() < n_items
    rep_pos = np.repeat(np.asarray(items_pos), num_neg)
    assert (negatives != rep_pos).mean() > 0.999
    if user_consumed is not None:
        bad = 0
        for j, u in enumerate(np.asarray(users).tolist()):
            cons = set(user_consumed.get(u, ()))
            bad += sum(int(n in cons) for n in negatives[j * num_neg:(j + 1) * num_neg].tolist())
        assert bad <= allow_consumed_fraction * len(negatives), f"{bad} consumed negatives"


def neg_probs_from_frequency(item_consumed, n_items, temperature):
    """negatives.py:85-93 — p_i ~ len(set(item_consumed[i]))**temperature."""
    f = np.array([len(set(item_consumed.get(i, ()))) for i in range(n_items)], dtype=np.float64)
    if temperature != 1.0:
        f = np.power(f, temperature)
    return f / f.sum()
