"""A numpy MODEL of the fused scorer's selection algorithm (csrc/score_topk_tc.cu header: coarse
fp16 scores of power-of-two scaled rows, error bound eps, speculative threshold from sampled block maxima, collection of every
item with coarse >= tau, finalize with the a-posteriori speculation check, cut at c_k - 2 eps,
consumed filter, exact fp32 re-score).  It checks the EXACTNESS ARGUMENT itself, independent of any
kernel: whenever the model accepts a row, its top-K equals the exact top-K of the oracle; rows it
cannot prove are flagged — including adversarial catalogues (near-ties inside 2 eps, a threshold
guessed too high, heavy users with capped k_row)."""
import numpy as np
import pytest

ERR_COEF = 0.00097705
KROW_MAX = 288


def _pow2_scale(nrm):
    """power of two s with s * nrm in [64, 128) — pow2_scale_for in the kernel source"""
    if not (nrm > 0):
        return 1.0
    m, x = np.frexp(np.float32(nrm))
    return float(np.ldexp(1.0, 7 - int(x)))


def _f16(x):
    """round-to-nearest-even to fp16 (subnormals kept), returned as float64"""
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float64)


def _exact_scores(u, I):
    acc = np.zeros(len(I), dtype=np.float32)
    for k in range(len(u)):                                   # sequential fp32 fma chain in k
        acc = (acc.astype(np.float64) + np.float64(u[k]) * I[:, k].astype(np.float64)).astype(np.float32)
    return acc


def model_row(u, I, consumed, K, stride=2, block=16, guess_scale=1.0, stats=None):
    # (the kernel samples every 16th 256-item tile in 128-item blocks; the model uses finer blocks so
    # that the speculation engages on catalogues small enough for a numpy test)
    """Returns (ids or None, status).  status 0 = accepted, 2 = too few collected, 3 = failed
    speculation, 5 = capped row not provable."""
    N, d = I.shape
    nu = float(np.linalg.norm(u.astype(np.float64))) * 1.0001
    ni = float(np.linalg.norm(I.astype(np.float64), axis=1).max()) * 1.0001
    su, si = _pow2_scale(nu), _pow2_scale(ni)
    # everything below lives in the row's SCALED units (coarse ~ su * si * exact)
    coarse = (_f16(u * np.float32(su))[None, :] * _f16(I * np.float32(si))).sum(1).astype(np.float32)
    d_pad = -(-d // 64) * 64
    eps = (ERR_COEF + d_pad * 2.4e-7) * (nu * su) * (ni * si) + np.sqrt(d_pad) * 6.2e-5 * (nu * su + ni * si)
    apply = len(consumed) > 0 and K + len(consumed) <= N
    k_full = K + (len(consumed) if apply else 0)
    k_row = min(k_full, KROW_MAX)
    capped = k_row < k_full
    # speculative threshold: pre_k-th largest maximum over the sampled 128-item blocks
    n_blocks = N // block
    sampled = [b for b in range(n_blocks) if (b // (256 // block)) % stride == 0]
    tau = -np.inf
    if len(sampled) * 1 >= 16:
        f = len(sampled) / max(n_blocks, 1)
        pre_k = 12 + int(np.ceil(2.0 * f * k_row))        # g_pre_margin, g_pre_coef of score_topk_tc.cu
        bm = np.sort([coarse[b * block:(b + 1) * block].max() for b in sampled])[::-1]
        if len(bm) >= pre_k:
            tau = bm[pre_k - 1] * guess_scale if bm[pre_k - 1] > 0 else bm[pre_k - 1] / guess_scale
    if stats is not None:
        stats["speculated"] = stats.get("speculated", 0) + int(np.isfinite(tau))
    collected = np.nonzero(coarse >= tau)[0]
    if len(collected) < k_row:
        return None, 2
    c_k = np.sort(coarse[collected])[::-1][k_row - 1]
    thr = c_k - 2 * eps
    if tau > thr:
        return None, 3                                        # the guess may have hidden a top item
    cand = collected[coarse[collected] >= thr]
    if apply:
        cand = cand[~np.isin(cand, consumed)]
    ex = _exact_scores(u.astype(np.float32), I[cand].astype(np.float32))
    order = np.lexsort((cand, -ex.astype(np.float64)))
    cand, ex = cand[order], ex[order]
    if len(cand) < K:
        return None, 5
    if capped and ex[K - 1] * (su * si) < thr + eps:          # an uncollected item could still beat it
        return None, 5
    return cand[:K], 0


def _catalogue(rng, N, d, mode):
    I = rng.standard_normal((N, d)).astype(np.float32)
    I /= np.linalg.norm(I, axis=1, keepdims=True)
    if mode == "near_ties":                                   # many items within 2 eps of each other
        base = I[0].copy()
        I[: N // 4] = base + 1e-3 * rng.standard_normal((N // 4, d)).astype(np.float32)
    return I


@pytest.mark.parametrize("mode", ["random", "near_ties"])
@pytest.mark.parametrize("K", [10, 100])
def test_accepted_rows_are_exact(mode, K):
    rng = np.random.default_rng(K + len(mode))
    N, d = 6000, 32
    I = _catalogue(rng, N, d, mode)
    accepted, stats = 0, {}
    for trial in range(12):
        u = rng.standard_normal(d).astype(np.float32)
        u /= np.linalg.norm(u)
        if mode == "near_ties":
            u = (I[0] + 0.3 * u).astype(np.float32)
        consumed = rng.choice(N, size=int(rng.integers(0, 60)), replace=False)
        ids, status = model_row(u, I, consumed, K, stats=stats)
        exact = _exact_scores(u, I)
        masked = exact.astype(np.float64).copy()
        if len(consumed) and K + len(consumed) <= N:
            masked[consumed] = -np.inf
        ref = np.lexsort((np.arange(N), -masked))[:K]
        if status == 0:
            accepted += 1
            np.testing.assert_array_equal(ids, ref)           # bit-identical ids, (score desc, id asc)
        else:
            assert status in (2, 3, 5)
    assert accepted >= 6                                      # the fast path is the common case
    assert stats.get("speculated", 0) >= 3                     # ... and speculative thresholds were in play


def test_overshooting_guess_is_detected_never_trusted():
    """A threshold pushed above c_k - 2 eps must flag the row (status 3 / 2), never return ids."""
    rng = np.random.default_rng(0)
    N, d, K = 6000, 32, 50
    I = _catalogue(rng, N, d, "random")
    flagged = 0
    for _ in range(10):
        u = rng.standard_normal(d).astype(np.float32)
        ids, status = model_row(u, I, np.zeros(0, dtype=np.int64), K, guess_scale=1.6)
        if status == 0:                                       # still provable -> must still be exact
            ref = np.lexsort((np.arange(N), -_exact_scores(u, I).astype(np.float64)))[:K]
            np.testing.assert_array_equal(ids, ref)
        else:
            flagged += 1
    assert flagged >= 5


def test_heavy_user_capped_rows_are_exact_or_flagged():
    rng = np.random.default_rng(1)
    N, d, K = 8000, 32, 100
    I = _catalogue(rng, N, d, "random")
    for c_u in (250, 600):                                    # K + c_u > KROW_MAX -> capped
        u = rng.standard_normal(d).astype(np.float32)
        consumed = rng.choice(N, size=c_u, replace=False)
        # worst case: the user consumed exactly its best items
        best = np.argsort(-_exact_scores(u, I))[: c_u // 2]
        consumed[: len(best)] = best
        consumed = np.unique(consumed)
        ids, status = model_row(u, I, consumed, K)
        masked = _exact_scores(u, I).astype(np.float64)
        masked[consumed] = -np.inf
        ref = np.lexsort((np.arange(N), -masked))[:K]
        if status == 0:
            np.testing.assert_array_equal(ids, ref)
        else:
            assert status == 5
