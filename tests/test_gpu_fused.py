"""GPU parity of the fused tensor-core scorer (b200_recommend_embed) against the exact
materialised path and the oracle.  IDs must be identical to the exact path (both use the same
exact-score definition and tie rule); vs the oracle (numpy sgemm) they may differ only inside
near-ties (<= 1e-6 relative score gap)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(seed, n_users, N, d, mean_c, normalise=True, scale=1.0):
    rng = np.random.default_rng(seed)
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    if normalise:
        U /= np.linalg.norm(U, axis=1, keepdims=True)
        I /= np.linalg.norm(I, axis=1, keepdims=True)
    U *= scale
    consumed = {}
    for u in range(n_users):
        c = int(min(rng.poisson(mean_c), N // 2))
        if c:
            consumed[u] = rng.choice(N, size=c, replace=False).tolist()
    return U, I, consumed


@pytest.mark.parametrize("n_users,N,d,K,mean_c,B", [
    (500, 5000, 64, 100, 30, 300),     # several row tiles? no: 3 tiles of 128 rows
    (200, 70001, 64, 100, 50, 130),    # N not a multiple of 256, 2 row tiles
    (64, 3231, 16, 10, 20, 64),        # C1-like catalogue, d padded 16 -> 64
    (300, 20000, 128, 50, 10, 257),    # two K blocks
    (50, 1000, 7, 5, 3, 50),           # odd width
    (300, 30000, 160, 50, 10, 300),    # three K blocks: one ring stage per k-block (d_pad > 128)
    (200, 25000, 256, 100, 20, 200),   # four K blocks, the widest supported embedding
])
def test_fused_equals_exact_and_oracle(n_users, N, d, K, mean_c, B):
    import torch
    from librecommender_b200.engine import EmbedScorer
    from oracle import ranking as orc

    U, I, consumed = _mk(d + N, n_users, N, d, mean_c)
    sc = EmbedScorer(U, I, N, consumed, n_users=n_users)
    rng = np.random.default_rng(1)
    users = rng.integers(0, n_users + 1, size=B)        # includes the OOV row sometimes
    uid = torch.as_tensor(users).cuda()
    ids_f, sc_f, status = sc.recommend_fused(uid, K, True, True)
    ids_e, sc_e = sc.recommend_exact(uid, K, True, True)
    torch.cuda.synchronize()
    assert not bool((status != 0).any()), status[status != 0].tolist()   # reason codes: see finalize_kernel
    np.testing.assert_array_equal(ids_f.cpu().numpy(), ids_e.cpu().numpy())
    np.testing.assert_array_equal(sc_f.cpu().numpy(), sc_e.cpu().numpy())
    ref = orc.recommend_from_embedding("ranking", users.tolist(), K, U, I, N, consumed, True)
    full = orc.embed_scores(U, I, users, N)
    got = ids_f.cpu().numpy()
    assert orc.near_tie_mask(ref, got, full, 1e-6).all()
    assert (got == ref).mean() > 0.995
    for r, u in enumerate(users.tolist()):
        assert not set(got[r].tolist()) & set(consumed.get(u, []))


def test_fused_no_filter_and_unnormalised():
    import torch
    from librecommender_b200.engine import EmbedScorer

    U, I, consumed = _mk(5, 200, 30000, 64, 40, normalise=False, scale=3.0)
    sc = EmbedScorer(U, I, 30000, consumed, n_users=200)
    uid = torch.arange(0, 200).cuda()
    for flt in (False, True):
        ids_f, _, status = sc.recommend_fused(uid, 100, flt, False)
        ids_e = sc.recommend_exact(uid, 100, flt, False)
        assert int(status.sum()) == 0
        np.testing.assert_array_equal(ids_f.cpu().numpy(), ids_e.cpu().numpy())


def test_fused_heavy_users_and_dense_ties_fall_back():
    import torch
    from librecommender_b200.engine import EmbedScorer

    rng = np.random.default_rng(3)
    n_users, N, d, K = 40, 4000, 64, 20
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    I[: N // 2] = I[0]                      # half the catalogue ties exactly
    consumed = {0: rng.choice(N, size=1500, replace=False).tolist(),   # heavy user: capped k_row + verification
                1: list(range(10)), 2: rng.choice(N, size=N - 10, replace=False).tolist()}  # cannot filter
    sc = EmbedScorer(U, I, N, consumed, n_users=n_users)
    uid = torch.arange(0, n_users).cuda()
    ids = sc.recommend_device(uid, K, True, False, path="auto")
    ids_e = sc.recommend_exact(uid, K, True, False)
    np.testing.assert_array_equal(ids.cpu().numpy(), ids_e.cpu().numpy())
    ids_f, _, status = sc.recommend_fused(uid, K, True, False)
    ok = (status == 0).cpu().numpy()
    # rows the fused path accepted are exact (the heavy user is verified a posteriori, not refused)
    np.testing.assert_array_equal(ids_f.cpu().numpy()[ok], ids_e.cpu().numpy()[ok])
    assert ok[0] and ok[3:].all()


def test_fused_adversarial_near_ties():
    """Scores packed within a few bf16 ulps: the candidate margin must still deliver the exact
    fp32 top-K (ids equal to the exact path)."""
    import torch
    from librecommender_b200.engine import EmbedScorer

    rng = np.random.default_rng(11)
    n_users, N, d, K = 130, 50000, 64, 100
    base = rng.standard_normal(d).astype(np.float32)
    I = (base[None, :] + 1e-3 * rng.standard_normal((N + 1, d))).astype(np.float32)
    U = (base[None, :] + 1e-2 * rng.standard_normal((n_users + 1, d))).astype(np.float32)
    sc = EmbedScorer(U, I, N, {}, n_users=n_users)
    uid = torch.arange(0, n_users).cuda()
    ids = sc.recommend_device(uid, K, True, False)
    ids_e = sc.recommend_exact(uid, K, True, False)
    np.testing.assert_array_equal(ids.cpu().numpy(), ids_e.cpu().numpy())


def test_public_api_uses_fused_path_and_matches_oracle():
    from librecommender_b200 import recommend_from_embedding
    from oracle import ranking as orc

    U, I, consumed = _mk(9, 1000, 100000, 64, 50)
    model = types.SimpleNamespace(task="ranking", n_items=100000, n_users=1000, user_consumed=consumed)
    users = list(range(0, 1000, 2))
    got = recommend_from_embedding(model, users, 100, U, I, True, False)
    ref = orc.recommend_from_embedding("ranking", users, 100, U, I, 100000, consumed, True)
    full = orc.embed_scores(U, I, users, 100000)
    assert orc.near_tie_mask(ref, got, full, 1e-6).all()
    assert (got == ref).mean() > 0.995


def test_fused_heavy_users_with_top_ranked_history():
    """Consumed items that ARE the user's best-scoring items (the realistic case) and exceed the
    candidate budget: the capped rows must either be proven exact or be flagged — never wrong."""
    import torch
    from librecommender_b200.engine import EmbedScorer

    rng = np.random.default_rng(17)
    n_users, N, d, K = 64, 30000, 64, 50
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    full = U[:n_users] @ I[:N].T
    consumed = {}
    for u in range(n_users):
        c = [150, 400, 1200][u % 3]
        consumed[u] = np.argsort(-full[u])[:c].tolist()        # exactly the top-c items
    sc = EmbedScorer(U, I, N, consumed, n_users=n_users)
    uid = torch.arange(0, n_users).cuda()
    ids_f, _, status = sc.recommend_fused(uid, K, True, False)
    ids_e = sc.recommend_exact(uid, K, True, False)
    ok = (status == 0).cpu().numpy()
    np.testing.assert_array_equal(ids_f.cpu().numpy()[ok], ids_e.cpu().numpy()[ok])
    assert ok[0::3].all()                    # K + 150 fits the budget
    assert not ok[2::3].any()                # 1200 top-ranked consumed items cannot be proven -> flagged
    got = sc.recommend_device(uid, K, True, False).cpu().numpy()
    np.testing.assert_array_equal(got, ids_e.cpu().numpy())
