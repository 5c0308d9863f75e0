"""Hostile inputs for the fused scorer, through the device API with the exact-path repair (`recommend_device`):
whatever the coarse pass does, ids AND scores must equal the exact materialised path bit for bit (same exact-score
definition, same tie rule), and every flagged row must have been repaired.  Shapes / values the main parity tests
do not reach: zero vectors, all-negative scores, exact duplicates across 8-column group and 256-column tile
borders, widths 1 / 17 / 200, one user, odd batch sizes around the cluster padding, duplicated consumed entries,
huge and tiny norms in one catalogue."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(U, I, N, consumed, users, K, n_users):
    import torch

    from librecommender_b200.engine import EmbedScorer

    sc = EmbedScorer(U, I, N, consumed, n_users=n_users)
    uid = torch.as_tensor(np.asarray(users, dtype=np.int64)).cuda()
    ids, scores = sc.recommend_device(uid, K, True, True)
    ids_e, sc_e = sc.recommend_exact(uid, K, True, True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ids.cpu().numpy(), ids_e.cpu().numpy())
    np.testing.assert_array_equal(scores.cpu().numpy(), sc_e.cpu().numpy())
    got = ids.cpu().numpy()
    assert (got >= 0).all() and (got < N).all()
    for r, u in enumerate(users):
        assert len(set(got[r].tolist())) == K
        assert not set(got[r].tolist()) & set(consumed.get(int(u), []))
    return sc.last_fallback_rows


@pytest.mark.parametrize("d", [1, 17, 64, 200])
def test_zero_vectors_and_negative_scores(d):
    rng = np.random.default_rng(d)
    n_users, N, K = 40, 3000, 20
    U = np.abs(rng.standard_normal((n_users + 1, d))).astype(np.float32)
    I = -np.abs(rng.standard_normal((N + 1, d))).astype(np.float32)      # every score <= 0
    U[3] = 0.0                                                         # all scores equal (0): pure tie-break by id
    I[100:140] = 0.0                                                   # zero item rows: score exactly 0 = the maximum
    consumed = {u: rng.choice(N, size=5, replace=False).tolist() for u in range(n_users)}
    _run(U, I, N, consumed, list(range(n_users + 1)), K, n_users)


def test_exact_duplicates_across_group_and_tile_borders():
    rng = np.random.default_rng(7)
    n_users, N, d, K = 30, 4096 + 300, 64, 50
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    # copies of strong rows placed on both sides of 8-column and 256-column borders: ties decide the order
    strong = 3.0 * U[:6] / np.linalg.norm(U[:6], axis=1, keepdims=True)
    for j, pos in enumerate([7, 8, 255, 256, 257, 4095, 4096, 2047, 2048, 15, 16, 4103]):
        I[pos] = strong[j % 6]
    consumed = {u: [8, 256, 9] for u in range(0, n_users, 2)}
    _run(U, I, N, consumed, list(range(n_users)), K, n_users)


@pytest.mark.parametrize("B", [1, 2, 127, 129, 255, 257])
def test_batch_sizes_around_the_cluster_padding(B):
    rng = np.random.default_rng(B)
    n_users, N, d, K = 400, 9000, 32, 10
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    consumed = {u: rng.choice(N, size=int(rng.integers(0, 40)), replace=False).tolist() for u in range(n_users)}
    consumed = {u: c for u, c in consumed.items() if c}
    _run(U, I, N, consumed, rng.integers(0, n_users + 1, B).tolist(), K, n_users)


def test_mixed_magnitudes_and_duplicated_consumed_entries():
    rng = np.random.default_rng(11)
    n_users, N, d, K = 64, 20000, 48, 100
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    I[::7] *= 1e4                        # huge and tiny rows in one catalogue (one power-of-two scale for all)
    I[1::7] *= 1e-4
    U[::3] *= 1e3
    consumed = {}
    for u in range(n_users):
        c = rng.choice(N, size=30, replace=False).tolist()
        consumed[u] = c + c[:10]         # duplicates (the reference's lists may hold them)
    _run(U, I, N, consumed, list(range(n_users)), K, n_users)


def test_long_python_list_is_converted_chunk_by_chunk_and_matches_array_input():
    """> 16 384 users as a python list (the reference's calling convention): the host seam converts and uploads the
    ids per launch chunk; the result must equal the one-shot numpy-array input and the exact path."""
    import torch

    from librecommender_b200.engine import EmbedScorer

    rng = np.random.default_rng(21)
    n_users, N, d, K = 60000, 6000, 32, 10
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    consumed = {int(u): rng.choice(N, size=8, replace=False).tolist() for u in rng.choice(n_users, 5000, replace=False)}
    sc = EmbedScorer(U, I, N, consumed, n_users=n_users)
    users = rng.integers(0, n_users + 1, 40001)
    a = sc.recommend(users.tolist(), K, True, False)
    b = sc.recommend(users, K, True, False)
    np.testing.assert_array_equal(a, b)
    e = sc.recommend_exact(torch.as_tensor(users).cuda(), K, True, False).cpu().numpy()
    np.testing.assert_array_equal(a, e)
    assert a.dtype == np.int64 and a.shape == (40001, K)
