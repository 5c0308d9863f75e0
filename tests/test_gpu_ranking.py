"""GPU parity: rank_recommendations / recommend_from_embedding through the C-ABI vs the oracle
and the reference-generated golden vectors.  IDs bit-exact (no ties in these inputs); scores
within 1e-5 relative."""
import glob
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _dict_from_csr(indptr, idx):
    return {u: idx[indptr[u]:indptr[u + 1]].tolist()
            for u in range(len(indptr) - 1) if indptr[u + 1] > indptr[u]}


def test_reference_known_answers():
    # reference tests/test_rank_reco.py:7-87
    from librecommender_b200 import rank_recommendations

    user_ids = [1, 2]
    preds = np.array([-0.1, -0.01, 0, 0.1, 0.01, 1, -2, 4, 5, 6])
    consumed = {1: [3, 4], 2: [4]}
    with pytest.raises(ValueError):
        rank_recommendations("ranking", user_ids, preds, 12, 5, consumed, True, False, False)
    ids = rank_recommendations("ranking", user_ids, preds, 2, 5, consumed, True, False, False)
    assert ids.shape == (2, 2) and ids.dtype == np.int64
    np.testing.assert_array_equal(ids, [[2, 1], [3, 2]])
    ids = rank_recommendations("ranking", user_ids, preds, 4, 5, consumed, True, False, False)
    np.testing.assert_array_equal(ids, [[3, 4, 2, 1], [3, 2, 0, 1]])
    _, scores = rank_recommendations("ranking", user_ids, preds, 2, 5, consumed, True, False, True)
    assert scores.shape == (2, 2) and (np.diff(scores, axis=1) <= 0).all()
    ids = rank_recommendations("ranking", user_ids, preds.reshape(2, 5), 2, 5, consumed, True, False, False)
    np.testing.assert_array_equal(ids, [[2, 1], [3, 2]])


def test_random_rec_invariants():
    # reference tests/test_rank_reco.py:90-143
    from librecommender_b200 import rank_recommendations

    user_ids = [1, 2]
    preds = np.array([-0.1, -1e8, 0, 0.1, 0.01, 1e8, -0.01, 1e7, 0.1, 0.01])
    consumed = {1: [3, 4], 2: [4]}
    for _ in range(5):
        rec = rank_recommendations("ranking", user_ids, preds, 2, 5, consumed, True, True, False)
        assert rec.shape == (2, 2)
        assert 0 in rec[0] and 2 in rec[0] and 0 in rec[1]
        rec = rank_recommendations("ranking", user_ids, preds, 4, 5, consumed, True, True, False)
        assert rec.shape == (2, 4)
        assert 1 not in rec[0] and 1 in rec[1]
        _, sc = rank_recommendations("ranking", user_ids, preds, 2, 5, consumed, True, True, True)
        assert (np.diff(sc, axis=1) <= 0).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ranking_*.npz"))))
def test_rank_golden(path):
    from librecommender_b200 import rank_recommendations

    g = np.load(path)
    consumed = _dict_from_csr(g["indptr"], g["idx"])
    uids = g["user_ids"].tolist()
    K, N = int(g["K"]), int(g["N"])
    np.testing.assert_array_equal(
        rank_recommendations("ranking", uids, g["preds"], K, N, consumed, True, False, False), g["ids"])
    np.testing.assert_array_equal(
        rank_recommendations("ranking", uids, g["preds"], K, N, consumed, False, False, False),
        g["ids_nofilter"])
    ids, sc = rank_recommendations("ranking", uids, g["preds"].reshape(-1), K, N, consumed, True, False, True)
    np.testing.assert_array_equal(ids, g["ids_flat"])
    np.testing.assert_allclose(sc, g["scores_ranking"], rtol=1e-5)
    ids, sc = rank_recommendations("rating", uids, g["preds"], K, N, consumed, True, False, True)
    np.testing.assert_array_equal(ids, g["ids_rating"])
    np.testing.assert_array_equal(sc, g["scores_rating"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "embed_*.npz"))))
def test_embed_golden(path):
    from librecommender_b200 import recommend_from_embedding
    from oracle.ranking import near_tie_mask

    g = np.load(path)
    consumed = _dict_from_csr(g["indptr"], g["idx"])
    uids = g["user_ids"].tolist()
    K, N = int(g["K"]), int(g["N"])
    U, I = g["U"], g["I"]
    model = types.SimpleNamespace(task="ranking", n_items=N, user_consumed=consumed,
                                  n_users=U.shape[0] - 1)
    for flt, key in ((True, "ids"), (False, "ids_nofilter")):
        got = recommend_from_embedding(model, uids, K, U, I, flt, False)
        assert got.dtype == np.int64 and got.shape == g[key].shape
        assert near_tie_mask(g[key], got, g["full_scores"], rel_tol=1e-6).all()
        assert (got == g[key]).mean() > 0.999


def test_ties_lowest_id_first_and_all_equal():
    from librecommender_b200 import rank_recommendations
    from oracle.ranking import rank_recommendations as orc_rank

    N = 20000
    preds = np.zeros((3, N), dtype=np.float32)            # every score equal
    preds[1, ::7] = 1.0                                    # many ties at the top
    preds[2] = np.repeat(np.arange(N // 4, dtype=np.float32), 4)[::-1]
    consumed = {0: [0, 1, 5], 1: [7, 14], 2: []}
    got = rank_recommendations("rating", [0, 1, 2], preds, 50, N, consumed, True, False, False)
    ref = orc_rank("rating", [0, 1, 2], preds, 50, N, consumed, True)
    np.testing.assert_array_equal(got, ref)


def test_large_random_vs_oracle():
    from librecommender_b200 import rank_recommendations
    from oracle.ranking import rank_recommendations as orc_rank

    rng = np.random.default_rng(77)
    B, N, K = 5, 300_007, 2000
    preds = rng.standard_normal((B, N)).astype(np.float32)
    consumed = {u: rng.choice(N, size=int(rng.integers(1, 3000)), replace=False).tolist() for u in range(B)}
    ids, sc = rank_recommendations("rating", list(range(B)), preds, K, N, consumed, True, False, True)
    rid, rsc = orc_rank("rating", list(range(B)), preds, K, N, consumed, True, True)
    np.testing.assert_array_equal(ids, rid)
    np.testing.assert_array_equal(sc, rsc)
    for u in range(B):
        assert not set(ids[u].tolist()) & set(consumed[u])


def test_predict_from_embedding_gather_dot():
    from librecommender_b200.engine import EmbedScorer
    from oracle.ranking import predict_from_embedding as orc_pred

    rng = np.random.default_rng(9)
    U = rng.standard_normal((101, 16)).astype(np.float32)
    I = rng.standard_normal((57, 16)).astype(np.float32)
    users = rng.integers(0, 101, size=1000)
    items = rng.integers(0, 57, size=1000)
    sc = EmbedScorer(U, I, 56)
    got = sc.predict(users, items, mode=1)
    np.testing.assert_allclose(got, orc_pred(U, I, users, items), rtol=1e-5, atol=1e-7)
    got = sc.predict(users, items, mode=2, lo=1.0, hi=5.0)
    np.testing.assert_allclose(got, orc_pred(U, I, users, items, "rating", 1.0, 5.0), rtol=1e-5)
