"""Generate golden vectors for the ranking path by running the UNMODIFIED reference
(``/root/reference``, TF/gensim stubbed) in the build container.

    python tests/golden/gen_ranking.py      # writes tests/golden/ranking_*.npz

Inputs are seeded; scores are continuous random floats, so the reference's answer is
unique (no ties) and comparable bit-for-bit at the ID level.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from libreco.recommendation import rank_recommendations, recommend_from_embedding  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def consumed_to_arrays(consumed, n_users):
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u, items in consumed.items():
        indptr[u + 1] = len(items)
    indptr = np.cumsum(indptr)
    idx = np.zeros(int(indptr[-1]), dtype=np.int32)
    for u, items in consumed.items():
        idx[indptr[u]:indptr[u + 1]] = items
    return indptr, idx


def make_consumed(rng, n_users, n_items, mean_len, big_user=None, big_len=None, skip=()):
    consumed = {}
    for u in range(n_users):
        if u in skip:
            continue
        c = int(min(rng.poisson(mean_len), n_items))
        items = rng.choice(n_items, size=c, replace=False).tolist()
        if c >= 4 and u % 3 == 0:  # non-consecutive duplicates survive the Rust dedup
            items.append(items[0])
            items.append(items[2])
        consumed[u] = items
    if big_user is not None:
        consumed[big_user] = rng.choice(n_items, size=big_len, replace=False).tolist()
    return consumed


def case_rank(seed, B, N, K, mean_len, name):
    rng = np.random.default_rng(seed)
    n_users = B + 3
    consumed = make_consumed(rng, n_users, N, mean_len, big_user=1, big_len=N - K + 1, skip=(2,))
    user_ids = rng.permutation(n_users)[:B].tolist()
    if 1 not in user_ids:
        user_ids[0] = 1
    if 2 not in user_ids:
        user_ids[-1] = 2
    preds = rng.standard_normal((B, N)).astype(np.float32)
    ids = rank_recommendations("ranking", user_ids, preds, K, N, consumed, True, False, False)
    ids_nf = rank_recommendations("ranking", user_ids, preds, K, N, consumed, False, False, False)
    ids_s, scores = rank_recommendations("ranking", user_ids, preds.reshape(-1), K, N, consumed,
                                         True, False, True)
    ids_r, scores_r = rank_recommendations("rating", user_ids, preds, K, N, consumed, True, False, True)
    indptr, idx = consumed_to_arrays(consumed, n_users)
    np.savez_compressed(
        os.path.join(OUT, f"ranking_{name}.npz"), user_ids=np.array(user_ids), preds=preds, K=K, N=N,
        indptr=indptr, idx=idx, ids=ids, ids_nofilter=ids_nf, ids_flat=ids_s, scores_ranking=scores,
        ids_rating=ids_r, scores_rating=scores_r)
    print(name, ids.shape)


def case_embed(seed, n_users, N, d, B, K, name):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users + 1, d)) / np.sqrt(d)).astype(np.float32)
    I = (rng.standard_normal((N + 1, d)) / np.sqrt(d)).astype(np.float32)
    consumed = make_consumed(rng, n_users, N, 12)
    user_ids = rng.choice(n_users, size=B, replace=False).tolist()
    user_ids[-1] = n_users  # OOV user row (embed_base.py:153-161 uses it for default_recs)
    model = types.SimpleNamespace(task="ranking", n_items=N, user_consumed=consumed)
    ids = recommend_from_embedding(model, user_ids, K, U, I, True, False)
    ids_nf = recommend_from_embedding(model, user_ids, K, U, I, False, False)
    full = U[user_ids] @ I[:N].T
    indptr, idx = consumed_to_arrays(consumed, n_users)
    np.savez_compressed(
        os.path.join(OUT, f"embed_{name}.npz"), U=U, I=I, user_ids=np.array(user_ids), K=K, N=N,
        indptr=indptr, idx=idx, ids=ids, ids_nofilter=ids_nf, full_scores=full)
    print(name, ids.shape)


if __name__ == "__main__":
    case_rank(11, 9, 517, 10, 20, "small")
    case_rank(12, 4, 20011, 100, 60, "wide")
    case_rank(13, 3, 40, 37, 2, "k_near_n")
    case_embed(21, 60, 333, 16, 17, 20, "d16")
    case_embed(22, 40, 2500, 64, 8, 100, "d64")
    case_embed(23, 30, 200, 7, 5, 50, "d7")
