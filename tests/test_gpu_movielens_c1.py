"""C1 (BASELINE.json configs[0]): FM / DeepFM on the reference's own sample data with the feature
tables produced by the reference's DatasetFeat pipeline (tests/golden/movielens_feat.npz, generated
by tests/golden/gen_movielens_feat.py from examples/sample_data/sample_movielens_merged.csv with the
columns of examples/feat_ranking_example.py:34-41; FM hyper-parameters of :207-221: embed 16, use_bn).
Weights are glorot-uniform (TensorFlow's own init stream cannot be reproduced without TensorFlow):
parity is on the forward pass and on recommend_user given identical weights."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "movielens_feat.npz")


def _spec():
    g = np.load(GOLD)
    spec = dict(
        n_users=int(g["n_users"]), n_items=int(g["n_items"]),
        user_sparse_col_index=g["user_sparse_col_index"].tolist(), item_sparse_col_index=g["item_sparse_col_index"].tolist(),
        user_dense_col_index=g["user_dense_col_index"].tolist(), item_dense_col_index=g["item_dense_col_index"].tolist(),
        user_sparse_unique=g["user_sparse_unique"], item_sparse_unique=g["item_sparse_unique"],
        user_dense_unique=g["user_dense_unique"].astype(np.float32), item_dense_unique=None,
        sparse_vocab=int(g["sparse_vocab"]))
    spec["n_sparse"] = len(spec["user_sparse_col_index"]) + len(spec["item_sparse_col_index"])
    spec["n_dense"] = len(spec["user_dense_col_index"]) + len(spec["item_dense_col_index"])
    consumed = {u: g["consumed_idx"][g["consumed_indptr"][u]:g["consumed_indptr"][u + 1]].tolist()
                for u in range(spec["n_users"])}
    return g, spec, consumed


@pytest.mark.parametrize("name", ["FM", "DeepFM"])
def test_c1_predict_and_recommend(name):
    from librecommender_b200 import feat_models as fm
    from oracle import ranking as orc
    from oracle import tf_models as tm

    g, spec, consumed = _spec()
    rng = np.random.default_rng(42)
    if name == "FM":
        w, fwd = tm.make_fm_weights(rng, spec, 16, use_bn=True), tm.fm_forward
    else:
        w, fwd = tm.make_deepfm_weights(rng, spec, 16, (128, 64, 32), True), tm.deepfm_forward
    model = getattr(fm, name)(spec, w, consumed)
    # predict on real training rows with their own feature rows (predict_tf_feat)
    u, it = g["train_users"], g["train_items"]
    ref = fwd(w, u, it, g["train_sparse"].astype(np.int64), g["train_dense"], dtype=np.float64)
    ref32 = fwd(w, u, it, g["train_sparse"].astype(np.int64), g["train_dense"], dtype=np.float32)
    got = model.logits(u, it, sparse_rows=g["train_sparse"], dense_rows=g["train_dense"]).cpu().numpy()
    scale = np.maximum(np.abs(ref), np.abs(ref).mean())
    # raw "age" (1..56) multiplies a whole embedding row: 0.5((sum e)^2 - sum e^2) cancels — in fp32
    # the REFERENCE graph itself is only good to ~2e-5 relative here (ref32 vs ref64).  Tolerance =
    # 1e-5 relative + the fp32 rounding scale of the cancelling sums (1e-6 x their magnitude).
    Pm, _ = tm._stacked_embeds(tm._cast(w, np.float64), u, it, g["train_sparse"].astype(np.int64),
                               g["train_dense"], np.float64)
    cond = 0.5 * (np.square(Pm.sum(1)) + np.square(Pm).sum(1)).sum(1)
    assert np.abs(ref32 - ref).max() > 1e-5 * np.abs(ref).mean() * 0.5        # documents the conditioning
    assert (np.abs(got - ref) <= 1e-5 * scale + 1e-6 * cond + 1e-6).all(), float(np.abs(got - ref).max())
    # recommend_user for a slice of users, n_rec = 7 (examples) and 100
    users = np.arange(0, spec["n_users"], 37)
    N = spec["n_items"]
    uu, ii = np.repeat(users, N), np.tile(np.arange(N), len(users))
    sp, de = tm.row_features(spec, uu, ii)
    preds = fwd(w, uu, ii, sp, de, dtype=np.float64).astype(np.float32)
    for n_rec in (7, 100):
        got_ids = model.recommend(users, n_rec, True)
        ref_ids = orc.rank_recommendations("ranking", users.tolist(), preds, n_rec, N, consumed, True)
        assert orc.near_tie_mask(ref_ids, got_ids, preds.reshape(len(users), N), 1e-5).all()
        assert (got_ids == ref_ids).mean() > 0.97
        for r, usr in enumerate(users.tolist()):
            assert not set(got_ids[r].tolist()) & set(consumed[usr])
