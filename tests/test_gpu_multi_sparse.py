"""GPU parity of the multi-sparse combiners (reference libreco/tfops/features.py:47-118) against the
numpy restatement in oracle/tf_models.py (TF half: parity unpinned, see its header): the pooling
kernel itself, and FM / DeepFM logits + recommendations on a layout with user- and item-side
multi-sparse fields (padding = the field's OOV index, rows with no sub-feature at all)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=1e-5):
    scale = np.maximum(np.abs(ref), np.abs(ref).mean())
    assert (np.abs(got - ref) <= tol * scale + 1e-6).all(), float(np.abs(got - ref).max())


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("K", [1, 16, 40])
def test_combine_kernel_matches_oracle(combiner, K):
    import torch

    from librecommender_b200 import _lib
    from librecommender_b200.feat_models import _COMBINERS
    from oracle import tf_models as tm

    rng = np.random.default_rng(K)
    V, n, ln, oov = 500, 3000, 5, 499
    table = rng.standard_normal((V, K)).astype(np.float32)
    idx = rng.integers(0, V - 1, size=(n, ln)).astype(np.int32)
    idx[rng.random((n, ln)) < 0.4] = oov
    idx[:5] = oov
    info = dict(field_offset=[0], field_len=[ln], feat_oov=[oov])
    ref = tm.multi_sparse_combine(table.astype(np.float64) if K > 1 else table[:, 0].astype(np.float64),
                                  idx, info, combiner)
    ref = ref[:, 0] if K > 1 else ref
    t, i = torch.from_numpy(table).cuda(), torch.from_numpy(idx).cuda()
    out = torch.empty((n, K), device="cuda")
    _lib.check(_lib.lib.b200_multi_sparse_combine(_lib.ptr(t), K, K, _lib.ptr(i), ln, ln, n, oov, _COMBINERS[combiner],
                                                  _lib.ptr(out), K, _lib.current_stream()))
    got = out.cpu().numpy()
    np.testing.assert_allclose(got, ref.reshape(n, K), rtol=2e-6, atol=2e-6)
    assert (got[:5] == 0).all()                       # div_no_nan rows


def _ms_case(seed):
    from oracle import tf_models as tm

    rng = np.random.default_rng(seed)
    spec = tm.make_multi_sparse_spec(rng, 200, 350, [9, 30], [12, 6, 25],
                                     [("item", 18, 3), ("user", 7, 2), ("item", 40, 4)], 1, 2)
    return rng, spec


@pytest.mark.parametrize("combiner", ["sqrtn", "mean", "sum", "normal"])
def test_fm_and_deepfm_with_multi_sparse_fields(combiner):
    from librecommender_b200.feat_models import FM, DeepFM
    from oracle import tf_models as tm

    rng, spec = _ms_case(3)
    info = spec["multi_sparse_combine_info"]
    reduced = spec["n_sparse"] - (sum(info["field_len"]) - len(info["field_len"]))
    spec_w = dict(spec, n_sparse=reduced if combiner != "normal" else spec["n_sparse"])
    K = 16
    users = rng.integers(0, spec["n_users"] + 1, size=999)
    items = rng.integers(0, spec["n_items"] + 1, size=999)
    sparse, dense = tm.row_features(spec, users, items)
    for make, fwd, cls in ((tm.make_fm_weights, tm.fm_forward, FM), (tm.make_deepfm_weights, tm.deepfm_forward, DeepFM)):
        w = make(rng, spec_w, K)
        w["multi_sparse"] = dict(info, combiner=combiner)
        w["multi_sparse_combiner"] = combiner
        ref64 = fwd(w, users, items, sparse, dense, dtype=np.float64)
        model = cls(spec, w)
        assert model.spec.n_sparse == spec_w["n_sparse"]
        got = model.logits(users, items).cpu().numpy()
        _close(got, ref64)
        # all-items recommendation agrees with the oracle's ranking of its own scores
        uid = rng.integers(0, spec["n_users"], size=12)
        ids = model.recommend(uid, 10, filter_consumed=False)
        ids = ids[0] if isinstance(ids, tuple) else ids
        for r, u in enumerate(uid):
            all_items = np.arange(spec["n_items"])
            sp, de = tm.row_features(spec, np.full(spec["n_items"], u), all_items)
            sc = fwd(w, np.full(spec["n_items"], u), all_items, sp, de, dtype=np.float64)
            kth = np.sort(sc)[-10]
            assert (sc[ids[r]] >= kth - 1e-5 * max(1.0, abs(kth))).all()


def test_default_combiner_is_sqrtn_like_the_reference():
    """deepfm.py:107 / fm.py: multi_sparse_combiner defaults to "sqrtn"."""
    from librecommender_b200.feat_models import FM
    from oracle import tf_models as tm

    rng, spec = _ms_case(5)
    info = spec["multi_sparse_combine_info"]
    reduced = spec["n_sparse"] - (sum(info["field_len"]) - len(info["field_len"]))
    w = tm.make_fm_weights(rng, dict(spec, n_sparse=reduced), 8)
    users = rng.integers(0, spec["n_users"], size=300)
    items = rng.integers(0, spec["n_items"], size=300)
    sparse, dense = tm.row_features(spec, users, items)
    got = FM(spec, w).logits(users, items).cpu().numpy()          # no combiner key given
    w["multi_sparse"] = dict(info, combiner="sqrtn")
    _close(got, tm.fm_forward(w, users, items, sparse, dense, dtype=np.float64))


@pytest.mark.parametrize("name", ["FM", "DeepFM"])
def test_real_movielens_multi_sparse_layout(name):
    """The reference's own DataInfo layout (examples/multi_sparse_example.py columns): engine on the
    unique tables vs the oracle fed with the reference's own per-row index matrix."""
    from librecommender_b200 import feat_models as fm
    from oracle import tf_models as tm
    from _fixtures import load_multi_sparse_spec as load_spec

    g, spec = load_spec()
    info = spec["multi_sparse_combine_info"]
    reduced = spec["n_sparse"] - (sum(info["field_len"]) - len(info["field_len"]))
    rng = np.random.default_rng(7)
    spec_w = dict(spec, n_sparse=reduced)
    if name == "FM":
        w, fwd = tm.make_fm_weights(rng, spec_w, 16, use_bn=False), tm.fm_forward
    else:
        w, fwd = tm.make_deepfm_weights(rng, spec_w, 16, (128, 64, 32), False), tm.deepfm_forward
    w["multi_sparse"] = dict(info, combiner="sqrtn")           # the reference's default combiner
    u, it = g["train_users"], g["train_items"]
    ref = fwd(w, u, it, g["train_sparse"].astype(np.int64), g["train_dense"], dtype=np.float64)
    model = getattr(fm, name)(spec, w)
    assert model.spec.n_sparse == reduced == 3
    got = model.logits(u, it).cpu().numpy()
    # raw "age" up to 56 makes the FM pairwise term cancel in fp32 (see test_gpu_movielens_c1.py)
    Pm, _ = tm._stacked_embeds(tm._cast(w, np.float64), u, it, g["train_sparse"].astype(np.int64),
                               g["train_dense"], np.float64)
    cond = 0.5 * (np.square(Pm.sum(1)) + np.square(Pm).sum(1)).sum(1)
    scale = np.maximum(np.abs(ref), np.abs(ref).mean())
    assert (np.abs(got - ref) <= 1e-5 * scale + 1e-6 * cond + 1e-6).all(), float(np.abs(got - ref).max())
