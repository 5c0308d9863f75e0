"""WideDeep on the DeepFM engine: the weight mapping (feat_models.wide_deep_weights) is checked on the CPU by pushing
the mapped dict through the DeepFM restatement and the original variables through the WideDeep restatement
(oracle.tf_models: deepfm.py:155-174 vs wide_deep.py:150-176)."""
import numpy as np

from oracle import tf_models as tm


def test_mapping_reproduces_wide_deep_logits():
    from librecommender_b200.feat_models import wide_deep_weights

    rng = np.random.default_rng(3)
    spec = tm.make_spec(rng, 50, 80, [7, 30], [11, 5], 1, 2)
    base = tm.make_deepfm_weights(rng, spec, 8, (16, 8), True)
    wd = dict(user_wide=base["user_linear"], item_wide=base["item_linear"], sparse_wide=base["sparse_linear"],
              dense_wide=base["dense_linear"], wide_kernel=base["lin_kernel"], wide_bias=np.float32(0.03),
              user_deep=base["user_embeds"], item_deep=base["item_embeds"], sparse_deep=base["sparse_embeds"],
              dense_deep=base["dense_embeds"], mlp=base["mlp"],
              deep_kernel=rng.standard_normal(8).astype(np.float32), deep_bias=np.float32(-0.02))
    w = wide_deep_weights(**wd)
    users, items = rng.integers(0, 50, 300), rng.integers(0, 80, 300)
    sparse, dense = tm.row_features(spec, users, items)
    a = tm.wide_deep_forward(wd, users, items, sparse, dense, dtype=np.float64)
    b = tm.deepfm_forward(w, users, items, sparse, dense, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    assert w["out_kernel"].shape == (1 + 8 + 8,) and w["out_kernel"][0] == 1.0 and (w["out_kernel"][1:9] == 0).all()
