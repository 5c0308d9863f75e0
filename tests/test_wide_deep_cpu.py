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


def test_load_reference_wide_deep_from_npz(tmp_path):
    """A synthetic ``<name>_tf_variables.npz`` with the reference's WideDeep variable names -> engine weights."""
    from librecommender_b200.weights_io import load_reference_wide_deep

    rng = np.random.default_rng(5)
    spec = tm.make_spec(rng, 30, 40, [7], [11, 5], 1, 1)
    base = tm.make_deepfm_weights(rng, spec, 8, (16, 8), True)
    mlp = base["mlp"]
    arrays = {"embedding/user_wide_var:0": base["user_linear"].reshape(-1, 1), "embedding/item_wide_var:0": base["item_linear"].reshape(-1, 1),
              "embedding/sparse_wide_var:0": base["sparse_linear"].reshape(-1, 1), "embedding/dense_wide_var:0": base["dense_linear"].reshape(-1, 1),
              "embedding/user_deep_var:0": base["user_embeds"], "embedding/item_deep_var:0": base["item_embeds"],
              "embedding/sparse_deep_var:0": base["sparse_embeds"], "embedding/dense_deep_var:0": base["dense_embeds"],
              "wide_term/kernel:0": base["lin_kernel"].reshape(-1, 1), "wide_term/bias:0": np.array([0.04], np.float32),
              "deep_term/kernel:0": rng.standard_normal((8, 1)).astype(np.float32), "deep_term/bias:0": np.array([-0.01], np.float32)}
    for i in range(2):
        arrays[f"deep/deep_layer{i + 1}/kernel:0"] = mlp["kernels"][i]
        arrays[f"deep/deep_layer{i + 1}/bias:0"] = mlp["biases"][i]
    for j, bn in enumerate([mlp["bn_in"]] + list(mlp["bns"])):
        scope = "deep/batch_normalization" + ("" if j == 0 else f"_{j}")
        for k, v in (("gamma", bn["gamma"]), ("beta", bn["beta"]), ("moving_mean", bn["mean"]), ("moving_variance", bn["var"])):
            arrays[f"{scope}/{k}:0"] = v
    np.savez(tmp_path / "wd_tf_variables.npz", **arrays)
    w = load_reference_wide_deep(str(tmp_path), "wd", 2, True)
    users, items = rng.integers(0, 30, 100), rng.integers(0, 40, 100)
    sparse, dense = tm.row_features(spec, users, items)
    wd = dict(user_wide=base["user_linear"], item_wide=base["item_linear"], sparse_wide=base["sparse_linear"],
              dense_wide=base["dense_linear"], wide_kernel=base["lin_kernel"], wide_bias=np.float32(0.04),
              user_deep=base["user_embeds"], item_deep=base["item_embeds"], sparse_deep=base["sparse_embeds"],
              dense_deep=base["dense_embeds"], mlp=mlp, deep_kernel=arrays["deep_term/kernel:0"].reshape(-1),
              deep_bias=np.float32(-0.01))
    np.testing.assert_allclose(tm.deepfm_forward(w, users, items, sparse, dense, dtype=np.float64),
                               tm.wide_deep_forward(wd, users, items, sparse, dense, dtype=np.float64), rtol=1e-12, atol=1e-12)


def test_set_regularisation_validates_like_the_reference():
    """training.set_regularisation mirrors tfops/configs.py:20-26 (reg must be a positive float) — host logic only."""
    import types

    import pytest

    from librecommender_b200.training import set_regularisation

    tr = types.SimpleNamespace()
    set_regularisation(tr, reg=1e-3, lr_decay=True, decay_steps=50, decay_rate=0.9)
    assert tr.reg == 1e-3 and tr.decay_steps == 50 and tr.decay_rate == 0.9
    set_regularisation(tr, reg=None, lr_decay=False, decay_steps=50)
    assert tr.reg == 0.0 and tr.decay_steps == 0
    for bad in (-1.0, 0.0, 1):
        if bad == 0.0:
            continue                      # falsy reg = no regulariser (reg_config returns None)
        with pytest.raises(ValueError, match="reg must be float and positive"):
            set_regularisation(tr, reg=bad)
