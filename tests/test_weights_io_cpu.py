"""Round trips through the reference's on-disk weight formats (utils/save_load.py:39-98,
bases/embed_base.py:289-330); with the reference mounted, its own loader reads our files."""
import os

import numpy as np
import pytest

from oracle.ref_loader import load_reference, reference_available


def test_embed_and_tf_variable_roundtrip(tmp_path):
    from librecommender_b200 import weights_io as io

    rng = np.random.default_rng(0)
    U, I = rng.standard_normal((11, 8)).astype(np.float32), rng.standard_normal((7, 8)).astype(np.float32)
    io.save_embed_model(str(tmp_path), "m", U, I)
    u2, i2 = io.load_embed_model(str(tmp_path), "m")
    np.testing.assert_array_equal(u2, U)
    np.testing.assert_array_equal(i2, I)
    raw = np.load(os.path.join(tmp_path, "m.npz"))
    assert set(raw.files) == {"user_embed", "item_embed"}            # the keys EmbedBase.load reads

    w = dict(user_embeds=U, item_embeds=I, sparse_embeds=rng.standard_normal((30, 8)).astype(np.float32),
             user_linear=rng.standard_normal(11).astype(np.float32), item_linear=rng.standard_normal(7).astype(np.float32),
             sparse_linear=rng.standard_normal(30).astype(np.float32), lin_kernel=rng.standard_normal(5).astype(np.float32))
    io.save_tf_variables(str(tmp_path), "fm", w, extra_names={"lin_kernel": "dense/kernel:0"})
    raw = np.load(os.path.join(tmp_path, "fm_tf_variables.npz"))
    # shapes = the reference's var_shape values (fm.py:181-249): id tables [V, 1], feature tables 1-D
    assert raw["embedding/user_linear_var:0"].shape == (11, 1)
    assert raw["embedding/item_linear_var:0"].shape == (7, 1)
    assert raw["embedding/sparse_linear_var:0"].shape == (30,)
    back = io.load_tf_variables(str(tmp_path), "fm", extra_names={"lin_kernel": "dense/kernel:0"})
    for k in w:
        np.testing.assert_array_equal(np.asarray(back[k]).reshape(w[k].shape), w[k])
    io.save_default_recs(str(tmp_path), "m", np.arange(20))
    np.testing.assert_array_equal(io.load_default_recs(str(tmp_path), "m"), np.arange(20))


@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_reference_loader_reads_our_default_recs(tmp_path):
    load_reference()
    from libreco.utils.save_load import load_default_recs as ref_load

    from librecommender_b200 import weights_io as io

    io.save_default_recs(str(tmp_path), "m", np.arange(2000))
    np.testing.assert_array_equal(ref_load(str(tmp_path), "m"), np.arange(2000))
