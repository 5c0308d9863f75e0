"""The NGCF Laplacian builder is pure index arithmetic on torch tensors: run it on the CPU device and
pin it bit-for-bit against the reference module's matrix (tests/golden/ngcf_*.npz)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["ngcf_d16.npz", "ngcf_d64.npz"])
def test_ngcf_laplacian_bit_equal_on_cpu(name):
    from librecommender_b200.consumed import ConsumedCSR
    from librecommender_b200.ngcf import build_ngcf_laplacian_csr

    g = np.load(os.path.join(GOLD, name))
    n_users, n_items = int(g["n_users"]), int(g["n_items"])
    indptr, col, val = build_ngcf_laplacian_csr(ConsumedCSR(g["indptr"], g["idx"]), n_users, n_items,
                                                device=torch.device("cpu"))
    rows = np.repeat(np.arange(n_users + n_items), np.diff(indptr.numpy()))
    order = np.lexsort((g["lap_col"], g["lap_row"]))
    np.testing.assert_array_equal(rows, g["lap_row"][order])
    np.testing.assert_array_equal(col.numpy(), g["lap_col"][order])
    np.testing.assert_array_equal(val.numpy(), g["lap_val"][order])
    # every row sums to one (row-normalised with self loops), isolated nodes keep only the self loop
    sums = np.bincount(rows, weights=val.numpy().astype(np.float64), minlength=n_users + n_items)
    np.testing.assert_allclose(sums, 1.0, rtol=1e-6)
