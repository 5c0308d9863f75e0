"""C1 fixture: the reference's OWN data pipeline on its own sample data
(examples/feat_ranking_example.py:27-41: sample_movielens_merged.csv, chrono split, DatasetFeat with
5 sparse + 1 dense columns) — the resulting DataInfo feature tables and consumed lists, saved so that
the GPU box (no reference tree there) can drive the FM / DeepFM engines with REAL index layouts.

    python tests/golden/gen_movielens_feat.py
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

load_reference()
from libreco.data import DatasetFeat, split_by_ratio_chrono  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    data = pd.read_csv(os.path.join(REFERENCE_ROOT, "examples/sample_data/sample_movielens_merged.csv"))
    train, evald = split_by_ratio_chrono(data, test_size=0.2)
    sparse_col = ["sex", "occupation", "genre1", "genre2", "genre3"]
    dense_col = ["age"]
    user_col = ["sex", "age", "occupation"]
    item_col = ["genre1", "genre2", "genre3"]
    train_data, di = DatasetFeat.build_trainset(train, user_col, item_col, sparse_col, dense_col)
    uc = di.user_consumed
    n_users, n_items = di.n_users, di.n_items
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u in range(n_users):
        indptr[u + 1] = indptr[u] + len(uc[u])
    idx = np.concatenate([np.asarray(uc[u], dtype=np.int32) for u in range(n_users)])
    np.savez_compressed(
        os.path.join(OUT, "movielens_feat.npz"),
        n_users=n_users, n_items=n_items,
        user_sparse_col_index=np.asarray(di.user_sparse_col.index), item_sparse_col_index=np.asarray(di.item_sparse_col.index),
        user_dense_col_index=np.asarray(di.user_dense_col.index), item_dense_col_index=np.asarray(di.item_dense_col.index),
        user_sparse_unique=di.user_sparse_unique, item_sparse_unique=di.item_sparse_unique,
        user_dense_unique=di.user_dense_unique if di.user_dense_unique is not None else np.zeros((0, 0), np.float32),
        item_dense_unique=di.item_dense_unique if di.item_dense_unique is not None else np.zeros((0, 0), np.float32),
        consumed_indptr=indptr, consumed_idx=idx,
        train_users=train_data.user_indices[:4096], train_items=train_data.item_indices[:4096],
        train_sparse=train_data.sparse_indices[:4096], train_dense=train_data.dense_values[:4096],
        sparse_vocab=int(max(di.user_sparse_unique.max(), di.item_sparse_unique.max()) + 1))
    print(n_users, n_items, di.user_sparse_unique.shape, di.item_sparse_unique.shape,
          None if di.user_dense_unique is None else di.user_dense_unique.shape, di.user_sparse_col.index,
          di.item_sparse_col.index, di.user_dense_col.index, di.item_dense_col.index)
