"""Multi-sparse fixture from the reference's OWN pipeline (examples/multi_sparse_example.py:11-25):
sample_movielens_merged.csv with genre1..3 declared as one multi-sparse field, pad value "missing".
Stores the DataInfo layout (unique tables, column indices, MultiSparseInfo) so that the GPU box can
run the engines on the real index layout, plus a CPU check that `_spec_get` reads a live DataInfo.

    python tests/golden/gen_movielens_multi_sparse.py
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


def build():
    data = pd.read_csv(os.path.join(REFERENCE_ROOT, "examples/sample_data/sample_movielens_merged.csv"))
    train, _ = split_by_ratio_chrono(data, test_size=0.2)
    return DatasetFeat.build_trainset(
        train_data=train, user_col=["sex", "age", "occupation"], item_col=["genre1", "genre2", "genre3"],
        sparse_col=["sex", "occupation"], dense_col=["age"], multi_sparse_col=[["genre1", "genre2", "genre3"]],
        pad_val=["missing"])


if __name__ == "__main__":
    train_data, di = build()
    info = di.multi_sparse_combine_info
    np.savez_compressed(
        os.path.join(OUT, "movielens_multi_sparse.npz"),
        n_users=di.n_users, n_items=di.n_items,
        user_sparse_col_index=np.asarray(di.user_sparse_col.index), item_sparse_col_index=np.asarray(di.item_sparse_col.index),
        user_dense_col_index=np.asarray(di.user_dense_col.index), item_dense_col_index=np.asarray(di.item_dense_col.index),
        user_sparse_unique=di.user_sparse_unique, item_sparse_unique=di.item_sparse_unique,
        user_dense_unique=di.user_dense_unique,
        field_offset=np.asarray(info.field_offset), field_len=np.asarray(info.field_len),
        feat_oov=np.asarray(info.feat_oov),
        train_users=train_data.user_indices[:2048], train_items=train_data.item_indices[:2048],
        train_sparse=train_data.sparse_indices[:2048], train_dense=train_data.dense_values[:2048],
        sparse_vocab=int(max(di.user_sparse_unique.max(), di.item_sparse_unique.max()) + 1))
    print(di.n_users, di.n_items, info, di.user_sparse_col.index, di.item_sparse_col.index,
          di.item_sparse_unique[:3], di.item_sparse_unique[-1])
