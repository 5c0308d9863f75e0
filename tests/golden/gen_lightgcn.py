"""Golden vectors for LightGCN propagation from the UNMODIFIED reference module
(libreco/algorithms/torch_modules/lightgcn_module.py) on CPU.

    python tests/golden/gen_lightgcn.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
import torch  # noqa: E402
from libreco.algorithms.torch_modules.lightgcn_module import LightGCNModel  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(seed, n_users, n_items, d, n_layers, mean_deg, name):
    rng = np.random.default_rng(seed)
    consumed = {}
    w = 1.0 / np.arange(1, n_items + 1)
    w /= w.sum()
    for u in range(n_users):
        c = max(1, int(min(rng.poisson(mean_deg), n_items)))
        items = rng.choice(n_items, size=c, replace=False, p=w).tolist()
        if c > 3:
            items.append(items[0])           # duplicates collapse to one edge
        consumed[u] = items
    consumed[n_users - 1] = []               # isolated user -> zero row
    torch.manual_seed(seed)
    m = LightGCNModel(n_users, n_items, d, n_layers, 0.0, consumed, torch.device("cpu"))
    with torch.no_grad():
        ue, ie = m.embedding_propagation(use_dropout=False)
    lap = m.laplacian_matrix.coalesce()
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u in range(n_users):
        indptr[u + 1] = indptr[u] + len(consumed[u])
    idx = np.concatenate([np.asarray(consumed[u], dtype=np.int32) for u in range(n_users)])
    np.savez_compressed(
        os.path.join(OUT, f"lightgcn_{name}.npz"), n_users=n_users, n_items=n_items, n_layers=n_layers,
        indptr=indptr, idx=idx,
        user_init=m.user_init_embeds.weight.detach().numpy(), item_init=m.item_init_embeds.weight.detach().numpy(),
        user_out=ue.numpy(), item_out=ie.numpy(),
        lap_row=lap.indices()[0].numpy(), lap_col=lap.indices()[1].numpy(), lap_val=lap.values().numpy())
    print(name, ue.shape, ie.shape, lap._nnz())


if __name__ == "__main__":
    case(31, 200, 120, 16, 3, 8, "d16")
    case(32, 150, 400, 64, 2, 15, "d64")
    case(33, 60, 50, 10, 4, 5, "d10")
