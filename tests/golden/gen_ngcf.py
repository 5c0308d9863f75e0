"""Golden vectors for NGCF propagation from the UNMODIFIED reference module
(libreco/algorithms/torch_modules/ngcf_module.py) on CPU.

    python tests/golden/gen_ngcf.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
import torch  # noqa: E402
from libreco.algorithms.torch_modules.ngcf_module import NGCFModel  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(seed, n_users, n_items, d, layers, mean_deg, name):
    rng = np.random.default_rng(seed)
    consumed = {}
    w = 1.0 / np.arange(1, n_items + 1)
    w /= w.sum()
    for u in range(n_users):
        c = max(1, int(min(rng.poisson(mean_deg), n_items)))
        items = rng.choice(n_items, size=c, replace=False, p=w).tolist()
        if c > 3:
            items.append(items[0])
        consumed[u] = items
    consumed[n_users - 1] = []               # isolated user: only its self loop
    torch.manual_seed(seed)
    m = NGCFModel(n_users, n_items, d, list(layers), 0.0, 0.0, consumed, torch.device("cpu"))
    with torch.no_grad():
        for k in range(len(layers)):         # non-zero biases so that they are exercised
            m.weight_dict[f"b_self_{k}"].normal_(0, 0.05)
            m.weight_dict[f"b_pair_{k}"].normal_(0, 0.05)
        ue, ie = m.embedding_propagation(use_dropout=False)
    lap = m.laplacian_matrix.coalesce()
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u in range(n_users):
        indptr[u + 1] = indptr[u] + len(consumed[u])
    idx = np.concatenate([np.asarray(consumed[u], dtype=np.int32) for u in range(n_users)])
    data = dict(n_users=n_users, n_items=n_items, indptr=indptr, idx=idx,
                user_embed=m.embedding_dict["user_embed"].detach().numpy(),
                item_embed=m.embedding_dict["item_embed"].detach().numpy(),
                user_out=ue.numpy(), item_out=ie.numpy(), lap_row=lap.indices()[0].numpy(),
                lap_col=lap.indices()[1].numpy(), lap_val=lap.values().numpy())
    for k, v in m.weight_dict.items():
        data[k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, f"ngcf_{name}.npz"), **data)
    print(name, ue.shape, ie.shape, lap._nnz())


if __name__ == "__main__":
    case(41, 120, 90, 16, (16, 16, 16), 7, "d16")
    case(42, 80, 150, 64, (64, 32), 12, "d64")
