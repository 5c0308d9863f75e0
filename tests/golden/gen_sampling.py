"""Golden vectors for the host parity samplers from the UNMODIFIED reference
(libreco/sampling/negatives.py) — numpy PCG64 / Python random streams under fixed seeds.

    python tests/golden/gen_sampling.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from libreco.sampling.negatives import (  # noqa: E402
    neg_probs_from_frequency, negatives_from_popular, negatives_from_random, negatives_from_unconsumed)

OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    data = {}
    gen = np.random.default_rng(123)
    n_items, n_users = 500, 80
    items_pos = gen.integers(0, n_items, size=300)
    users = gen.integers(0, n_users, size=300)
    consumed = {u: gen.choice(n_items, size=int(gen.integers(1, 60)), replace=False).tolist() for u in range(n_users)}
    item_consumed = {i: [] for i in range(n_items)}
    for u, its in consumed.items():
        for i in its:
            item_consumed[i].append(u)
    for i in range(n_items):
        item_consumed[i].append(0)
    seed = 42 % 3407 * 11
    for num_neg in (1, 3):
        rng = np.random.default_rng(seed)
        data[f"random_{num_neg}"] = negatives_from_random(rng, n_items, items_pos, num_neg)
        rng = np.random.default_rng(seed)   # replace=True branch: more samples than items
        data[f"random_big_{num_neg}"] = negatives_from_random(rng, 50, items_pos % 50, num_neg)
        probs = neg_probs_from_frequency(item_consumed, n_items, 0.75)
        rng = np.random.default_rng(seed)
        data[f"popular_{num_neg}"] = negatives_from_popular(rng, n_items, items_pos, num_neg, probs=probs)
        random.seed(seed)
        cs = [set(consumed[u]) for u in range(n_users)]
        data[f"unconsumed_{num_neg}"] = negatives_from_unconsumed(cs, users, items_pos, n_items, num_neg)
    data["probs"] = probs
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u in range(n_users):
        indptr[u + 1] = indptr[u] + len(consumed[u])
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), items_pos=items_pos, users=users, n_items=n_items,
                        n_users=n_users, indptr=indptr,
                        idx=np.concatenate([np.asarray(consumed[u], dtype=np.int32) for u in range(n_users)]),
                        seed=seed, **data)
    print({k: v.shape for k, v in data.items()})
