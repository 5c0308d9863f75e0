"""Golden vectors for the collate-time behaviour sequences from the UNMODIFIED reference
(libreco/batch/sequence.py: get_interacted_seqs mode="recent", get_recent_seqs).

    python tests/golden/gen_sequences.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from libreco.batch.sequence import get_interacted_seqs, get_recent_seqs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    g = np.random.default_rng(77)
    n_users, n_items = 60, 200
    consumed = {}
    for u in range(n_users):
        ln = int(g.integers(1, 45))
        its = g.integers(0, n_items, ln).tolist()         # repeats allowed (not consecutive-equal)
        its = [x for k, x in enumerate(its) if k == 0 or x != its[k - 1]]
        consumed[u] = its
    n = 700
    users = g.integers(0, n_users, n)
    items = np.array([consumed[u][int(g.integers(0, len(consumed[u])))] if g.random() < 0.6
                      else int(g.integers(0, n_items)) for u in users])
    sets = {u: set(v) for u, v in consumed.items()}
    data = {"users": users, "items": items}
    for L in (5, 10, 40):
        random.seed(1234)
        seqs, lens = get_interacted_seqs(users, items, consumed, n_items, "recent", L, sets, None)
        data[f"seqs_{L}"], data[f"lens_{L}"] = seqs, lens
        rs, rl = get_recent_seqs(n_users, consumed, n_items, L)
        data[f"recent_{L}"], data[f"recent_lens_{L}"] = rs, rl
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    for u in range(n_users):
        indptr[u + 1] = indptr[u] + len(consumed[u])
    data["indptr"] = indptr
    data["idx"] = np.concatenate([np.asarray(consumed[u], dtype=np.int32) for u in range(n_users)])
    data["n_items"] = n_items
    np.savez_compressed(os.path.join(OUT, "sequences.npz"), **data)
    print("wrote sequences.npz", {k: np.shape(v) for k, v in data.items()})
