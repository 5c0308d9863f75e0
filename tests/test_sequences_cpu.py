"""Host-side sequence builders against golden vectors of the unmodified reference
(libreco/batch/sequence.py via tests/golden/gen_sequences.py)."""
import os
import random

import numpy as np

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sequences.npz"))


def _consumed():
    indptr, idx = G["indptr"], G["idx"]
    return {u: idx[indptr[u]:indptr[u + 1]].tolist() for u in range(len(indptr) - 1)}


def test_recent_sequences_match_reference():
    from librecommender_b200.consumed import ConsumedCSR
    from librecommender_b200.feat_models import recent_sequences, recent_sequences_csr

    cons = _consumed()
    n_users, n_items = len(cons), int(G["n_items"])
    csr = ConsumedCSR(G["indptr"], G["idx"])
    for L in (5, 10, 40):
        for seqs, lens in (recent_sequences(cons, n_users, n_items, L), recent_sequences_csr(csr, n_items, L)):
            assert seqs.dtype == np.int32 and lens.dtype == np.int32
            np.testing.assert_array_equal(seqs, G[f"recent_{L}"])
            np.testing.assert_array_equal(lens, G[f"recent_lens_{L}"])


def test_interacted_positions_consume_the_reference_random_stream():
    """The parity helper draws exactly one random.randrange per not-in-history sample, in order:
    replaying it reproduces the reference's windows for those samples."""
    from librecommender_b200.collate import interacted_positions_host

    cons = _consumed()
    users, items = G["users"], G["items"]
    random.seed(1234)
    pos = interacted_positions_host(cons, users, items)
    L = 10
    seqs, lens = G[f"seqs_{L}"], G[f"lens_{L}"]
    n_items = int(G["n_items"])
    checked = 0
    for j in np.nonzero(pos >= 0)[0]:
        p = int(pos[j])
        hist = cons[int(users[j])]
        want = hist[max(0, p - L):p]
        assert lens[j] == (1 if p == 0 else min(p, L))
        np.testing.assert_array_equal(seqs[j, :len(want)], want)
        assert (seqs[j, len(want):] == n_items).all()
        checked += 1
    assert checked > 50
