"""GPU parity of b200_interacted_seqs (collate-time history windows) against golden vectors of the
unmodified reference (libreco/batch/sequence.py:33-71, mode "recent") — bit-exact in parity mode."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sequences.npz"))


def _setup():
    import torch

    from librecommender_b200.consumed import ConsumedCSR

    csr = ConsumedCSR(G["indptr"], G["idx"])
    users = torch.tensor(G["users"], device="cuda")
    items = torch.tensor(G["items"], device="cuda")
    cons = {u: G["idx"][G["indptr"][u]:G["indptr"][u + 1]].tolist() for u in range(len(G["indptr"]) - 1)}
    return csr, cons, users, items


@pytest.mark.parametrize("L", [5, 10, 40])
def test_interacted_seqs_parity_mode_bit_exact(L):
    import torch

    from librecommender_b200.collate import DeviceSequenceBuilder, interacted_positions_host

    csr, cons, users, items = _setup()
    random.seed(1234)
    pos = interacted_positions_host(cons, G["users"], G["items"])
    b = DeviceSequenceBuilder(csr, L, int(G["n_items"]))
    seqs, lens = b(users, items, torch.tensor(pos, device="cuda"))
    np.testing.assert_array_equal(seqs.cpu().numpy(), G[f"seqs_{L}"])
    np.testing.assert_array_equal(lens.cpu().numpy(), G[f"lens_{L}"])
    assert seqs.dtype == torch.int32 and lens.dtype == torch.int32


def test_interacted_seqs_fast_mode_invariants():
    """Philox positions for negatives: samples whose item IS in the history are identical to the
    reference; the others hold a valid prefix window of the user's history; deterministic per
    (seed, step)."""
    import torch

    from librecommender_b200.collate import DeviceSequenceBuilder

    csr, cons, users, items = _setup()
    L, n_items = 10, int(G["n_items"])
    b1, b2 = DeviceSequenceBuilder(csr, L, n_items, seed=7), DeviceSequenceBuilder(csr, L, n_items, seed=7)
    s1, l1 = b1(users, items)
    s2, l2 = b2(users, items)
    assert torch.equal(s1, s2) and torch.equal(l1, l2)
    s1, l1 = s1.cpu().numpy(), l1.cpu().numpy()
    in_hist = np.array([int(i) in set(cons[int(u)]) for u, i in zip(G["users"], G["items"])])
    np.testing.assert_array_equal(s1[in_hist], G[f"seqs_{L}"][in_hist])
    np.testing.assert_array_equal(l1[in_hist], G[f"lens_{L}"][in_hist])
    for j in np.nonzero(~in_hist)[0]:
        hist = cons[int(G["users"][j])]
        ln = int(l1[j])
        w = s1[j][s1[j] != n_items].tolist() if ln > 1 or s1[j, 0] != n_items else []
        assert 1 <= ln <= L and len(w) in (ln, 0)
        if w:   # a contiguous window of the history ending before some position
            assert any(hist[k:k + len(w)] == w for k in range(len(hist)))
    s3, _ = b1(users, items)       # next step -> different draws somewhere
    assert not np.array_equal(s3.cpu().numpy()[~in_hist], s1[~in_hist])


def test_interacted_seqs_large_random_against_python_restatement():
    import torch

    from librecommender_b200.collate import DeviceSequenceBuilder
    from librecommender_b200.consumed import ConsumedCSR

    rng = np.random.default_rng(3)
    n_users, n_items, L = 500, 3000, 50
    lens_u = rng.integers(1, 400, n_users)
    indptr = np.concatenate([[0], np.cumsum(lens_u)]).astype(np.int64)
    idx = rng.integers(0, n_items, indptr[-1]).astype(np.int32)
    csr = ConsumedCSR(indptr, idx)
    n = 20000
    users = rng.integers(0, n_users, n)
    offs = (rng.random(n) * lens_u[users]).astype(np.int64)
    items = idx[indptr[users] + offs].astype(np.int64)
    b = DeviceSequenceBuilder(csr, L, n_items)
    seqs, lens = b(torch.tensor(users, device="cuda"), torch.tensor(items, device="cuda"))
    seqs, lens = seqs.cpu().numpy(), lens.cpu().numpy()
    for j in rng.integers(0, n, 2000):
        hist = idx[indptr[users[j]]:indptr[users[j] + 1]].tolist()
        p = hist.index(int(items[j]))
        want = hist[max(0, p - L):p]
        assert lens[j] == (1 if p == 0 else min(p, L))
        assert seqs[j, :len(want)].tolist() == want and (seqs[j, len(want):] == n_items).all()
