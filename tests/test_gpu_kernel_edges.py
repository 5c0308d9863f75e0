"""Edge shapes of the secondary C-ABI kernels against plain numpy / torch references: empty inputs, single rows,
K == N selections, rows with no non-zeros and rows far above the long-row threshold in the SpMM, sequence pooling
with empty / full / all-pad sequences, the column reduction's cluster path on awkward shapes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_spmm_empty_rows_and_very_long_rows():
    import torch

    from librecommender_b200.lightgcn import SpmmGraph

    rng = np.random.default_rng(0)
    n, d = 3000, 64
    deg = rng.integers(0, 6, n)
    deg[::97] = 0                     # empty rows
    deg[5] = 9000                     # > 8 chunks of the long-row path
    deg[1234] = 1025                  # just above the threshold
    deg[2999] = 1024                  # exactly at the threshold (short path)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, n, indptr[-1]).astype(np.int32)
    val = rng.standard_normal(indptr[-1]).astype(np.float32)
    E = rng.standard_normal((n, d)).astype(np.float32)
    g = SpmmGraph(torch.as_tensor(indptr).cuda(), torch.as_tensor(col).cuda(), torch.as_tensor(val).cuda())
    out = g.spmm(torch.as_tensor(E).cuda()).cpu().numpy()
    ref = np.zeros((n, d), dtype=np.float64)
    rows = np.repeat(np.arange(n), deg)
    np.add.at(ref, rows, val[:, None].astype(np.float64) * E[col].astype(np.float64))
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() <= 2e-5 * scale
    assert (out[::97] == 0).all()


@pytest.mark.parametrize("R,K", [(1, 1), (7, 3), (8192, 1), (513, 130), (100000, 33)])
def test_col_reduce_cluster_path_shapes(R, K):
    import torch

    from librecommender_b200 import _lib

    rng = np.random.default_rng(R + K)
    X = rng.standard_normal((R, K)).astype(np.float32)
    Y = rng.standard_normal((R, K)).astype(np.float32)
    w = rng.standard_normal(R).astype(np.float32)
    Xd, Yd, wd = (torch.as_tensor(a).cuda() for a in (X, Y, w))
    out = torch.full((K,), 0.5, dtype=torch.float32, device="cuda")     # the kernel ADDS into out
    _lib.check(_lib.lib.b200_col_reduce(_lib.ptr(Xd), Xd.stride(0), R, K, _lib.ptr(wd), _lib.ptr(Yd), Yd.stride(0),
                                        _lib.ptr(out), _lib.current_stream()))
    ref = 0.5 + (X.astype(np.float64) * Y * w[:, None]).sum(axis=0)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-6, atol=2e-6 * np.sqrt(R))
    out2 = torch.zeros(K, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.b200_col_reduce(_lib.ptr(Xd), Xd.stride(0), R, K, None, None, 0, _lib.ptr(out2),
                                        _lib.current_stream()))
    out3 = torch.zeros(K, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.b200_col_reduce(_lib.ptr(Xd), Xd.stride(0), R, K, None, None, 0, _lib.ptr(out3),
                                        _lib.current_stream()))
    np.testing.assert_array_equal(out2.cpu().numpy(), out3.cpu().numpy())          # deterministic
    np.testing.assert_allclose(out2.cpu().numpy(), X.astype(np.float64).sum(axis=0), rtol=2e-6, atol=2e-6 * np.sqrt(R))


def test_seq_pool_forward_backward_edge_sequences():
    import torch

    from librecommender_b200 import _lib

    rng = np.random.default_rng(3)
    n_items, K, T, R = 50, 16, 7, 40
    E = rng.standard_normal((n_items + 1, K)).astype(np.float32)
    lens = rng.integers(0, T + 1, R).astype(np.int32)
    lens[0], lens[1] = 0, T
    seqs = np.full((R, T), n_items, dtype=np.int32)
    for r in range(R):
        seqs[r, :lens[r]] = rng.integers(0, n_items, lens[r])
    seqs[2, :] = n_items                  # a row whose positions are ALL the pad id although len says 3
    lens[2] = 3
    Ed, sd, ld = torch.as_tensor(E).cuda(), torch.as_tensor(seqs).cuda(), torch.as_tensor(lens).cuda()
    rows = torch.arange(R, dtype=torch.int64, device="cuda")
    out = torch.empty((R, K), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.b200_seq_pool(_lib.ptr(Ed), Ed.stride(0), K, n_items, _lib.ptr(sd), sd.stride(0), _lib.ptr(ld), T,
                                      _lib.ptr(rows), R, 0, 0, _lib.ptr(out), out.stride(0), _lib.current_stream()))
    Ez = E.copy()
    Ez[n_items] = 0
    ref = Ez[seqs].sum(axis=1) / np.where(lens > 0, np.sqrt(np.maximum(lens, 1)), np.inf)[:, None]
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    dout = rng.standard_normal((R, K)).astype(np.float32)
    g = torch.zeros((n_items + 1, K), dtype=torch.float32, device="cuda")
    dd = torch.as_tensor(dout).cuda()
    _lib.check(_lib.lib.b200_seq_pool_backward(_lib.ptr(dd), dd.stride(0), K, n_items, _lib.ptr(sd), sd.stride(0),
                                               _lib.ptr(ld), T, _lib.ptr(rows), R, _lib.ptr(g), g.stride(0),
                                               _lib.current_stream()))
    gref = np.zeros((n_items + 1, K), dtype=np.float64)
    for r in range(R):
        if lens[r] > 0:
            for t in range(T):
                if seqs[r, t] != n_items:
                    gref[seqs[r, t]] += dout[r] / np.sqrt(lens[r])
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-5, atol=1e-5)
    assert (g[n_items] == 0).all()


def test_topk_whole_row_and_single_column():
    import torch

    from librecommender_b200.engine import EmbedScorer

    rng = np.random.default_rng(5)
    n_users, N, d = 20, 37, 8
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    sc = EmbedScorer(U, I, N, {}, n_users=n_users)
    uid = torch.arange(n_users, device="cuda")
    ids, scores = sc.recommend_exact(uid, N, True, True)            # n_rec == n_items: a full sort of every row
    full = (U[:n_users].astype(np.float64) @ I[:N].astype(np.float64).T)
    np.testing.assert_array_equal(np.sort(ids.cpu().numpy(), axis=1), np.tile(np.arange(N), (n_users, 1)))
    assert (np.diff(scores.cpu().numpy(), axis=1) <= 0).all()
    ids1 = sc.recommend_exact(uid, 1, True, False).cpu().numpy()
    np.testing.assert_array_equal(ids1[:, 0], full.argmax(axis=1))
    with pytest.raises(ValueError, match="exceeds num of items"):
        sc.recommend_exact(uid, N + 1, True, False)
