"""2-GPU check of the sharded LightGCN propagation (NCCL all-gather per layer + local CUDA SpMM):
bit-for-bit equal to the single-GPU propagation.  Skipped on boxes with one GPU (the CPU gloo test
tests/test_lightgcn_sharded_cpu.py covers the sharding logic there)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_propagation_two_gpus_bit_exact():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LG_USERS="300000", LG_ITEMS="40000")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(root, "tools", "lightgcn_sharded_check.py")],
                       capture_output=True, text=True, env=env, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["max_abs_err_vs_single_gpu"] == 0.0
    assert max(d["nnz_per_rank"]) <= 1.1 * min(d["nnz_per_rank"])
    assert d["two_phase_max_abs_err"] <= 1e-5 * max(d["ref_scale"], 1.0)      # different summation order: not bit-exact


def test_row_sharded_table_two_gpus():
    """Index / row all-to-all over NCCL around the CUDA gather and scatter-add kernels."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ST_ROWS="300000", ST_IDS="65536")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29579",
                        os.path.join(root, "tools", "sharded_table_check.py")],
                       capture_output=True, text=True, env=env, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["lookup_exact"] and d["scatter_add_max_err"] < 1e-3


def test_gather_and_scatter_rows_single_gpu():
    """The two local kernels of the sharded table against torch indexing (world size 1 path)."""
    import torch

    from librecommender_b200.parallel import RowShardedTable

    g = torch.Generator(device="cuda").manual_seed(3)
    full = torch.randn(5000, 48, device="cuda", generator=g)
    t = RowShardedTable(full.clone(), 5000)
    ids = torch.randint(0, 5000, (20000,), device="cuda", generator=g)
    assert torch.equal(t.lookup(ids), full[ids])
    grads = torch.randn(20000, 48, device="cuda", generator=g)
    t.scatter_add(ids, grads)
    expect = full.clone().index_add_(0, ids, grads)
    assert float((t.local - expect).abs().max()) < 1e-4
    assert t.lookup(ids[:0]).shape == (0, 48)


@pytest.mark.parametrize("G,d", [(1, 16), (2, 16), (4, 64), (3, 20), (8, 32)])
def test_peer_gather_and_scatter_kernels_virtual_shards(G, d):
    """b200_peer_gather_rows / b200_peer_scatter_add_rows with G shards that all live on THIS GPU: the
    row -> (shard id % G, slot id / G) addressing, the vectorised and the generic paths, the float
    atomics — exactly what runs over NVLink peer pointers when the shards belong to G GPUs."""
    import ctypes

    import torch

    from librecommender_b200 import _lib

    n_rows = 100_003
    g = torch.Generator(device="cuda").manual_seed(G * 100 + d)
    full = torch.randn(n_rows, d, device="cuda", generator=g)
    rows_loc = -(-n_rows // G)
    shards = []
    for r in range(G):
        s = torch.zeros(rows_loc, d, device="cuda")
        part = full[r::G]
        s[: part.shape[0]] = part
        shards.append(s)
    ptrs = (ctypes.c_void_p * G)(*[s.data_ptr() for s in shards])
    ids = torch.randint(0, n_rows, (300_001,), device="cuda", generator=g)
    out = torch.empty(ids.numel(), d, device="cuda")
    _lib.check(_lib.lib.b200_peer_gather_rows(ptrs, G, d, d, _lib.ptr(ids), ids.numel(), _lib.ptr(out), d,
                                              _lib.current_stream()))
    assert torch.equal(out, full[ids])
    grads = torch.randn(ids.numel(), d, device="cuda", generator=g)
    _lib.check(_lib.lib.b200_peer_scatter_add_rows(ptrs, G, d, d, _lib.ptr(ids), ids.numel(), _lib.ptr(grads), d,
                                                   _lib.current_stream()))
    expect = full.clone().index_add_(0, ids, grads)
    got = torch.empty_like(full)
    for r in range(G):
        got[r::G] = shards[r][: full[r::G].shape[0]]
    assert float((got - expect).abs().max()) < 2e-4


@pytest.mark.parametrize("G", [2, 3])
def test_item_sharded_recommend_virtual_shards_equals_single_gpu(G):
    """SURVEY 8e row 1, item-sharded variant: G shards of the catalogue scored one after the other in ONE process
    (the all-gather is a stack), merged with b200_topk_rows: ids and exact scores equal the single-scorer result."""
    import numpy as np
    import torch

    from librecommender_b200.consumed import ConsumedCSR
    from librecommender_b200.engine import EmbedScorer
    from librecommender_b200.parallel import ItemShardScorer, merge_topk_shards, shard_bounds

    rng = np.random.default_rng(G)
    n_users, N, d, K = 500, 30001, 64, 100
    U = rng.standard_normal((n_users + 1, d)).astype(np.float32)
    I = rng.standard_normal((N + 1, d)).astype(np.float32)
    I[7000] = I[12] ; I[20000] = I[12]; I[29999] = I[12]          # equal scores across shards: the id order must survive
    consumed = {u: rng.choice(N, size=int(rng.integers(0, 60)), replace=False).tolist() for u in range(n_users)}
    consumed = {u: c for u, c in consumed.items() if c}
    csr = ConsumedCSR.from_dict(consumed, n_users)
    full = EmbedScorer(U, I, N, csr, n_users=n_users)
    users = torch.as_tensor(rng.integers(0, n_users + 1, 300)).cuda()
    ref_ids, ref_sc = full.recommend_device(users, K, True, True)
    ids_l, sc_l = [], []
    for g in range(G):
        lo, hi = shard_bounds(N, G, g)
        sh = ItemShardScorer(U, torch.as_tensor(I[lo:hi]), lo, csr.indptr, csr.idx, n_users=n_users)
        a, b = sh.local_topk(users, K)
        ids_l.append(a)
        sc_l.append(b)
    ids, sc = merge_topk_shards(torch.stack(ids_l), torch.stack(sc_l), K)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids.cpu().numpy())
    np.testing.assert_array_equal(sc.cpu().numpy(), ref_sc.cpu().numpy())
    tiny = ItemShardScorer(U, torch.as_tensor(I[:120]), 0, csr.indptr, csr.idx, n_users=n_users)
    with pytest.raises(ValueError, match="too small"):
        tiny.local_topk(users, K)
