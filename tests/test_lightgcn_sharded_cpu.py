"""CPU, world_size 2, gloo: the sharding logic of multi-GPU LightGCN propagation (row partition,
column remap into the gathered layout, one all-gather per layer, un-permute).  The local multiply is
a torch.sparse stand-in — the CUDA SpMM cannot run here — so this covers exactly the code that is
new at N > 1; the result must equal the single-process scipy oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(seed, n_users, n_items):
    from oracle import lightgcn as ol

    rng = np.random.default_rng(seed)
    consumed = {u: rng.choice(n_items, size=int(rng.integers(1, min(12, n_items + 1))), replace=False).tolist() for u in range(n_users)}
    L = ol.build_laplacian(n_users, n_items, consumed).tocsr()
    L.sort_indices()
    return consumed, L


def _worker(rank, world, port, n_users, n_items, d, n_layers, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import LightGCNShardPlan, gather_embeddings, propagate_sharded

    _, L = _graph(0, n_users, n_items)
    E0 = torch.from_numpy(np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32))
    plan = LightGCNShardPlan(n_users, n_items, world)
    lptr, lcol, lval = plan.shard_csr(torch.from_numpy(L.indptr.astype(np.int64)),
                                      torch.from_numpy(L.indices.astype(np.int32)),
                                      torch.from_numpy(L.data.astype(np.float32)), rank)
    Lloc = torch.sparse_csr_tensor(lptr, lcol.to(torch.int64), lval, size=(plan.slab, world * plan.slab))
    def spmm_local(full, acc, final_div):      # stand-in for the CUDA SpMM with its fused epilogue
        out = Lloc @ full
        acc += out
        if final_div > 0:
            acc /= final_div
        return out

    out_local = propagate_sharded(plan, spmm_local, plan.scatter_rows(E0, rank), n_layers)
    ue, ie = gather_embeddings(plan, out_local)
    q.put((rank, ue.numpy(), ie.numpy(), int(lval.numel())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_users,n_items", [(37, 23), (40, 10), (5, 64)])
def test_sharded_propagation_matches_oracle_gloo_world2(n_users, n_items):
    from oracle import lightgcn as ol

    world, port, d, n_layers = 2, _free_port(), 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_users, n_items, d, n_layers, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, L = _graph(0, n_users, n_items)
    E0 = np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32)
    ref = np.concatenate(ol.propagate(L, E0[:n_users], E0[n_users:], n_layers))
    assert sum(r[3] for r in res) == L.nnz                          # the row blocks tile the matrix
    for rank, ue, ie, _ in res:
        np.testing.assert_allclose(np.concatenate([ue, ie]), ref, rtol=1e-5, atol=1e-6)


def test_plan_layout_roundtrip():
    from librecommender_b200.parallel import LightGCNShardPlan

    for nu, ni, w in ((10, 3, 4), (7, 7, 2), (1, 9, 8)):
        plan = LightGCNShardPlan(nu, ni, w)
        pos = plan.position(torch.arange(nu + ni))
        assert len(set(pos.tolist())) == nu + ni and int(pos.max()) < w * plan.slab
        seen = []
        for r in range(w):
            nodes, slots = plan.local_nodes(r)
            np.testing.assert_array_equal(plan.position(nodes).numpy(), r * plan.slab + slots.numpy())
            seen += nodes.tolist()
        assert sorted(seen) == list(range(nu + ni))


@pytest.mark.parametrize("world", [1, 3, 4, 8])
def test_plan_single_process_simulation_any_world(world):
    """The same partition / remap logic for world sizes the gloo test does not spawn: all ranks are
    simulated in one process (the all-gather is a concatenation of the per-rank blocks)."""
    from librecommender_b200.parallel import LightGCNShardPlan
    from oracle import lightgcn as ol

    n_users, n_items, d, n_layers = 29, 17, 5, 3
    _, L = _graph(4, n_users, n_items)
    E0 = torch.from_numpy(np.random.default_rng(2).standard_normal((n_users + n_items, d)).astype(np.float32))
    plan = LightGCNShardPlan(n_users, n_items, world)
    ip = torch.from_numpy(L.indptr.astype(np.int64))
    ci = torch.from_numpy(L.indices.astype(np.int32))
    va = torch.from_numpy(L.data.astype(np.float32))
    locals_, nnz = [], 0
    for r in range(world):
        lptr, lcol, lval = plan.shard_csr(ip, ci, va, r)
        nnz += int(lval.numel())
        locals_.append(torch.sparse_csr_tensor(lptr, lcol.to(torch.int64), lval, size=(plan.slab, world * plan.slab)))
    assert nnz == L.nnz
    cur = [plan.scatter_rows(E0, r) for r in range(world)]
    acc = [c.clone() for c in cur]
    for _ in range(n_layers):
        full = torch.cat(cur, dim=0)                         # what all_gather_into_tensor produces
        cur = [locals_[r] @ full for r in range(world)]
        acc = [a + c for a, c in zip(acc, cur)]
    out = plan.unpermute(torch.cat(acc, dim=0) / (n_layers + 1)).numpy()
    ref = np.concatenate(ol.propagate(L, E0[:n_users].numpy(), E0[n_users:].numpy(), n_layers))
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


# ---- overlapped variant: column blocks per source rank + ring exchange (gloo) ------------------
def _worker_overlap(rank, world, port, n_users, n_items, d, n_layers, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import (LightGCNShardPlan, RingExchange, gather_embeddings,
                                              propagate_sharded_overlap, split_column_blocks)

    _, L = _graph(0, n_users, n_items)
    E0 = torch.from_numpy(np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32))
    plan = LightGCNShardPlan(n_users, n_items, world)
    lptr, lcol, lval = plan.shard_csr(torch.from_numpy(L.indptr.astype(np.int64)),
                                      torch.from_numpy(L.indices.astype(np.int32)),
                                      torch.from_numpy(L.data.astype(np.float32)), rank)
    blocks = split_column_blocks(lptr, lcol, lval, plan.slab, world)
    mats = [torch.sparse_csr_tensor(p, c.to(torch.int64), v, size=(plan.slab, plan.slab)) for p, c, v in blocks]

    def block_spmm(g, E_block, acc):           # stand-in for SpmmGraph.spmm(acc=acc, acc_init=False)
        acc += mats[g] @ E_block

    ex = RingExchange(world, rank)
    out_local = propagate_sharded_overlap(plan, block_spmm, plan.scatter_rows(E0, rank), n_layers, ex, rank)
    ue, ie = gather_embeddings(plan, out_local)
    q.put((rank, ue.numpy(), ie.numpy(), sum(int(v.numel()) for _, _, v in blocks)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_overlapped_propagation_matches_oracle_gloo(world):
    from oracle import lightgcn as ol

    n_users, n_items, port, d, n_layers = 37, 23, _free_port(), 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, n_users, n_items, d, n_layers, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, L = _graph(0, n_users, n_items)
    E0 = np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32)
    ref = np.concatenate(ol.propagate(L, E0[:n_users], E0[n_users:], n_layers))
    assert sum(r[3] for r in res) == L.nnz                          # the column blocks tile the matrix
    for rank, ue, ie, _ in res:
        np.testing.assert_allclose(np.concatenate([ue, ie]), ref, rtol=1e-5, atol=1e-6)


def test_split_column_blocks_preserves_rows_and_order():
    from librecommender_b200.parallel import LightGCNShardPlan, split_column_blocks

    n_users, n_items, world = 29, 17, 4
    _, L = _graph(4, n_users, n_items)
    plan = LightGCNShardPlan(n_users, n_items, world)
    ip = torch.from_numpy(L.indptr.astype(np.int64))
    ci = torch.from_numpy(L.indices.astype(np.int32))
    va = torch.from_numpy(L.data.astype(np.float32))
    for r in range(world):
        lptr, lcol, lval = plan.shard_csr(ip, ci, va, r)
        dense = torch.sparse_csr_tensor(lptr, lcol.to(torch.int64), lval, size=(plan.slab, world * plan.slab)).to_dense()
        blocks = split_column_blocks(lptr, lcol, lval, plan.slab, world)
        for g, (p, c, v) in enumerate(blocks):
            assert c.numel() == 0 or int(c.max()) < plan.slab
            blk = torch.sparse_csr_tensor(p, c.to(torch.int64), v, size=(plan.slab, plan.slab)).to_dense()
            np.testing.assert_array_equal(blk.numpy(), dense[:, g * plan.slab:(g + 1) * plan.slab].numpy())


# ---- two-phase variant: own block while the slabs travel, every other block in ONE product (gloo) -------------
def _worker_two_phase(rank, world, port, n_users, n_items, d, n_layers, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import (LightGCNShardPlan, RingExchange, gather_embeddings,
                                              propagate_sharded_two_phase, split_local_remote)

    _, L = _graph(0, n_users, n_items)
    E0 = torch.from_numpy(np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32))
    plan = LightGCNShardPlan(n_users, n_items, world)
    lptr, lcol, lval = plan.shard_csr(torch.from_numpy(L.indptr.astype(np.int64)),
                                      torch.from_numpy(L.indices.astype(np.int32)),
                                      torch.from_numpy(L.data.astype(np.float32)), rank)
    (p0, c0, v0), (p1, c1, v1) = split_local_remote(lptr, lcol, lval, plan.slab, rank)
    own = torch.sparse_csr_tensor(p0, c0.to(torch.int64), v0, size=(plan.slab, plan.slab))
    rest = torch.sparse_csr_tensor(p1, c1.to(torch.int64), v1, size=(plan.slab, world * plan.slab))

    def local_spmm(E, acc):
        acc += own @ E

    def remote_spmm(G, acc):
        acc += rest @ G

    ex = RingExchange(world, rank)
    out_local = propagate_sharded_two_phase(plan, local_spmm, remote_spmm, plan.scatter_rows(E0, rank), n_layers, ex,
                                            rank)
    ue, ie = gather_embeddings(plan, out_local)
    q.put((rank, ue.numpy(), ie.numpy(), int(v0.numel()) + int(v1.numel()),
           bool(c1.numel() == 0 or ((c1.to(torch.int64) // plan.slab) != rank).all())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_phase_propagation_matches_oracle_gloo(world):
    from oracle import lightgcn as ol

    n_users, n_items, port, d, n_layers = 37, 23, _free_port(), 8, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_two_phase, args=(r, world, port, n_users, n_items, d, n_layers, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, L = _graph(0, n_users, n_items)
    E0 = np.random.default_rng(1).standard_normal((n_users + n_items, d)).astype(np.float32)
    ref = np.concatenate(ol.propagate(L, E0[:n_users], E0[n_users:], n_layers))
    assert sum(r[3] for r in res) == L.nnz                          # own + rest tile the matrix
    for rank, ue, ie, _, rest_has_no_own_columns in res:
        assert rest_has_no_own_columns
        np.testing.assert_allclose(np.concatenate([ue, ie]), ref, rtol=1e-5, atol=1e-6)
