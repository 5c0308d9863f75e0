"""CPU, world_size 2, gloo: routing logic of the row-sharded embedding table (owner = row % G, index
all-to-all, row all-to-all, original order restored; transposed path for gradients).  The two local
operations (row gather / scatter-add) are torch stand-ins here, the CUDA kernels on GPUs."""
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


def _worker(rank, world, port, n_rows, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import RowShardedTable

    full = torch.from_numpy(np.random.default_rng(0).standard_normal((n_rows, d)).astype(np.float32))
    local = RowShardedTable.shard(full, world, rank).clone()

    def scatter(tab, slots, rows):
        tab.index_add_(0, slots, rows)

    t = RowShardedTable(local, n_rows, gather_fn=lambda tab, slots: tab[slots], scatter_fn=scatter)
    rng = np.random.default_rng(100 + rank)
    ids = torch.from_numpy(rng.integers(0, n_rows, 37 + 5 * rank))          # different batch sizes per rank
    rows = t.lookup(ids)
    grads = torch.from_numpy(rng.standard_normal((len(ids), d)).astype(np.float32))
    t.scatter_add(ids, grads)
    empty = t.lookup(torch.zeros(0, dtype=torch.int64))
    q.put((rank, ids.numpy(), rows.numpy(), grads.numpy(), t.local.numpy(), tuple(empty.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [101, 64])
def test_row_sharded_lookup_and_scatter_gloo_world2(n_rows):
    world, port, d = 2, _free_port(), 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.random.default_rng(0).standard_normal((n_rows, d)).astype(np.float32)
    expect = full.copy()
    for rank, ids, rows, grads, local, eshape in res:
        np.testing.assert_array_equal(rows, full[ids])                      # lookups return the caller's rows
        np.add.at(expect, ids, grads)
        assert eshape == (0, d)
    for rank, _, _, _, local, _ in res:
        np.testing.assert_allclose(local, expect[rank::world], rtol=1e-6, atol=1e-6)   # gradients reached the owners
