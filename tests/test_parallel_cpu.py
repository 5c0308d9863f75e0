"""CPU, world_size 2, gloo: the user-sharding host logic of the multi-GPU recommend path
(sharding bounds, order-preserving gather).  The per-rank compute is a stub — the CUDA scorer
cannot run here — so this covers exactly the code that is new at N > 1."""
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


def _worker(rank, world, port, n_users, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import recommend_sharded, shard_users

    users = np.arange(100, 100 + n_users)
    seen = []

    def fake_recommend(local, n_rec):
        seen.append(local.copy())
        return np.stack([local * 10 + k for k in range(n_rec)], axis=1)

    out = recommend_sharded(fake_recommend, users, 3)
    mine = shard_users(users, world, rank)
    q.put((rank, out, seen[0] if seen else np.zeros(0), mine))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_users", [7, 8, 1])
def test_sharded_recommend_gloo_world2(n_users):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_users, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    users = np.arange(100, 100 + n_users)
    expect = np.stack([users * 10 + k for k in range(3)], axis=1)
    covered = np.concatenate([r[3] for r in sorted(res, key=lambda x: x[0])])
    np.testing.assert_array_equal(covered, users)                  # shards tile the batch, in order
    for rank, out, seen, mine in res:
        np.testing.assert_array_equal(out, expect)                  # every rank holds the full result
        np.testing.assert_array_equal(seen, mine)


def test_shard_bounds_balanced():
    from librecommender_b200.parallel import shard_bounds

    for n in (0, 1, 5, 8, 8191, 8192):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker_mean(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import sharded_mean_row

    full = torch.from_numpy(np.random.default_rng(0).standard_normal((101, 6)).astype(np.float32))
    local = full[rank::world].contiguous()
    q.put((rank, sharded_mean_row(local, local.shape[0], 101).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_oov_mean_row_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_mean, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.random.default_rng(0).standard_normal((101, 6)).astype(np.float32)
    for _, m in res:
        np.testing.assert_allclose(m, full.mean(axis=0), rtol=1e-6, atol=1e-7)


def test_per_rank_sampler_streams_differ():
    from librecommender_b200.sampling import rank_stream_seed

    seeds = [rank_stream_seed(462, r) for r in range(8)]
    assert seeds[0] == 462 and len(set(seeds)) == 8 and all(0 <= s < 2 ** 63 for s in seeds)


def test_restrict_consumed_to_shard_keeps_order_and_translates_ids():
    from librecommender_b200.parallel import restrict_consumed_to_shard

    indptr = np.array([0, 4, 4, 7, 9], dtype=np.int64)                  # user 1 has no history
    idx = np.array([50, 3, 120, 51, 7, 199, 100, 0, 99], dtype=np.int32)
    lptr, lidx = restrict_consumed_to_shard(indptr, idx, 50, 150)
    np.testing.assert_array_equal(lptr, [0, 3, 3, 4, 5])
    np.testing.assert_array_equal(lidx, [0, 70, 1, 50, 49])             # arrival order kept, ids minus 50
    lptr, lidx = restrict_consumed_to_shard(indptr, idx, 0, 50)
    np.testing.assert_array_equal(lptr, [0, 1, 1, 2, 3])
    np.testing.assert_array_equal(lidx, [3, 7, 0])
    lptr, lidx = restrict_consumed_to_shard(np.array([0, 0], dtype=np.int64), np.zeros(0, np.int32), 0, 10)
    np.testing.assert_array_equal(lptr, [0, 0])
    assert lidx.size == 0


def _worker_item_sharded(rank, world, port, q):
    import os

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from librecommender_b200.parallel import recommend_item_sharded, shard_bounds

    rng = np.random.default_rng(0)
    n_users, N, d, K = 40, 300, 8, 7
    U = rng.standard_normal((n_users, d)).astype(np.float32)
    I = rng.standard_normal((N, d)).astype(np.float32)
    lo, hi = shard_bounds(N, world, rank)

    class Shard:                                  # stand-in for ItemShardScorer (numpy scores of this rank's items)
        def local_topk(self, users, n_rec):
            sc = U[users.numpy()] @ I[lo:hi].T
            order = np.lexsort((np.broadcast_to(np.arange(hi - lo), sc.shape), -sc), axis=1)[:, :n_rec]
            return torch.as_tensor(order + lo), torch.as_tensor(np.take_along_axis(sc, order, axis=1))

    def merge(ids, scores, n_rec):                # stand-in for merge_topk_shards (the CUDA kernel)
        G, B, K_ = ids.shape
        ci = ids.permute(1, 0, 2).reshape(B, G * K_).numpy()
        cs = scores.permute(1, 0, 2).reshape(B, G * K_).numpy()
        order = np.lexsort((ci, -cs), axis=1)[:, :n_rec]
        return torch.as_tensor(np.take_along_axis(ci, order, axis=1)), torch.as_tensor(np.take_along_axis(cs, order, axis=1))

    users = torch.as_tensor(np.arange(0, n_users, 3))
    ids, sc = recommend_item_sharded(Shard(), users, K, merge=merge)
    full = U[users.numpy()] @ I.T
    ref = np.lexsort((np.broadcast_to(np.arange(N), full.shape), -full), axis=1)[:, :K]
    q.put((rank, bool((ids.numpy() == ref).all()), bool(np.allclose(sc.numpy(), np.take_along_axis(full, ref, axis=1)))))
    dist.barrier()
    dist.destroy_process_group()


def test_item_sharded_recommend_gloo_world2():
    """all-gather plumbing of recommend_item_sharded with numpy stand-ins for the shard scorer and the merge kernel."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_item_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ids_ok, sc_ok in res:
        assert ids_ok and sc_ok, rank
