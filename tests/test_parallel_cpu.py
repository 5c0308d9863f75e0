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
