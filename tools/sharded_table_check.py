"""2+-GPU check + timing of the row-sharded embedding table (index all-to-all -> CUDA row gather ->
row all-to-all over NCCL).  Launch with torchrun, one rank per GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from librecommender_b200.parallel import RowShardedTable

    n_rows, d, n_ids = int(os.environ.get("ST_ROWS", 4_000_000)), 64, int(os.environ.get("ST_IDS", 1 << 20))
    g = torch.Generator(device=dev).manual_seed(1)
    full = torch.randn(n_rows, d, device=dev, generator=g)                  # same on every rank (same seed)
    t = RowShardedTable(RowShardedTable.shard(full, world, rank).clone(), n_rows)
    gi = torch.Generator(device=dev).manual_seed(100 + rank)
    ids = torch.randint(0, n_rows, (n_ids,), device=dev, generator=gi)
    rows = t.lookup(ids)
    ok_lookup = bool(torch.equal(rows, full[ids]))
    grads = torch.randn(n_ids, d, device=dev, generator=gi)
    t.scatter_add(ids, grads)
    # every rank's (ids, grads) must have reached the owners: rebuild the expectation with all of them
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_g = [torch.empty_like(grads) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_g, grads)
    else:
        all_ids, all_g = [ids], [grads]
    expect = full.clone()
    for i_, g_ in zip(all_ids, all_g):
        expect.index_add_(0, i_, g_)
    err = float((t.local - expect[rank::world]).abs().max())

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    ms = timed(lambda: t.lookup(ids))
    if rank == 0:
        print(json.dumps({"check": "row-sharded table", "world": world, "rows": n_rows, "d": d, "ids_per_rank": n_ids,
                          "lookup_exact": ok_lookup, "scatter_add_max_err": err, "lookup_ms": ms,
                          "rows_per_s_all_ranks": world * n_ids / (ms * 1e-3),
                          "gb_per_s_all_ranks": world * n_ids * d * 4 / (ms * 1e-3) / 1e9}), flush=True)
    assert ok_lookup and err < 1e-3, (ok_lookup, err)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
