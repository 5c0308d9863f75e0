"""Multi-GPU layout of the recommend path: independent users are sharded
across ranks, the item table (+ bf16 catalog) is replicated, there is NO data-path collective.
Only the optional result gather touches the process group.  One process per GPU
(``torchrun``), ``torch.distributed`` for the plumbing (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of n users: the first n % world ranks take one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_users(user_ids, world: int, rank: int):
    user_ids = np.asarray(user_ids)
    lo, hi = shard_bounds(len(user_ids), world, rank)
    return user_ids[lo:hi]


def recommend_sharded(recommend_fn, user_ids, n_rec, group=None, gather=True):
    """Run ``recommend_fn(local_user_ids, n_rec) -> int64[b, n_rec]`` on this rank's shard and
    (optionally) all-gather the ``[B, n_rec]`` result in the original user order on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local = shard_users(user_ids, world, rank)
    out = recommend_fn(local, n_rec) if len(local) else np.zeros((0, n_rec), dtype=np.int64)
    out = np.asarray(out, dtype=np.int64)
    if world == 1 or not gather:
        return out
    sizes = [shard_bounds(len(user_ids), world, r) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.full((max_rows, n_rec), -1, dtype=torch.int64, device=dev)
    buf[: len(out)] = torch.from_numpy(out).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    rows = [p[: hi - lo].cpu().numpy() for p, (lo, hi) in zip(parts, sizes)]
    return np.concatenate(rows, axis=0)
