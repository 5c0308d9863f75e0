"""Multi-GPU LightGCN propagation check + timing.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/lightgcn_sharded_check.py
Every rank builds the same synthetic bipartite graph, rank r keeps its row block of L and its
block of E; one NCCL all-gather per layer; the result is compared with the single-GPU propagation
(bit-for-bit: same CSR order inside a row, same fma chain) and timed with CUDA events (max over
ranks)."""
import json
import os
import sys

import numpy as np
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
    from librecommender_b200.consumed import ConsumedCSR
    from librecommender_b200.lightgcn import SpmmGraph, build_laplacian_csr, propagate
    from librecommender_b200.parallel import (LightGCNShardPlan, gather_embeddings, propagate_sharded,
                                              sharded_spmm_fn)

    n_users = int(os.environ.get("LG_USERS", 2_000_000))
    n_items = int(os.environ.get("LG_ITEMS", 200_000))
    d, n_layers = 64, 3
    g = torch.Generator(device=dev).manual_seed(5)
    deg = torch.clamp(torch.poisson(torch.full((n_users,), 50.0, device=dev), generator=g), 1, 2000).long()
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(deg, 0)
    nnz = int(indptr[-1])
    # Zipf(1.0)-like item popularity through an exponential transform of uniforms
    u = torch.rand(nnz, device=dev, generator=g)
    idx = (torch.exp(u * np.log(n_items)) - 1).clamp(0, n_items - 1).to(torch.int32)
    csr = ConsumedCSR.from_device_tensors(indptr, idx)
    ip, col, val = build_laplacian_csr(csr, n_users, n_items, dev)
    E0 = torch.randn(n_users + n_items, d, device=dev, generator=g) * 0.1

    plan = LightGCNShardPlan(n_users, n_items, world)
    lg = SpmmGraph(*plan.shard_csr(ip, col, val, rank))
    E0_loc = plan.scatter_rows(E0, rank)

    spmm_local = sharded_spmm_fn(lg, plan.slab)

    def run_sharded():
        return propagate_sharded(plan, spmm_local, E0_loc, n_layers)

    out_loc = run_sharded()
    ue, ie = gather_embeddings(plan, out_loc)
    full_graph = SpmmGraph(ip, col, val)
    ref = propagate(full_graph, E0, n_layers)
    got = torch.cat([ue, ie])
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())

    def timed(fn, iters=5):
        for _ in range(2):
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
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_sharded = timed(run_sharded)
    # two products per layer (own block while the slabs are pulled over NVLink peer memory, the rest in one)
    err_two, ms_two = None, None
    if world > 1:
        from librecommender_b200.parallel import (PeerPullExchange, acc_spmm_fn, propagate_sharded_two_phase,
                                                  split_local_remote)

        lptr, lcol, lval = plan.shard_csr(ip, col, val, rank)
        own, rest = split_local_remote(lptr, lcol, lval, plan.slab, rank)
        g_own, g_rest = SpmmGraph(*own), SpmmGraph(*rest)
        ex = PeerPullExchange(world, rank)

        def run_two():
            return propagate_sharded_two_phase(plan, acc_spmm_fn(g_own), acc_spmm_fn(g_rest), E0_loc, n_layers, ex, rank)

        ue2, ie2 = gather_embeddings(plan, run_two())
        err_two = float((torch.cat([ue2, ie2]) - ref).abs().max())
        ms_two = timed(run_two)
    ms_single = timed(lambda: propagate(full_graph, E0, n_layers))
    nnz_l = torch.tensor([lg.nnz], device=dev)
    nnz_all = [torch.zeros_like(nnz_l) for _ in range(world)]
    if world > 1:
        dist.all_gather(nnz_all, nnz_l)
    else:
        nnz_all = [nnz_l]
    if rank == 0:
        print(json.dumps({"check": "lightgcn sharded propagation", "world": world, "n_users": n_users,
                          "n_items": n_items, "nnz": int(col.numel()), "d": d, "layers": n_layers,
                          "max_abs_err_vs_single_gpu": err, "ref_scale": scale,
                          "nnz_per_rank": [int(x) for x in nnz_all],
                          "ms_sharded_3_layers": ms_sharded, "ms_single_gpu_3_layers": ms_single,
                          "speedup": ms_single / ms_sharded, "two_phase_max_abs_err": err_two,
                          "ms_two_phase_3_layers": ms_two,
                          "allgather_bytes_per_layer_per_rank": plan.world * plan.slab * d * 4}), flush=True)
    assert err <= 1e-6 * max(scale, 1.0), (err, scale)
    assert err_two is None or err_two <= 1e-5 * max(scale, 1.0), (err_two, scale)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
