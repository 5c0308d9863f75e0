"""One-GPU probe of the sharded LightGCN SpMM: time the local row block of a simulated world of G
ranks (no communication) against the full-graph SpMM, to separate kernel / layout effects from
exchange effects in the multi-GPU efficiency.  python tools/lightgcn_shard_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librecommender_b200.consumed import ConsumedCSR  # noqa: E402
from librecommender_b200.lightgcn import SpmmGraph, build_laplacian_csr  # noqa: E402
from librecommender_b200.parallel import LightGCNShardPlan, split_column_blocks  # noqa: E402


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    n_users, n_items, d = 2_000_000, 200_000, 64
    deg = torch.clamp(torch.poisson(torch.full((n_users,), 50.0, device=dev), generator=g), 1, 2000).long()
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(deg, 0)
    u = torch.rand(int(indptr[-1]), device=dev, generator=g)
    idx = (torch.exp(u * np.log(n_items)) - 1).clamp(0, n_items - 1).to(torch.int32)
    ip, col, val = build_laplacian_csr(ConsumedCSR.from_device_tensors(indptr, idx), n_users, n_items, dev)
    full = SpmmGraph(ip, col, val)
    E = torch.randn(n_users + n_items, d, device=dev, generator=g) * 0.1
    out = torch.empty_like(E)
    t_full = timeit(lambda: full.spmm(E, out=out))
    print(json.dumps({"what": "full graph, one layer", "ms": t_full, "nnz": full.nnz, "long_rows": full.n_long}), flush=True)
    for G in (2, 4, 8):
        plan = LightGCNShardPlan(n_users, n_items, G)
        lptr, lcol, lval = plan.shard_csr(ip, col, val, 0)
        lg = SpmmGraph(lptr, lcol, lval)
        Eg = torch.randn(G * plan.slab, d, device=dev, generator=g) * 0.1
        o = torch.empty((plan.slab, d), device=dev)
        t_loc = timeit(lambda: lg.spmm(Eg, out=o))
        blocks = [SpmmGraph(*b) for b in split_column_blocks(lptr, lcol, lval, plan.slab, G)]
        acc = torch.zeros((plan.slab, d), device=dev)
        Eb = [Eg[i * plan.slab:(i + 1) * plan.slab].contiguous() for i in range(G)]

        def run_blocks():
            acc.zero_()
            for i in range(G):
                if blocks[i].nnz:
                    blocks[i].spmm(Eb[i], out=None, acc=acc, acc_init=False, final_div=0.0)

        t_blk = timeit(run_blocks)
        print(json.dumps({"what": f"rank 0 of a world of {G}", "ms_local_block": t_loc, "ms_column_blocks": t_blk,
                          "nnz_local": lg.nnz, "ideal_ms": t_full / G, "kernel_efficiency": t_full / G / t_loc,
                          "long_rows": lg.n_long}), flush=True)


if __name__ == "__main__":
    main()
