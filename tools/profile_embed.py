"""Small driver for ncu captures of the fused embed scorer (same kernels as bench.py, smaller
user table so that set-up is quick).  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 \
        -o gpurun_out/prof_sweep python tools/profile_embed.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    sys.argv = [sys.argv[0]] + ["--users", os.environ.get("PROF_USERS", "400000")] + sys.argv[1:]
    args = bench.parse()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    U, I = bench.make_tables(args, dev)
    indptr, idx = bench.make_consumed_csr(args, dev)
    from librecommender_b200 import _lib
    from librecommender_b200.consumed import ConsumedCSR
    from librecommender_b200.engine import EmbedScorer

    if args.epi_warps or args.pre_coef:
        _lib.check(_lib.lib.b200_recommend_embed_tune(args.epi_warps, args.pre_coef))

    sc = EmbedScorer(U, I, args.items, ConsumedCSR.from_device_tensors(indptr, idx), n_users=args.users, device=dev)
    batches = [torch.from_numpy(b).to(dev) for b in bench.make_batches(args, 0, args.steps)]
    for b in batches:
        sc.recommend_device(b, args.topk, True, False, args.path)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
