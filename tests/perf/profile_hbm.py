"""Driver for ncu captures of the HBM-bound kernels (north_star: "committed ncu captures showing
achieved HBM GB/s (gather, SpMM)").  One SpMM layer on a Zipf bipartite graph, the DeepFM-shaped
feature gather (forward) and its gradient scatter (backward).
    ncu --set full --clock-control none -k regex:"spmm|feat_forward|feat_backward" -s 6 -c 6 \
        -o gpurun_out/prof_hbm python tests/perf/profile_hbm.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from librecommender_b200.lightgcn import SpmmGraph  # noqa: E402
from librecommender_b200.training import FMTrainer  # noqa: E402
from oracle import tf_models as tm  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(5)
n_users, n_items, d = 1_000_000, 100_000, 64
counts = torch.poisson(torch.full((n_users,), 50.0, device=dev), generator=g).clamp_(min=1, max=2000).long()
w = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float64)
cdf = (torch.cumsum(w, 0) / w.sum()).float()
owner = torch.repeat_interleave(torch.arange(n_users, device=dev), counts)
item = torch.searchsorted(cdf, torch.rand(owner.numel(), generator=g, device=dev)).clamp_(max=n_items - 1)
n = n_users + n_items
shift = (n - 1).bit_length()
und = torch.unique((owner << shift) | (item + n_users))
r, c = und >> shift, und & ((1 << shift) - 1)
key = torch.sort(torch.cat([und, (c << shift) | r])).values
rows, cols = key >> shift, (key & ((1 << shift) - 1)).to(torch.int32)
deg = torch.bincount(rows, minlength=n)
indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
indptr[1:] = torch.cumsum(deg, 0)
val = torch.rand(rows.numel(), device=dev, generator=g)
graph = SpmmGraph(indptr, cols.contiguous(), val.contiguous())
E = torch.randn(n, d, device=dev) * 0.1
out = torch.empty_like(E)

rng = np.random.default_rng(0)
us = [int(x) for x in np.exp(rng.uniform(np.log(10), np.log(2e5), 50))]
its = [int(x) for x in np.exp(rng.uniform(np.log(10), np.log(2e5), 50))]
spec = tm.make_spec(rng, 1_000_000, 100_000, us, its, 5, 5, interleave=False)
tr = FMTrainer(spec, tm.make_fm_weights(rng, spec, 16, True), use_bn=True)
R = 1 << 17
users = torch.as_tensor(rng.integers(0, 1_000_000, R)).cuda()
items = torch.as_tensor(rng.integers(0, 100_000, R)).cuda()
labels = torch.as_tensor((rng.random(R) < 0.3).astype(np.float32)).cuda()
for _ in range(3):
    graph.spmm(E, out=out)
    tr.step(users, items, labels)
torch.cuda.synchronize()
print(f"done nnz={graph.nnz} rows={n} d={d} feat_rows={R} F_s=100 F_d=10 K=16")
