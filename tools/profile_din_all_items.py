"""Launch list of DIN all-items scoring (feat_models.DIN.score_all_items) for a few users at the C4-like shape:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_din_all_items.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librecommender_b200 import synthetic as syn  # noqa: E402
from librecommender_b200.consumed import ConsumedCSR  # noqa: E402
from librecommender_b200.feat_models import DIN, recent_sequences_csr  # noqa: E402

rng = np.random.default_rng(0)
n_users, n_items, T = 200_000, 100_000, 50
spec = syn.make_spec(rng, n_users, n_items, [50, 1000], [1000, 10000, 100000], 1, 0, interleave=False)
deg = np.minimum(rng.poisson(80, n_users), 1000).astype(np.int64) + 1
indptr = np.concatenate([[0], np.cumsum(deg)])
idx = rng.integers(0, n_items, indptr[-1]).astype(np.int32)
csr = ConsumedCSR(indptr, idx)
seqs, lens = recent_sequences_csr(csr, n_items, T)
w = syn.make_seq_weights(rng, spec, 16, (128, 64, 32), True, din=True)
model = DIN(spec, w, seqs, lens, csr)
uid = rng.integers(0, n_users, int(os.environ.get("DIN_USERS", "4")))
model.recommend(uid[:2], 100, True)
model.recommend(uid, 100, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
model.recommend(uid, 100, True)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ms per user", e0.elapsed_time(e1) / len(uid))
