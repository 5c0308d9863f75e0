"""Driver for launch lists / ncu captures of the DeepFM training step at the C3 shape:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_train_step.py
The step of interest is bracketed by cudaProfilerStart/Stop (use `--profile-from-start off`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librecommender_b200 import synthetic as syn  # noqa: E402
from librecommender_b200.training import DeepFMTrainer  # noqa: E402

rng = np.random.default_rng(5)
us = [int(x) for x in np.exp(rng.uniform(np.log(10), np.log(2e5), 50))]
its = [int(x) for x in np.exp(rng.uniform(np.log(10), np.log(2e5), 50))]
spec = syn.make_spec(rng, 1_000_000, 100_000, us, its, 5, 5, interleave=False)
w = syn.make_deepfm_weights(rng, spec, 16, (128, 64, 32), True)
tr = DeepFMTrainer(spec, w, use_bn=True, lr=1e-3)
B = 8192
u = torch.as_tensor(rng.integers(0, 1_000_000, B)).cuda()
i = torch.as_tensor(rng.integers(0, 100_000, B)).cuda()
y = torch.as_tensor((rng.random(B) < 1 / 6).astype(np.float32)).cuda()
for _ in range(3):
    tr.step(u, i, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(u, i, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
