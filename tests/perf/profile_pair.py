"""Driver for ncu captures of the hoisted DeepFM pair kernel (csrc/pair.cu).
    ncu --set full --clock-control none --import-source on -k regex:deepfm_pair -s 1 -c 1 \
        -o gpurun_out/prof_pair python tests/perf/profile_pair.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from librecommender_b200.feat_models import DeepFM  # noqa: E402
from oracle import tf_models as tm  # noqa: E402

rng = np.random.default_rng(0)
spec = tm.make_spec(rng, 20000, 100_000, [50, 1000], [1000, 10000, 500], 1, 1, interleave=False)
w = tm.make_deepfm_weights(rng, spec, 16, (128, 64, 32), True)
model = DeepFM(spec, w)
uid = torch.as_tensor(rng.integers(0, 20000, 64)).cuda()
for _ in range(3):
    model.score_all_items(uid)
torch.cuda.synchronize()
print("done")
