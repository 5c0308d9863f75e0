"""Golden vectors for the training losses from the UNMODIFIED reference
(libreco/torchops/loss.py, torch CPU): values in float32 and float64 plus autograd gradients.

    python tests/golden/gen_losses.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from libreco.torchops import loss as ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _val_and_grads(fn, *tensors):
    ts = [torch.tensor(t, dtype=torch.float64, requires_grad=True) for t in tensors]
    v = fn(*ts)
    v.backward()
    v32 = fn(*[torch.tensor(t, dtype=torch.float32) for t in tensors])
    return float(v), float(v32), [t.grad.numpy() for t in ts]


if __name__ == "__main__":
    g = np.random.default_rng(2024)
    data = {}
    n = 4099
    logits = (g.standard_normal(n) * 4).astype(np.float32)
    logits[:6] = [0.0, 40.0, -40.0, 90.0, -90.0, 1e-4]            # both tails
    labels = (g.random(n) < 0.3).astype(np.float32)
    data["logits"], data["labels"] = logits, labels
    v, v32, (gl,) = _val_and_grads(lambda x: ref.binary_cross_entropy_loss(x, torch.tensor(labels, dtype=x.dtype)), logits)
    data["bce"], data["bce_f32"], data["bce_grad"] = v, v32, gl
    v, v32, (gl,) = _val_and_grads(lambda x: ref.focal_loss(x, torch.tensor(labels, dtype=x.dtype)), logits)
    data["focal"], data["focal_f32"], data["focal_grad"] = v, v32, gl
    v, v32, (gl,) = _val_and_grads(
        lambda x: ref.focal_loss(x, torch.tensor(labels, dtype=x.dtype), alpha=0.4, gamma=1.5), logits)
    data["focal_a4_g15"], data["focal_a4_g15_grad"] = v, gl

    m, f = 1031, 3
    pos = (g.standard_normal(m) * 3).astype(np.float32)
    neg = (g.standard_normal(m * f) * 3).astype(np.float32)
    pos[:3], neg[:3] = [50.0, -50.0, 0.0], [-50.0, 50.0, 0.0]
    data["pos"], data["neg"] = pos, neg
    rep = lambda p: p.repeat_interleave(f)                           # noqa: E731
    v, v32, (gp, gn) = _val_and_grads(lambda p, q: ref.bpr_loss(rep(p), q), pos, neg)
    data["bpr"], data["bpr_f32"], data["bpr_gpos"], data["bpr_gneg"] = v, v32, gp, gn
    v, v32, (gp, gn) = _val_and_grads(lambda p, q: ref.max_margin_loss(rep(p), q, 1.0), pos, neg)
    data["mm"], data["mm_f32"], data["mm_gpos"], data["mm_gneg"] = v, v32, gp, gn
    for mean in (True, False):
        v, v32, (gp, gn) = _val_and_grads(lambda p, q: ref.pairwise_bce_loss(p, q, mean=mean), pos, neg)
        data[f"pbce_{int(mean)}"], data[f"pbce_{int(mean)}_gpos"], data[f"pbce_{int(mean)}_gneg"] = v, gp, gn
        v, v32, (gp, gn) = _val_and_grads(lambda p, q: ref.pairwise_focal_loss(p, q, mean=mean), pos, neg)
        data[f"pfocal_{int(mean)}"], data[f"pfocal_{int(mean)}_gpos"], data[f"pfocal_{int(mean)}_gneg"] = v, gp, gn

    d = 16
    T = g.standard_normal((m, d)).astype(np.float32)
    P = g.standard_normal((m, d)).astype(np.float32)
    N = g.standard_normal((m * f, d)).astype(np.float32)
    data["T"], data["P"], data["N"] = T, P, N
    for rp in (True, False):
        ps, ns = ref.compute_pair_scores(torch.tensor(T), torch.tensor(P), torch.tensor(N), repeat_positives=rp)
        data[f"pair_pos_{int(rp)}"], data[f"pair_neg_{int(rp)}"] = ps.numpy(), ns.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **data)
    print("wrote losses.npz:", {k: (np.shape(v) if np.ndim(v) else float(v)) for k, v in data.items() if np.ndim(v) == 0})
