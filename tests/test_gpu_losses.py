"""GPU parity of the fused loss kernels (csrc/loss.cu, librecommender_b200/losses.py) against the
golden vectors of the unmodified reference (torchops/loss.py: values and autograd gradients) and the
float64 oracle.  Tolerance: 1e-6 relative on the loss value, 1e-6 absolute (x 1/n scale) on
gradients — fp32 transcendental error, the reductions themselves are exact to double."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))
F = 3


def _t(a, grad=False):
    import torch

    t = torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")
    t.requires_grad_(grad)
    return t


def _check(val, ref, tol=2e-6):
    assert abs(float(val) - ref) <= tol * max(abs(ref), 1.0), (float(val), ref)


def _check_grad(g, ref, scale):
    g = g.detach().cpu().numpy().astype(np.float64)
    assert g.shape == ref.shape
    assert np.abs(g - ref).max() <= 2e-6 * scale, float(np.abs(g - ref).max())


def test_pointwise_losses_match_reference_values_and_grads():
    from librecommender_b200 import losses as L

    n = len(G["logits"])
    for name, fn, key in (("bce", lambda x, y: L.binary_cross_entropy_loss(x, y), "bce"),
                          ("focal", lambda x, y: L.focal_loss(x, y), "focal"),
                          ("focal2", lambda x, y: L.focal_loss(x, y, alpha=0.4, gamma=1.5), "focal_a4_g15")):
        x, y = _t(G["logits"], True), _t(G["labels"])
        v = fn(x, y)
        _check(v, float(G[key]))
        (v * 3.0).backward()            # upstream gradient is honoured
        _check_grad(x.grad / 3.0, G[key + "_grad"], 1.0 / n * 4)


def test_mse_matches_oracle():
    from librecommender_b200 import losses as L
    from oracle import losses as ol

    rng = np.random.default_rng(1)
    p, y = rng.standard_normal(5000) * 2 + 3, rng.integers(1, 6, 5000).astype(np.float64)
    x = _t(p, True)
    v = L.mean_squared_error(x, _t(y))
    _check(v, ol.mean_squared_error(p.astype(np.float32).astype(np.float64), y))
    v.backward()
    _check_grad(x.grad, 2 * (p.astype(np.float32).astype(np.float64) - y) / 5000, 1e-2)


@pytest.mark.parametrize("repeat", [True, False])
def test_rank_losses_match_reference(repeat):
    """bpr / max-margin with positives repeated by the caller (reference) or broadcast in-kernel."""
    import torch

    from librecommender_b200 import losses as L

    m = len(G["pos"])
    for fn, key in ((lambda p, q: L.bpr_loss(p, q), "bpr"), (lambda p, q: L.max_margin_loss(p, q, 1.0), "mm")):
        pos, neg = _t(G["pos"], True), _t(G["neg"], True)
        p_in = pos.repeat_interleave(F) if repeat else pos
        v = fn(p_in, neg)
        _check(v, float(G[key]))
        v.backward()
        _check_grad(pos.grad, G[key + "_gpos"], F / (m * F))
        _check_grad(neg.grad, G[key + "_gneg"], 1.0 / (m * F))


@pytest.mark.parametrize("mean", [1, 0])
def test_pairwise_class_losses_match_reference(mean):
    from librecommender_b200 import losses as L

    n = len(G["pos"]) + len(G["neg"])
    for fn, key in ((L.pairwise_bce_loss, "pbce"), (L.pairwise_focal_loss, "pfocal")):
        pos, neg = _t(G["pos"], True), _t(G["neg"], True)
        v = fn(pos, neg, mean=bool(mean))
        _check(v, float(G[f"{key}_{mean}"]))
        v.backward()
        sc = 1.0 / n if mean else 1.0
        _check_grad(pos.grad, G[f"{key}_{mean}_gpos"], sc)
        _check_grad(neg.grad, G[f"{key}_{mean}_gneg"], sc)


def test_compute_pair_scores_matches_reference():
    from librecommender_b200 import losses as L

    for rp in (1, 0):
        ps, ns = L.compute_pair_scores(_t(G["T"]), _t(G["P"]), _t(G["N"]), bool(rp))
        np.testing.assert_allclose(ps.cpu().numpy(), G[f"pair_pos_{rp}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ns.cpu().numpy(), G[f"pair_neg_{rp}"], rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        L.compute_pair_scores(_t(G["T"]), _t(G["P"][:-1]), _t(G["N"]))
    with pytest.raises(ValueError):
        L.compute_pair_scores(_t(G["T"]), _t(G["P"]), _t(G["N"][:-1]))


@pytest.mark.parametrize("B,d,temp,use_corr,use_ids", [(257, 16, 1.0, False, False), (1024, 64, 0.1, True, True),
                                                       (100, 8, 0.5, True, False), (4500, 32, 0.2, False, True)])
def test_inbatch_softmax_matches_oracle(B, d, temp, use_corr, use_ids):
    """TwoTower in-batch softmax + logQ correction + accidental-hit removal (oracle: unpinned TF half),
    value and gradients w.r.t. both towers against float64 finite-difference-free analytic grads."""
    import torch

    from librecommender_b200 import losses as L
    from oracle import losses as ol

    rng = np.random.default_rng(B)
    U = rng.standard_normal((B, d)).astype(np.float32)
    I = rng.standard_normal((B, d)).astype(np.float32)
    U /= np.linalg.norm(U, axis=1, keepdims=True)
    I /= np.linalg.norm(I, axis=1, keepdims=True)
    corr = (rng.random(B) * 0.01 + 1e-4).astype(np.float32) if use_corr else None
    ids = rng.integers(0, B // 3, B) if use_ids else None
    ref = ol.softmax_cross_entropy(U.astype(np.float64), I.astype(np.float64), temp,
                                   corr.astype(np.float64) if use_corr else None, ids)
    u, i = _t(U, True), _t(I, True)
    v = L.softmax_cross_entropy(u, i, temp, _t(corr) if use_corr else None,
                                torch.tensor(ids, device="cuda") if use_ids else None)
    _check(v, ref, 5e-6)
    v.backward()
    # analytic float64 gradient of the oracle
    lg = ol.adjust_logits(U.astype(np.float64) @ I.astype(np.float64).T, temp,
                          corr.astype(np.float64) if use_corr else None, ids)
    p = np.exp(lg - lg.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    Gm = (p - np.eye(B)) / temp / B
    if use_ids:
        Gm[(ids[None, :] == ids[:, None]) & ~np.eye(B, dtype=bool)] = 0.0
    _check_grad(u.grad, Gm @ I.astype(np.float64), 5.0 / B / temp)
    _check_grad(i.grad, Gm.T @ U.astype(np.float64), 5.0 / B / temp)


def test_losses_reject_cpu_tensors_and_bad_shapes():
    import torch

    from librecommender_b200 import _lib
    from librecommender_b200 import losses as L

    with pytest.raises(_lib.B200Error):
        L.bpr_loss(torch.zeros(4), torch.zeros(4))
    with pytest.raises(_lib.B200Error):
        L.bpr_loss(torch.zeros(4, device="cuda"), torch.zeros(7, device="cuda"))
    with pytest.raises(ValueError):
        L.binary_cross_entropy_loss(torch.zeros(4, device="cuda"), torch.zeros(5, device="cuda"))
