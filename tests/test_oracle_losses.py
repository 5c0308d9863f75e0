"""Pins oracle/losses.py against golden vectors produced by the unmodified reference
(libreco/torchops/loss.py via tests/golden/gen_losses.py)."""
import os

import numpy as np

from oracle import losses as ol

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))
F = 3


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def test_pointwise_values_match_reference_float64():
    x, y = G["logits"].astype(np.float64), G["labels"].astype(np.float64)
    assert _rel(ol.binary_cross_entropy_loss(x, y), float(G["bce"])) < 1e-12
    assert _rel(ol.focal_loss(x, y), float(G["focal"])) < 1e-12
    assert _rel(ol.focal_loss(x, y, 0.4, 1.5), float(G["focal_a4_g15"])) < 1e-12


def test_pairwise_values_match_reference_float64():
    p, n = G["pos"].astype(np.float64), G["neg"].astype(np.float64)
    pr = np.repeat(p, F)
    assert _rel(ol.bpr_loss(pr, n), float(G["bpr"])) < 1e-12
    assert _rel(ol.max_margin_loss(pr, n, 1.0), float(G["mm"])) < 1e-12
    for mean in (1, 0):
        assert _rel(ol.pairwise_bce_loss(p, n, bool(mean)), float(G[f"pbce_{mean}"])) < 1e-12
        assert _rel(ol.pairwise_focal_loss(p, n, bool(mean)), float(G[f"pfocal_{mean}"])) < 1e-12


def test_compute_pair_scores_matches_reference():
    for rp in (1, 0):
        ps, ns = ol.compute_pair_scores(G["T"], G["P"], G["N"], bool(rp))
        np.testing.assert_allclose(ps, G[f"pair_pos_{rp}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ns, G[f"pair_neg_{rp}"], rtol=1e-5, atol=1e-5)
        assert ps.shape == G[f"pair_pos_{rp}"].shape


def test_softmax_cross_entropy_properties():
    """Unpinned TF half: sanity properties of the restatement (uniform logits -> log B; masking
    accidental hits can only lower the loss; correction shifts columns)."""
    rng = np.random.default_rng(0)
    B, d = 64, 8
    U, I = rng.standard_normal((B, d)), rng.standard_normal((B, d))
    assert abs(ol.softmax_cross_entropy(np.zeros((B, d)), I) - np.log(B)) < 1e-12
    ids = rng.integers(0, 20, B)
    assert ol.softmax_cross_entropy(U, I, 0.5, None, ids) <= ol.softmax_cross_entropy(U, I, 0.5) + 1e-12
    corr = rng.random(B)
    a = ol.softmax_cross_entropy(U, I, 1.0, corr)
    logits = U @ I.T - np.log(np.clip(corr, 1e-8, 1.0))[None, :]
    ref = (np.log(np.exp(logits).sum(1)) - np.diag(logits)).mean()
    assert abs(a - ref) < 1e-10
