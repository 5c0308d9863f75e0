"""CPU restatement of the reference's training losses (TEST INFRASTRUCTURE ONLY — never imported by
the product path).

* torch half (libreco/torchops/loss.py:5-90): PINNED by golden vectors produced with the
  unmodified reference (tests/golden/gen_losses.py -> tests/golden/losses.npz).
* TF half (libreco/tfops/loss.py:4-71, algorithms/two_tower.py:458-479): TensorFlow is absent;
  the sigmoid-CE / focal / max-margin formulas coincide with the torch half (pinned through it),
  the in-batch softmax with logit adjustment is PARITY UNPINNED (restated from the TF op
  semantics: divide_no_nan, clip_by_value, sparse_softmax_cross_entropy_with_logits).

All functions evaluate in the dtype of their inputs (pass float64 for the high-precision check).
"""
import numpy as np


def _bce(x, y):
    # torch / TF stable form: max(x, 0) - x*y + log1p(exp(-|x|))
    return np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    return np.where(x >= 0, 1 / (1 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1 + np.exp(-np.abs(x))))


def binary_cross_entropy_loss(logits, labels):
    """torchops/loss.py:5-6."""
    return _bce(logits, labels).mean()


def focal_elementwise(logits, labels, alpha=0.25, gamma=2.0):
    """torchops/loss.py:10-16 = tfops/loss.py:52-58."""
    w = labels * alpha + (1 - labels) * (1 - alpha)
    p = _sigmoid(logits)
    p_t = labels * p + (1 - labels) * (1 - p)
    return w * np.power(1.0 - p_t, gamma) * _bce(logits, labels)


def focal_loss(logits, labels, alpha=0.25, gamma=2.0):
    return focal_elementwise(logits, labels, alpha, gamma).mean()


def mean_squared_error(pred, labels):
    """tfops/loss.py:5-8."""
    return np.square(pred - labels).mean()


def bpr_loss(pos, neg):
    """torchops/loss.py:22-24: -mean(logsigmoid(pos - neg))."""
    d = pos - neg
    return -(np.minimum(d, 0) - np.log1p(np.exp(-np.abs(d)))).mean()


def max_margin_loss(pos, neg, margin):
    """torchops/loss.py:27-30 (target = 1): mean(max(0, -(pos - neg) + margin))."""
    return np.maximum(0, margin - (pos - neg)).mean()


def pairwise_bce_loss(pos, neg, mean=True):
    """torchops/loss.py:33-46."""
    v = np.concatenate([_bce(pos, np.ones_like(pos)), _bce(neg, np.zeros_like(neg))])
    return v.mean() if mean else v.sum()


def pairwise_focal_loss(pos, neg, mean=True):
    """torchops/loss.py:49-60."""
    v = np.concatenate([focal_elementwise(pos, np.ones_like(pos)), focal_elementwise(neg, np.zeros_like(neg))])
    return v.mean() if mean else v.sum()


def compute_pair_scores(targets, items_pos, items_neg, repeat_positives=True):
    """torchops/loss.py:63-90."""
    if len(targets) == len(items_pos) == len(items_neg):
        return (targets * items_pos).sum(1), (targets * items_neg).sum(1)
    factor = len(items_neg) // len(items_pos)
    pos = (targets * items_pos).sum(1)
    if repeat_positives:
        pos = np.repeat(pos, factor)
    neg = (targets[:, None, :] * items_neg.reshape(len(items_pos), factor, -1)).sum(2).ravel()
    return pos, neg


def adjust_logits(logits, temperature, correction=None, item_indices=None):
    """two_tower.py:458-479 with all_adjust=True."""
    logits = logits / temperature if temperature != 0 else np.zeros_like(logits)
    if correction is not None:
        logits = logits - np.log(np.clip(correction, 1e-8, 1.0)).reshape(1, -1)
    if item_indices is not None:
        eq = item_indices.reshape(1, -1) == item_indices.reshape(-1, 1)
        mask = eq & ~np.eye(len(item_indices), dtype=bool)
        logits = np.where(mask, np.finfo(np.float32).min, logits)
    return logits


def softmax_cross_entropy(user_embeds, item_embeds, temperature=1.0, correction=None, item_indices=None):
    """tfops/loss.py:67-71 + reduce_mean (:37-39): labels = arange(B)."""
    logits = adjust_logits(user_embeds @ item_embeds.T, temperature, correction, item_indices)
    mx = logits.max(axis=1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(axis=1))
    return (lse - np.diag(logits)).mean()
