"""Device-side batch collation — the part of
``libreco/batch/collators.py`` that sits between the interaction arrays and the model forward:

* ``PointwiseCollator.__call__`` (:225-252): users / items repeated ``num_neg + 1`` times, labels
  ``1, 0, 0, …``, the ``num_neg`` negatives of positive ``j`` interleaved right after it (:231-232);
* ``PairwiseCollator.__call__`` (:277-299): ``(users, items_pos, items_neg)`` with the positives
  repeated ``num_neg`` times when ``repeat_positives``.

Negatives come from :class:`~librecommender_b200.sampling.DeviceNegativeSampler` (fast mode) or from
a host array produced by the parity samplers.  The per-row FEATURE gathering of the reference
(``get_pointwise_feats`` / ``get_sampled_item_feats``, :254-267,460-490) has no counterpart here on
purpose: the feature kernels (``b200_feat_forward``) read the per-user / per-item unique tables
themselves from ``(user, item)``, so a collated batch is just the three id / label vectors.
Index bookkeeping only (repeat / interleave on device tensors); the sampling itself is the CUDA
kernel."""
from __future__ import annotations

import numpy as np


def adjust_batch_size(batch_size: int, num_neg: int, pairwise: bool) -> int:
    """libreco/batch/batch_data.py:93-105 — positives per step."""
    return max(1, int(batch_size / (num_neg if pairwise else num_neg + 1)))


class DevicePointwiseCollator:
    def __init__(self, sampler, num_neg: int, sampler_name: str = "random"):
        self.sampler, self.num_neg, self.sampler_name = sampler, int(num_neg), sampler_name

    def __call__(self, users_d, items_d, negatives_d=None):
        import torch

        r = self.num_neg + 1
        n = users_d.numel()
        if negatives_d is None:
            negatives_d = self.sampler.sample(users_d, items_d, self.num_neg, self.sampler_name)
        users = users_d.repeat_interleave(r)
        items = torch.empty(n * r, dtype=torch.int64, device=users_d.device)
        items.view(n, r)[:, 0] = items_d
        items.view(n, r)[:, 1:] = negatives_d.view(n, self.num_neg)
        labels = torch.zeros(n * r, dtype=torch.float32, device=users_d.device)
        labels[::r] = 1.0
        return users, items, labels


class DevicePairwiseCollator:
    def __init__(self, sampler, num_neg: int, sampler_name: str = "random", repeat_positives: bool = True):
        self.sampler, self.num_neg = sampler, int(num_neg)
        self.sampler_name, self.repeat_positives = sampler_name, repeat_positives

    def __call__(self, users_d, items_d, negatives_d=None):
        if negatives_d is None:
            negatives_d = self.sampler.sample(users_d, items_d, self.num_neg, self.sampler_name)
        if self.repeat_positives and self.num_neg > 1:
            return users_d.repeat_interleave(self.num_neg), items_d.repeat_interleave(self.num_neg), negatives_d
        return users_d, items_d, negatives_d
