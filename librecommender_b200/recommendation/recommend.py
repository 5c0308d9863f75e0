"""``recommend_from_embedding`` / ``construct_rec`` — drop-ins for
``libreco/recommendation/recommend.py:8-18,57-78``."""
from __future__ import annotations

import numpy as np

from ..engine import scorer_for
from .ranking import rank_recommendations


def construct_rec(data_info, user_ids, computed_recs, inner_id):
    """recommend.py:8-18 — inner ids → original ids (vectorised through one lookup array)."""
    out = {}
    if inner_id:
        for r, u in enumerate(user_ids):
            out[u] = np.array(computed_recs[r])
        return out
    id2item = data_info.id2item
    for r, u in enumerate(user_ids):
        out[data_info.id2user[u]] = np.array([id2item[i] for i in computed_recs[r]])
    return out


def recommend_from_embedding(
    model,
    user_ids,
    n_rec,
    user_embeddings,
    item_embeddings,
    filter_consumed,
    random_rec,
):
    """Same contract as the reference: ``int64[B, n_rec]`` inner item ids."""
    if n_rec > model.n_items:
        raise ValueError(f"`n_rec` {n_rec} exceeds num of items {model.n_items}")
    scorer = scorer_for(model, user_embeddings, item_embeddings)
    if random_rec:
        import torch

        uid = torch.as_tensor(np.asarray(user_ids, dtype=np.int64)).to(scorer.device)
        rows = scorer.score_rows(uid)
        return rank_recommendations(model.task, user_ids, rows, n_rec, model.n_items,
                                    scorer.csr, filter_consumed, True)
    return scorer.recommend(user_ids, n_rec, filter_consumed)
