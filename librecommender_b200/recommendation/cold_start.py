"""Cold-start recommendations — host-side, mirrors
``libreco/recommendation/cold_start.py:4-31`` (draws WITH replacement from
``default_recs`` / ``popular_items`` using ``data_info.np_rng``)."""
import numpy as np


def popular_recommendations(data_info, inner_id, n_rec):
    picked = data_info.np_rng.choice(data_info.popular_items, n_rec)
    if inner_id:
        return np.array([data_info.item2id[i] for i in picked])
    return picked


def _average_recommendations(data_info, default_recs, inner_id, n_rec):
    picked = data_info.np_rng.choice(default_recs, n_rec)
    if inner_id:
        return picked
    return np.array([data_info.id2item[i] for i in picked])


def cold_start_rec(data_info, default_recs, cold_start, users, n_rec, inner_id):
    if cold_start not in ("average", "popular"):
        raise ValueError(f"Unknown cold start strategy: {cold_start}")
    out = {}
    for u in users:
        if cold_start == "average":
            out[u] = _average_recommendations(data_info, default_recs, inner_id, n_rec)
        else:
            out[u] = popular_recommendations(data_info, inner_id, n_rec)
    return out
