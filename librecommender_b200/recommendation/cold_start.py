"""Cold-start recommendations for unknown users — host-side seam of
``libreco/recommendation/cold_start.py`` (called from ``recommend_user`` of every model base).

Behaviour kept from the reference: one draw WITH replacement of ``n_rec`` entries
per cold user from a fixed pool, consuming ``data_info.np_rng`` in user order —
strategy ``"average"`` draws from the model's ``default_recs`` (inner ids: the top items of the
mean-embedding user), ``"popular"`` from ``data_info.popular_items`` (original ids).  The pool and
the id conversion are resolved once per call instead of once per user."""
import numpy as np

_STRATEGIES = ("average", "popular")


def _pool_and_mapping(data_info, default_recs, strategy, inner_id):
    """(candidate pool, dict that converts a drawn entry to the requested id space or None)."""
    if strategy == "average":
        return np.asarray(default_recs), (None if inner_id else data_info.id2item)
    return np.asarray(data_info.popular_items), (data_info.item2id if inner_id else None)


def _draw(data_info, pool, mapping, n_rec):
    picked = data_info.np_rng.choice(pool, n_rec)
    if mapping is None:
        return picked
    return np.array([mapping[i] for i in picked])


def popular_recommendations(data_info, inner_id, n_rec):
    """One draw from the popular items (the reference's helper of the same name)."""
    pool, mapping = _pool_and_mapping(data_info, None, "popular", inner_id)
    return _draw(data_info, pool, mapping, n_rec)


def cold_start_rec(data_info, default_recs, cold_start, users, n_rec, inner_id):
    if cold_start not in _STRATEGIES:
        raise ValueError(f"Unknown cold start strategy: {cold_start}")
    pool, mapping = _pool_and_mapping(data_info, default_recs, cold_start, inner_id)
    return {u: _draw(data_info, pool, mapping, n_rec) for u in users}
