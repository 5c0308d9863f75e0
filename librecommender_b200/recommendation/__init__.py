"""Drop-in replacements for ``libreco.recommendation`` (same names, same signatures)."""
from .cold_start import cold_start_rec, popular_recommendations
from .ranking import rank_recommendations
from .recommend import check_dynamic_rec_feats, construct_rec, recommend_from_embedding, recommend_tf_feat

__all__ = [
    "check_dynamic_rec_feats",
    "cold_start_rec",
    "construct_rec",
    "popular_recommendations",
    "rank_recommendations",
    "recommend_from_embedding",
    "recommend_tf_feat",
]
