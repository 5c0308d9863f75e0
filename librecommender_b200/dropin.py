"""Drop-in wiring: patch an imported reference package (``libreco``) so that its OWN classes run the
hot path on this library — the code form of INTEGRATION.md ("What a reference maintainer changes").

    import libreco
    from librecommender_b200 import dropin
    dropin.install(libreco)            # LightGCN(...).fit(...); model.recommend_user(...) now run on the GPU
    dropin.uninstall()                 # the reference's numpy / torch-CPU path again

What is patched (seams of SURVEY.md §8b; nothing else of the reference changes):

* ``libreco.recommendation.{rank_recommendations, recommend_from_embedding, construct_rec}`` and the
  copies of those names that ``bases/embed_base.py:10``, ``bases/dyn_embed_base.py:8`` and
  ``recommendation/recommend.py`` hold — so ``EmbedBase.fit`` (default_recs, ``embed_base.py:153-161``),
  ``EmbedBase.recommend_user`` (``:190-251``) and ``DynEmbedBase.recommend_user`` call the CUDA path;
* ``libreco.algorithms.lightgcn.LightGCNModel`` (``algorithms/lightgcn.py:4,117-127``) → the
  differentiable K6 module with the reference's constructor / ``forward(use_dropout)`` contract;
* optionally the loss functions ``TorchTrainer._compute_loss`` uses
  (``training/torch_trainer.py:15-23,140-161``).
"""
from __future__ import annotations

import importlib

_saved: list = []


def _patch(mod, name, value):
    if hasattr(mod, name):
        _saved.append((mod, name, getattr(mod, name)))
        setattr(mod, name, value)


def install(libreco=None, losses: bool = True, lightgcn: bool = True) -> None:
    """Patch the reference package in place (idempotent: a second call re-installs)."""
    from . import recommendation as rec

    if libreco is None:
        libreco = importlib.import_module("libreco")
    uninstall()
    base = libreco.__name__
    mods = {}
    for sub in ("recommendation", "recommendation.recommend", "bases.embed_base", "bases.dyn_embed_base",
                "algorithms.lightgcn", "training.torch_trainer", "torchops"):
        try:
            mods[sub] = importlib.import_module(f"{base}.{sub}")
        except Exception:                      # a sub-module the installed reference cannot import
            mods[sub] = None
    for sub in ("recommendation", "recommendation.recommend", "bases.embed_base", "bases.dyn_embed_base"):
        m = mods[sub]
        if m is None:
            continue
        _patch(m, "rank_recommendations", rec.rank_recommendations)
        _patch(m, "recommend_from_embedding", rec.recommend_from_embedding)
        _patch(m, "construct_rec", rec.construct_rec)
    if lightgcn and mods["algorithms.lightgcn"] is not None:
        from .lightgcn import make_lightgcn_model_class

        _patch(mods["algorithms.lightgcn"], "LightGCNModel", make_lightgcn_model_class())
    if losses:
        from . import losses as L

        for sub in ("training.torch_trainer", "torchops"):
            m = mods[sub]
            if m is None:
                continue
            for name in ("binary_cross_entropy_loss", "bpr_loss", "compute_pair_scores", "focal_loss",
                         "max_margin_loss", "pairwise_bce_loss", "pairwise_focal_loss"):
                if hasattr(L, name):
                    _patch(m, name, getattr(L, name))


def uninstall() -> None:
    """Restore every patched name."""
    while _saved:
        mod, name, old = _saved.pop()
        setattr(mod, name, old)


def installed() -> bool:
    return bool(_saved)
