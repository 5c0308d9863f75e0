"""Weight interchange with the reference's on-disk formats (SURVEY.md §8b "Weight interchange",
Appendix C) — pure host code:

* ``<model_name>.npz`` with ``user_embed`` / ``item_embed`` (``EmbedBase.save(inference_only=True)``,
  ``libreco/bases/embed_base.py:289-295``; read back by ``EmbedBase.load``, :323-330);
* ``<model_name>_tf_variables.npz`` keyed by TF variable name (``utils/save_load.py:70-98``): the
  embedding-scope names are fixed by the reference's graph code, the un-named ``tf_dense`` /
  batch-norm variables carry TensorFlow's auto-generated names, so those go through an explicit
  name map;
* ``<model_name>_default_recs.npz`` (``utils/save_load.py:39-48``).
"""
from __future__ import annotations

import os

import numpy as np

EMBEDDING_SCOPE = {
    "user_embeds": "embedding/user_embeds_var:0", "item_embeds": "embedding/item_embeds_var:0",
    "sparse_embeds": "embedding/sparse_embeds_var:0", "dense_embeds": "embedding/dense_embeds_var:0",
    "user_linear": "embedding/user_linear_var:0", "item_linear": "embedding/item_linear_var:0",
    "sparse_linear": "embedding/sparse_linear_var:0", "dense_linear": "embedding/dense_linear_var:0",
}


def save_embed_model(path, model_name, user_embed, item_embed):
    """Write what ``EmbedBase.save(path, model_name, inference_only=True)`` writes for the variables."""
    os.makedirs(path, exist_ok=True)
    np.savez_compressed(os.path.join(path, model_name), user_embed=np.asarray(user_embed, dtype=np.float32),
                        item_embed=np.asarray(item_embed, dtype=np.float32))


def load_embed_model(path, model_name):
    """(user_embed, item_embed) of a reference-saved embed model (last rows = OOV)."""
    v = np.load(os.path.join(path, f"{model_name}.npz"))
    return v["user_embed"], v["item_embed"]


def to_tf_variables(weights, extra_names=None):
    """Engine weight dict -> ``{tf variable name: array}``.  Tables use the fixed embedding-scope
    names; every other entry needs a name in `extra_names` ({engine key: tf name}).  Shapes follow the
    reference's ``var_shape``s: ``user_linear_var`` / ``item_linear_var`` are ``[V, 1]``
    (``fm.py:181-196``), ``sparse_linear_var`` / ``dense_linear_var`` stay 1-D
    (``[sparse_feature_size]`` / ``[dense_field_size]``, ``fm.py:219-249``, ``deepfm.py:224-256``)."""
    out = {}
    for k, name in EMBEDDING_SCOPE.items():
        if weights.get(k) is not None:
            a = np.asarray(weights[k], dtype=np.float32)
            if k in ("user_linear", "item_linear"):
                a = a.reshape(-1, 1)
            elif k in ("sparse_linear", "dense_linear"):
                a = a.reshape(-1)
            out[name] = a
    for k, name in (extra_names or {}).items():
        out[name] = np.asarray(weights[k], dtype=np.float32)
    return out


def save_tf_variables(path, model_name, weights, extra_names=None):
    os.makedirs(path, exist_ok=True)
    np.savez_compressed(os.path.join(path, f"{model_name}_tf_variables"), **to_tf_variables(weights, extra_names))


def load_tf_variables(path, model_name, extra_names=None):
    """Inverse of :func:`save_tf_variables` (also reads files written by the reference)."""
    from .feat_models import from_tf_variables

    npz = np.load(os.path.join(path, f"{model_name}_tf_variables.npz"))
    w = from_tf_variables(npz)
    for k, name in (extra_names or {}).items():
        w[k] = npz[name]
    return w


def save_default_recs(path, model_name, default_recs):
    np.savez_compressed(os.path.join(path, f"{model_name}_default_recs"), default_recs=np.asarray(default_recs))


def load_default_recs(path, model_name):
    return np.load(os.path.join(path, f"{model_name}_default_recs.npz"))["default_recs"]
