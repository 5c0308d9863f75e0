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


# ------------------------------------------------------------------------------------------------
# TensorFlow auto-generated variable names of the un-named layers
# ------------------------------------------------------------------------------------------------
# The reference never names its heads: ``tf_dense(units=1)`` (layers/dense.py:52-80) becomes a
# ``tf.keras.layers.Dense`` / ``tf.layers.dense`` whose variables TensorFlow names ``dense``,
# ``dense_1``, ... in CREATION ORDER; ``tf.layers.batch_normalization`` likewise
# (``batch_normalization``, ``batch_normalization_1``, ...), prefixed by the enclosing
# ``tf.variable_scope`` — ``dense_nn`` opens ``<name>`` (default "mlp") and names its Dense layers
# ``<name>_layer<i>`` (layers/dense.py:28-33).  The creation order per model is read off the graph
# builders: fm.py:152-171, deepfm.py:158-174, din.py:205-218 (+ the "attention" dense_nn,
# layers/attention.py:47-53), youtube_ranking.py:208-217, two_tower.py:400-409.  No TensorFlow exists in
# this environment, so this table is RESTATED from TensorFlow's documented uniquifying rule and is
# unverified against a real checkpoint; ``resolve_tf_names`` therefore checks every expected name AND
# shape against the file and reports exactly what is missing instead of guessing.
def _bn_names(prefix):
    return {k: f"{prefix}/{v}:0" for k, v in (("gamma", "gamma"), ("beta", "beta"), ("mean", "moving_mean"),
                                             ("var", "moving_variance"))}


def _mlp_names(scope, n_layers, use_bn):
    names = {"kernels": [f"{scope}/{scope}_layer{i}/kernel:0" for i in range(1, n_layers + 1)],
             "biases": [f"{scope}/{scope}_layer{i}/bias:0" for i in range(1, n_layers + 1)]}
    if use_bn:
        names["bn_in"] = _bn_names(f"{scope}/batch_normalization")
        names["bns"] = [_bn_names(f"{scope}/batch_normalization_{i}") for i in range(1, n_layers)]
    return names


def default_tf_names(model_name, n_hidden, use_bn, use_tf_attention=False):
    """{engine weight key: TF variable name (or nested dict / list of names)} for the auto-named
    variables of `model_name` in {"FM", "DeepFM", "DIN", "YouTubeRanking", "TwoTower"}."""
    if model_name == "FM":
        out = {"lin_kernel": "dense/kernel:0", "lin_bias": "dense/bias:0",
               "pw_kernel": "dense_1/kernel:0", "pw_bias": "dense_1/bias:0"}
        if use_bn:
            out["fm_bn"] = _bn_names("batch_normalization")
        return out
    if model_name == "DeepFM":
        return {"lin_kernel": "dense/kernel:0", "lin_bias": "dense/bias:0", "mlp": _mlp_names("mlp", n_hidden, use_bn),
                "out_kernel": "dense_1/kernel:0", "out_bias": "dense_1/bias:0"}
    if model_name == "DIN":
        out = {"mlp": _mlp_names("mlp", n_hidden, use_bn), "out_kernel": "dense/kernel:0", "out_bias": "dense/bias:0"}
        if not use_tf_attention:
            out["attention"] = {"k1": "attention/attention_layer1/kernel:0", "b1": "attention/attention_layer1/bias:0",
                                "k2": "attention/attention_layer2/kernel:0", "b2": "attention/attention_layer2/bias:0"}
        return out
    if model_name == "YouTubeRanking":
        return {"mlp": _mlp_names("mlp", n_hidden, use_bn), "out_kernel": "dense/kernel:0", "out_bias": "dense/bias:0"}
    if model_name == "TwoTower":
        return {"user_tower": _mlp_names("user_tower", n_hidden, use_bn),
                "item_tower": _mlp_names("item_tower", n_hidden, use_bn)}
    raise ValueError(f"no TensorFlow name table for model `{model_name}`")


def resolve_tf_names(npz, names):
    """Read the (possibly nested) name table out of `npz`; a missing variable raises a ``KeyError`` that
    lists the expected name and the names the file does contain."""
    def take(n):
        if isinstance(n, dict):
            return {k: take(v) for k, v in n.items()}
        if isinstance(n, list):
            return [take(v) for v in n]
        if n not in npz:
            have = sorted(k for k in npz.files if not k.startswith("embedding/"))
            raise KeyError(f"TF variable `{n}` not in the file; non-embedding variables present: {have}")
        return np.asarray(npz[n])
    return take(names)


def load_reference_tf_model(path, model_name, arch, n_hidden, use_bn, use_tf_attention=False, extra_names=None):
    """Engine weight dict of a model saved by the reference (``save_tf_variables``,
    utils/save_load.py:70-98) WITHOUT a hand-written name map: the embedding-scope variables by their
    fixed names, the heads / MLPs / batch-norms through :func:`default_tf_names` (override single entries
    with `extra_names`)."""
    from .feat_models import from_tf_variables

    npz = np.load(os.path.join(path, f"{model_name}_tf_variables.npz"))
    w = from_tf_variables(npz)
    names = default_tf_names(arch, n_hidden, use_bn, use_tf_attention)
    names.update(extra_names or {})
    w.update(resolve_tf_names(npz, names))
    if use_tf_attention:
        w["use_tf_attention"] = True
    return w


def load_reference_wide_deep(path, model_name, n_hidden, use_bn):
    """Weight dict for :class:`feat_models.DeepFM` from a WideDeep model saved by the reference
    (``<model_name>_tf_variables.npz``; variables of ``libreco/algorithms/wide_deep.py:150-262``: ``embedding/
    {user,item,sparse,dense}_{wide,deep}_var``, ``wide_term``, the ``deep`` dense_nn stack, ``deep_term``) through
    :func:`feat_models.wide_deep_weights`."""
    from .feat_models import wide_deep_weights

    npz = np.load(os.path.join(path, f"{model_name}_tf_variables.npz"))

    def emb(name):
        key = f"embedding/{name}:0"
        return np.asarray(npz[key]) if key in npz else None

    rest = resolve_tf_names(npz, {"mlp": _mlp_names("deep", n_hidden, use_bn), "wide_kernel": "wide_term/kernel:0",
                                  "wide_bias": "wide_term/bias:0", "deep_kernel": "deep_term/kernel:0",
                                  "deep_bias": "deep_term/bias:0"})
    return wide_deep_weights(emb("user_wide_var"), emb("item_wide_var"), emb("sparse_wide_var"), emb("dense_wide_var"),
                             rest["wide_kernel"], rest["wide_bias"], emb("user_deep_var"), emb("item_deep_var"),
                             emb("sparse_deep_var"), emb("dense_deep_var"), rest["mlp"], rest["deep_kernel"],
                             rest["deep_bias"])
