"""FM training step on the device: the reference's graph in training mode
(``libreco/algorithms/fm.py:140-172``), mean sigmoid cross entropy (``tfops/loss.py:14-18``) and
``tf.train.AdamOptimizer`` grouped with the batch-norm update ops
(``libreco/training/tf_trainer.py:112-123``), one call per mini-batch:

    gather + FM term (b200_feat_forward)  ->  BN with batch statistics (b200_bn_train_forward)
    -> Dense(1, elu) head (b200_fm_head_forward) -> loss + d loss / d logit (b200_pointwise_loss)
    -> head backward (b200_fm_head_backward) -> scatter of the field gradients (b200_feat_backward)
    -> TF-Adam over every variable (b200_adam_dense)

Every variable, Adam slot and gradient buffer is a device tensor; nothing returns to the host
during a step (the loss stays a device scalar).  Embedding variables follow TensorFlow's
``_apply_sparse_shared`` semantics: m and v are decayed over the whole variable and the whole
variable moves, i.e. a dense Adam step with a zero-filled gradient.  torch: memory only.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .feat_models import FeatSpec, FeatTablesStruct, _dev

BN_EPS = 1e-3          # tf.layers.batch_normalization defaults
BN_MOMENTUM = 0.99
BETA1, BETA2 = 0.9, 0.999

_TABLES = ("user_embeds", "item_embeds", "sparse_embeds", "dense_embeds",
           "user_linear", "item_linear", "sparse_linear", "dense_linear")


class FMTrainer:
    """Owns the FM variables of ``fm.py`` (scope "embedding" tables + the two Dense(1) heads + BN),
    their Adam slots and gradient buffers.  ``weights`` uses the inference layout of
    ``feat_models.FM`` / ``oracle.tf_models.make_fm_weights``."""

    def __init__(self, spec, weights, use_bn=True, lr=1e-3, epsilon=1e-5, device=None):
        import torch

        self._torch = torch
        K = int(weights["user_embeds"].shape[1])
        self.spec = spec if isinstance(spec, FeatSpec) else FeatSpec(spec, K, device)
        self.device, self.K = self.spec.device, K
        self.use_bn, self.lr, self.epsilon, self.t = bool(use_bn), float(lr), float(epsilon), 0
        f32 = torch.float32
        p = {k: _dev(weights[k], self.device, f32).clone() for k in _TABLES if weights.get(k) is not None}
        p["lin_kernel"] = _dev(np.asarray(weights["lin_kernel"]).reshape(-1), self.device, f32).clone()
        p["lin_bias"] = _dev(np.asarray(weights["lin_bias"]).reshape(1), self.device, f32).clone()
        p["pw_kernel"] = _dev(np.asarray(weights["pw_kernel"]).reshape(-1), self.device, f32).clone()
        p["pw_bias"] = _dev(np.asarray(weights["pw_bias"]).reshape(1), self.device, f32).clone()
        if self.use_bn:
            bn = weights.get("fm_bn")
            one, zero = np.ones(K, np.float32), np.zeros(K, np.float32)
            p["bn_gamma"] = _dev(bn["gamma"] if bn else one, self.device, f32).clone()
            p["bn_beta"] = _dev(bn["beta"] if bn else zero, self.device, f32).clone()
            self.moving_mean = _dev(bn["mean"] if bn else zero, self.device, f32).clone()
            self.moving_var = _dev(bn["var"] if bn else one, self.device, f32).clone()
        self.params = p
        self.grads = {k: torch.zeros_like(v) for k, v in p.items()}
        self.m = {k: torch.zeros_like(v) for k, v in p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in p.items()}
        T = FeatTablesStruct()
        for k in _TABLES:
            setattr(T, k, p[k].data_ptr() if k in p else None)
        self.tables = T
        self._buf = {}

    def _buffers(self, R):
        torch = self._torch
        if self._buf.get("R") != R:
            K, dev, f32 = self.K, self.device, torch.float32
            nb = int(_lib.lib.b200_fm_head_backward_workspace_bytes(R, K))
            self._buf = dict(
                R=R, S=torch.empty((R, K), dtype=f32, device=dev), Q=torch.empty((R, K), dtype=f32, device=dev),
                pw=torch.empty((R, K), dtype=f32, device=dev), y=torch.empty((R, K), dtype=f32, device=dev),
                dpw=torch.empty((R, K), dtype=f32, device=dev), lin=torch.empty(R, dtype=f32, device=dev),
                z=torch.empty(R, dtype=f32, device=dev), logit=torch.empty(R, dtype=f32, device=dev),
                dlogit=torch.empty(R, dtype=f32, device=dev), loss=torch.empty((), dtype=f32, device=dev),
                mean=torch.empty(K, dtype=f32, device=dev), var=torch.empty(K, dtype=f32, device=dev),
                ws=torch.empty(nb, dtype=torch.uint8, device=dev),
                lws=torch.empty(int(_lib.lib.b200_loss_workspace_bytes()), dtype=torch.uint8, device=dev))
        return self._buf

    def forward(self, users_d, items_d):
        """Training-mode logits of the batch (batch statistics in the BN); fills the step buffers."""
        lib, st, p, K = _lib.lib, _lib.current_stream(), self.params, self.K
        R = int(users_d.numel())
        b = self._buffers(R)
        L = self.spec.layout
        _lib.check(lib.b200_feat_forward(
            ctypes.byref(L), ctypes.byref(self.tables), _lib.ptr(users_d), _lib.ptr(items_d), R, 0, 0,
            None, 0, _lib.ptr(b["pw"]), K, _lib.ptr(b["lin"]), None, _lib.ptr(p["lin_kernel"]), 0.0,
            None, None, None, 0.0, _lib.ptr(b["S"]), _lib.ptr(b["Q"]), K, st))
        y = b["pw"]
        if self.use_bn:
            _lib.check(lib.b200_bn_train_forward(
                _lib.ptr(b["pw"]), K, R, K, _lib.ptr(p["bn_gamma"]), _lib.ptr(p["bn_beta"]), BN_EPS, BN_MOMENTUM,
                _lib.ptr(b["y"]), K, _lib.ptr(b["mean"]), _lib.ptr(b["var"]), _lib.ptr(self.moving_mean),
                _lib.ptr(self.moving_var), st))
            y = b["y"]
        _lib.check(lib.b200_fm_head_forward(_lib.ptr(y), K, R, K, _lib.ptr(p["pw_kernel"]), _lib.ptr(p["pw_bias"]),
                                            _lib.ptr(b["lin"]), _lib.ptr(p["lin_bias"]), _lib.ptr(b["z"]),
                                            _lib.ptr(b["logit"]), st))
        return b["logit"]

    def step(self, users_d, items_d, labels_d):
        """One optimisation step on (users, items, labels) device tensors; returns the device loss."""
        torch = self._torch
        lib, st, p, g, K = _lib.lib, _lib.current_stream(), self.params, self.grads, self.K
        users_d = users_d.to(torch.int64).contiguous()
        items_d = items_d.to(torch.int64).contiguous()
        labels_d = labels_d.to(torch.float32).contiguous()
        R = int(users_d.numel())
        self.forward(users_d, items_d)
        b = self._buf
        _lib.check(lib.b200_pointwise_loss(_lib.ptr(b["logit"]), _lib.ptr(labels_d), R, 0, 0.25, 2.0, _lib.ptr(b["loss"]),
                                           _lib.ptr(b["dlogit"]), _lib.ptr(b["lws"]), b["lws"].numel(), st))
        bn = self.use_bn
        _lib.check(lib.b200_fm_head_backward(
            _lib.ptr(b["dlogit"]), _lib.ptr(b["z"]), _lib.ptr(b["pw"]), K, R, K,
            _lib.ptr(b["mean"]) if bn else None, _lib.ptr(b["var"]) if bn else None,
            _lib.ptr(p["bn_gamma"]) if bn else None, _lib.ptr(p["bn_beta"]) if bn else None, BN_EPS,
            _lib.ptr(p["pw_kernel"]), _lib.ptr(b["dpw"]), K, _lib.ptr(g["pw_kernel"]), _lib.ptr(g["pw_bias"]),
            _lib.ptr(g["bn_gamma"]) if bn else None, _lib.ptr(g["bn_beta"]) if bn else None,
            _lib.ptr(g["lin_bias"]), _lib.ptr(b["ws"]), b["ws"].numel(), st))
        gp = lambda k: _lib.ptr(g[k]) if k in g else None      # noqa: E731
        _lib.check(lib.b200_feat_backward(
            ctypes.byref(self.spec.layout), ctypes.byref(self.tables), _lib.ptr(users_d), _lib.ptr(items_d), R,
            _lib.ptr(b["dpw"]), K, _lib.ptr(b["S"]), K, None, 0, _lib.ptr(b["dlogit"]), _lib.ptr(p["lin_kernel"]),
            gp("user_embeds"), gp("item_embeds"), gp("sparse_embeds"), gp("dense_embeds"), gp("user_linear"),
            gp("item_linear"), gp("sparse_linear"), gp("dense_linear"), gp("lin_kernel"), st))
        self.t += 1
        for k in p:
            _lib.check(lib.b200_adam_dense(_lib.ptr(p[k]), _lib.ptr(self.m[k]), _lib.ptr(self.v[k]), _lib.ptr(g[k]),
                                           p[k].numel(), self.lr, BETA1, BETA2, self.epsilon, self.t, st))
        return b["loss"]

    def export_weights(self):
        """Inference weight dict (feat_models.FM layout) with the BN moving statistics."""
        p = self.params
        w = {k: p[k].cpu().numpy() for k in _TABLES if k in p}
        w.update(lin_kernel=p["lin_kernel"].cpu().numpy(), lin_bias=np.float32(p["lin_bias"].cpu().numpy()[0]),
                 pw_kernel=p["pw_kernel"].cpu().numpy(), pw_bias=np.float32(p["pw_bias"].cpu().numpy()[0]))
        if self.use_bn:
            w["fm_bn"] = dict(gamma=p["bn_gamma"].cpu().numpy(), beta=p["bn_beta"].cpu().numpy(),
                              mean=self.moving_mean.cpu().numpy(), var=self.moving_var.cpu().numpy())
        return w
