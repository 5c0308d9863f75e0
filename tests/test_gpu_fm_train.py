"""GPU parity of the device FM training step (librecommender_b200/training.py, csrc/train.cu)
against oracle/fm_train.py (numpy float64; gradient math pinned to torch autograd by
tests/test_fm_train_cpu.py; TensorFlow conventions unpinned): per-variable parameters after 1 and 4
steps, loss values, BN moving statistics, and inference parity of the exported weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed, use_bn, n_users=300, n_items=500, R=2048, K=16):
    from oracle import tf_models as tm

    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, n_users, n_items, [7, 30], [11, 5, 40], 1, 2)
    w = tm.make_fm_weights(rng, spec, K, use_bn)
    batches = []
    for _ in range(4):
        users = rng.integers(0, n_users, R)
        items = rng.integers(0, n_items, R)
        labels = (rng.random(R) < 0.35).astype(np.float32)
        batches.append((users, items, labels))
    return spec, w, batches


@pytest.mark.parametrize("use_bn", [True, False])
def test_fm_training_steps_match_oracle(use_bn):
    import torch

    from librecommender_b200.training import FMTrainer
    from oracle import fm_train as ft
    from oracle import tf_models as tm

    spec, w, batches = _case(17, use_bn)
    lr, eps = 1e-2, 1e-5
    tr = FMTrainer(spec, w, use_bn=use_bn, lr=lr, epsilon=eps)
    st = ft.init_state(w, use_bn, dtype=np.float64)
    for step, (users, items, labels) in enumerate(batches):
        sparse, dense = tm.row_features(spec, users, items)
        ref_loss = ft.train_step(st, users, items, sparse, dense, labels, lr, eps)
        loss = tr.step(torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda(), torch.as_tensor(labels).cuda())
        assert abs(float(loss) - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (step, float(loss), ref_loss)
        if step in (0, 3):
            for k, ref in st["params"].items():
                got = tr.params[k].cpu().numpy().astype(np.float64).reshape(ref.shape)
                # Adam normalises the update to ~lr per element: compare on that scale
                err = np.abs(got - ref).max()
                assert err <= 3e-3 * lr * (step + 1) + 1e-6, (step, k, err)
    if use_bn:
        np.testing.assert_allclose(tr.moving_mean.cpu().numpy(), st["moving_mean"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(tr.moving_var.cpu().numpy(), st["moving_var"], rtol=1e-4, atol=1e-6)
    # exported weights drive the inference engine to the oracle's inference logits
    from librecommender_b200.feat_models import FM

    w_dev, w_ref = tr.export_weights(), ft.export_weights(st)
    users, items, _ = batches[0]
    sparse, dense = tm.row_features(spec, users, items)
    ref = tm.fm_forward(w_ref, users, items, sparse, dense, dtype=np.float64)
    got = FM(spec, w_dev).logits(users, items).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


def test_gradients_of_one_batch_match_oracle():
    """The raw gradient buffers before the optimiser touches them (Adam zeroes them afterwards)."""
    import ctypes

    import torch

    from librecommender_b200 import _lib
    from librecommender_b200.training import BN_EPS, FMTrainer
    from oracle import fm_train as ft
    from oracle import tf_models as tm

    spec, w, batches = _case(23, True, R=1500)
    users, items, labels = batches[0]
    tr = FMTrainer(spec, w, use_bn=True)
    st = ft.init_state(w, True)
    sparse, dense = tm.row_features(spec, users, items)
    ref_loss, ref_out, ref_g, _ = ft.forward_backward(st["params"], True, users, items, sparse, dense, labels)
    u, i, y = torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda(), torch.as_tensor(labels).cuda()
    logits = tr.forward(u, i)
    np.testing.assert_allclose(logits.cpu().numpy(), ref_out, rtol=2e-5, atol=2e-5)
    b, p, g, K, R = tr._buf, tr.params, tr.grads, tr.K, len(users)
    lib, stm = _lib.lib, _lib.current_stream()
    _lib.check(lib.b200_pointwise_loss(_lib.ptr(b["logit"]), _lib.ptr(y), R, 0, 0.25, 2.0, _lib.ptr(b["loss"]),
                                       _lib.ptr(b["dlogit"]), _lib.ptr(b["lws"]), b["lws"].numel(), stm))
    _lib.check(lib.b200_fm_head_backward(
        _lib.ptr(b["dlogit"]), _lib.ptr(b["z"]), _lib.ptr(b["pw"]), K, R, K, _lib.ptr(b["mean"]), _lib.ptr(b["var"]),
        _lib.ptr(p["bn_gamma"]), _lib.ptr(p["bn_beta"]), BN_EPS, _lib.ptr(p["pw_kernel"]), _lib.ptr(b["dpw"]), K,
        _lib.ptr(g["pw_kernel"]), _lib.ptr(g["pw_bias"]), _lib.ptr(g["bn_gamma"]), _lib.ptr(g["bn_beta"]),
        _lib.ptr(g["lin_bias"]), _lib.ptr(b["ws"]), b["ws"].numel(), stm))
    _lib.check(lib.b200_feat_backward(
        ctypes.byref(tr.spec.layout), ctypes.byref(tr.tables), _lib.ptr(u), _lib.ptr(i), R, _lib.ptr(b["dpw"]), K,
        _lib.ptr(b["S"]), K, None, 0, _lib.ptr(b["dlogit"]), _lib.ptr(p["lin_kernel"]), _lib.ptr(g["user_embeds"]),
        _lib.ptr(g["item_embeds"]), _lib.ptr(g["sparse_embeds"]), _lib.ptr(g["dense_embeds"]),
        _lib.ptr(g["user_linear"]), _lib.ptr(g["item_linear"]), _lib.ptr(g["sparse_linear"]),
        _lib.ptr(g["dense_linear"]), _lib.ptr(g["lin_kernel"]), stm))
    torch.cuda.synchronize()
    assert abs(float(b["loss"]) - ref_loss) < 1e-5
    for k, ref in ref_g.items():
        got = g[k].cpu().numpy().astype(np.float64).reshape(ref.shape)
        scale = max(np.abs(ref).max(), 1e-8)
        assert np.abs(got - ref).max() <= 2e-4 * scale + 1e-9, (k, float(np.abs(got - ref).max()), scale)
