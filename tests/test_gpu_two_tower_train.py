"""GPU parity of the device TwoTower training step (librecommender_b200/training.py::TwoTowerTrainer) against
oracle/two_tower_train.py (torch float64 autograd): loss and raw gradients of one batch, parameters and BN
moving statistics after 3 steps, exported weights in the inference towers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed, use_bn, hidden=(64, 32), B=1024, K=16, n_users=400, n_items=600):
    from oracle import tf_models as tm

    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, n_users, n_items, [7, 30, 12], [11, 5, 40, 8], 1, 2)
    w = tm.make_two_tower_weights(rng, spec, K, hidden, use_bn)
    batches = []
    for _ in range(3):
        users, items = rng.integers(0, n_users, B), rng.integers(0, n_items, B)
        corr = (np.bincount(items, minlength=n_items) / B)[items].astype(np.float32)
        batches.append((users, items, corr))
    return spec, w, batches


def _feats(spec, users, items):
    return (spec["user_sparse_unique"][users], spec["user_dense_unique"][users],
            spec["item_sparse_unique"][items], spec["item_dense_unique"][items])


def _map(name):
    """oracle name (user_W0 = [din, dout]) -> (trainer name, transpose?)."""
    for which in ("user", "item"):
        if name.startswith(f"{which}_W"):
            return f"{which}_Wt{name[len(which) + 2:]}", True
    return name, False


@pytest.mark.parametrize("use_bn,norm,temp,hits", [(True, False, 1.0, False), (True, True, 0.2, True),
                                                   (False, True, 0.5, False), (False, False, 2.0, True)])
def test_gradients_of_one_batch_match_oracle(use_bn, norm, temp, hits):
    import torch

    from librecommender_b200.training import TwoTowerTrainer
    from oracle import two_tower_train as tt

    spec, w, batches = _case(5, use_bn)
    users, items, corr = batches[0]
    tr = TwoTowerTrainer(spec, w, use_bn=use_bn, norm_embed=norm, temperature=temp, remove_accidental_hits=hits)
    st = tt.init_state(w, use_bn)
    ref_loss, ref_g, _, rU, rV = tt.forward_backward(st, users, items, _feats(spec, users, items), norm=norm,
                                                     temperature=temp, correction=corr, remove_hits=hits)
    u, i, c = torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda(), torch.as_tensor(corr).cuda()
    loss = tr.forward_backward(u, i, c)
    torch.cuda.synchronize()
    U, V = tr._last
    np.testing.assert_allclose(U.cpu().numpy(), rU, rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(V.cpu().numpy(), rV, rtol=3e-5, atol=3e-5)
    assert abs(float(loss) - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    # some gradients are mathematically zero (the softmax is invariant to the item tower's last bias): the
    # absolute term is a fraction of the largest gradient of the whole model (fp32 summation noise)
    gmax = max(np.abs(v).max() for v in ref_g.values())
    for k, ref in ref_g.items():
        name, tr_ = _map(k)
        got = tr.grads[name].cpu().numpy().astype(np.float64)
        got = got.T if tr_ else got
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-3 * scale + 2e-5 * gmax, (k, float(np.abs(got - ref).max()), scale, gmax)


@pytest.mark.parametrize("use_bn", [True, False])
def test_training_steps_match_oracle_and_export(use_bn):
    import torch

    from librecommender_b200.feat_models import TwoTower
    from librecommender_b200.training import TwoTowerTrainer
    from oracle import tf_models as tm
    from oracle import two_tower_train as tt

    spec, w, batches = _case(11, use_bn)
    lr, eps = 1e-2, 1e-5
    tr = TwoTowerTrainer(spec, w, use_bn=use_bn, norm_embed=True, temperature=0.5, lr=lr, epsilon=eps)
    st = tt.init_state(w, use_bn)
    for step, (users, items, corr) in enumerate(batches):
        ref_loss = tt.train_step(st, users, items, _feats(spec, users, items), lr, eps, norm=True, temperature=0.5,
                                 correction=corr)
        loss = tr.step(torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda(), torch.as_tensor(corr).cuda())
        assert abs(float(loss) - ref_loss) <= 2e-3 * max(1.0, abs(ref_loss)) * (step + 1), (step, float(loss), ref_loss)
    # Adam normalises the step size, so parameters that moved agree to a fraction of lr per step
    for k, ref in st["params"].items():
        name, tr_ = _map(k)
        got = tr.params[name].cpu().numpy().astype(np.float64)
        got = got.T if tr_ else got
        assert np.abs(got - ref).max() <= 0.35 * lr * len(batches), (k, float(np.abs(got - ref).max()))
        assert np.median(np.abs(got - ref)) <= 0.02 * lr, (k, float(np.median(np.abs(got - ref))))
    if use_bn:
        for name, (mm, mv) in st["moving"].items():
            np.testing.assert_allclose(tr.moving[name][0].cpu().numpy(), mm, rtol=2e-3, atol=2e-4)
            np.testing.assert_allclose(tr.moving[name][1].cpu().numpy(), mv, rtol=2e-3, atol=2e-4)
    # exported weights drive the inference towers
    w2 = tr.export_weights()
    tower = TwoTower(spec, w2, norm_embed=True)
    ids = np.arange(50)
    got = tower.tower("user", ids).cpu().numpy()
    want = tm.tower_forward(w2, ids, spec["user_sparse_unique"][ids], spec["user_dense_unique"][ids], "user", True,
                            dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-5)


def test_learned_temperature_is_refused():
    from librecommender_b200.training import TwoTowerTrainer

    spec, w, _ = _case(1, False)
    with pytest.raises(ValueError, match="learned temperature"):
        TwoTowerTrainer(spec, w, use_bn=False, temperature=0.0)


def test_graph_replay_equals_eager_steps():
    import torch

    from librecommender_b200.training import TwoTowerTrainer

    spec, w, batches = _case(13, True)
    a = TwoTowerTrainer(spec, w, use_bn=True, norm_embed=True, temperature=0.5, lr=1e-2)
    b = TwoTowerTrainer(spec, w, use_bn=True, norm_embed=True, temperature=0.5, lr=1e-2)
    for users, items, corr in batches + batches:
        u, i, c = torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda(), torch.as_tensor(corr).cuda()
        la, lb = float(a.step(u, i, c)), float(b.step_graph(u, i, c))
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la)), (la, lb)
    assert a.t == b.t == 6
    for k in a.params:
        assert (a.params[k] - b.params[k]).abs().max().item() <= 2e-4, k
