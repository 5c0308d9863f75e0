"""GPU parity of the device DIN training step (librecommender_b200/training.py::DINTrainer, incl. the attention
backward kernel b200_din_attention_backward) against oracle/din_train.py (torch float64 autograd): attention output,
logits, loss and every raw gradient of one batch (tables incl. the part that arrives through the item feature table,
attention weights), parameters after one step, exported weights in the inference model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed, use_bn, hidden=(64, 32), R=1024, K=16, T=10, n_users=200, n_items=300, n_is=(11, 5, 40), n_id=1):
    from oracle import tf_models as tm

    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, n_users, n_items, [7, 30], list(n_is), 1, n_id)
    w = tm.make_seq_weights(rng, spec, K, hidden, use_bn, din=True)
    batches = []
    for _ in range(2):
        users, items = rng.integers(0, n_users, R), rng.integers(0, n_items, R)
        lens = rng.integers(1, T + 1, R)
        seqs = np.full((R, T), n_items, dtype=np.int32)
        for r in range(R):
            seqs[r, :lens[r]] = rng.integers(0, n_items, lens[r])
        batches.append((users, items, seqs, lens.astype(np.int32), (rng.random(R) < 0.35).astype(np.float32)))
    return spec, w, batches


def _name(k):
    if k.startswith("W") and k[1:].isdigit():
        return "Wt" + k[1:], True
    return k, False


@pytest.mark.parametrize("use_bn,hidden,n_is,n_id", [(True, (64, 32), (11, 5, 40), 1), (False, (48,), (9,), 0),
                                                     (True, (128, 64, 32), (6, 13, 21), 2)])
def test_gradients_of_one_batch_match_oracle(use_bn, hidden, n_is, n_id):
    import torch

    from librecommender_b200.training import DINTrainer
    from oracle import din_train as dt_
    from oracle import tf_models as tm

    spec, w, batches = _case(5, use_bn, hidden, n_is=n_is, n_id=n_id)
    users, items, seqs, lens, labels = batches[0]
    tr = DINTrainer(spec, w, use_bn=use_bn)
    st = dt_.init_state(w, use_bn)
    sparse, dense = tm.row_features(spec, users, items)
    ref_loss, ref_out, ref_g, _, ref_att = dt_.forward_backward(st, spec, users, items, seqs, lens, sparse, dense, labels)
    cu = lambda a: torch.as_tensor(a).cuda()      # noqa: E731
    logits = tr.forward(cu(users), cu(items), cu(seqs), cu(lens))
    np.testing.assert_allclose(logits.cpu().numpy(), ref_out, rtol=3e-5, atol=3e-5)
    loss = tr.backward(cu(labels))
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss) < 2e-5
    gmax = max(np.abs(v).max() for v in ref_g.values())
    for k, ref in ref_g.items():
        name, tr_ = _name(k)
        got = tr.grads[name].cpu().numpy().astype(np.float64)
        got = (got.T if tr_ else got).reshape(ref.shape)
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-3 * scale + 2e-5 * gmax, (k, float(np.abs(got - ref).max()), scale, gmax)


def test_training_steps_match_oracle_and_export():
    import torch

    from librecommender_b200.feat_models import DIN
    from librecommender_b200.training import DINTrainer
    from oracle import din_train as dt_
    from oracle import tf_models as tm

    spec, w, batches = _case(11, True)
    lr, eps = 1e-2, 1e-5
    tr = DINTrainer(spec, w, use_bn=True, lr=lr, epsilon=eps)
    st = dt_.init_state(w, True)
    cu = lambda a: torch.as_tensor(a).cuda()      # noqa: E731
    for step, (users, items, seqs, lens, labels) in enumerate(batches):
        sparse, dense = tm.row_features(spec, users, items)
        ref_loss = dt_.train_step(st, spec, users, items, seqs, lens, sparse, dense, labels, lr, eps)
        loss = tr.step(cu(users), cu(items), cu(seqs), cu(lens), cu(labels))
        assert abs(float(loss) - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss)) * (step + 1), (step, float(loss), ref_loss)
        if step == 0:
            for k, ref in st["params"].items():
                name, tr_ = _name(k)
                got = tr.params[name].cpu().numpy().astype(np.float64)
                got = (got.T if tr_ else got).reshape(ref.shape)
                assert np.abs(got - ref).max() <= 2e-2 * lr + 1e-6, (k, float(np.abs(got - ref).max()))
    w2 = tr.export_weights()
    n_users, n_items, T = 200, 300, 10
    rng = np.random.default_rng(3)
    lens_u = rng.integers(1, T + 1, n_users + 1).astype(np.int32)
    seqs_u = np.full((n_users + 1, T), n_items, dtype=np.int32)
    for u in range(n_users + 1):
        seqs_u[u, :lens_u[u]] = rng.integers(0, n_items, lens_u[u])
    model = DIN(spec, w2, seqs_u, lens_u)
    users, items = rng.integers(0, n_users, 300), rng.integers(0, n_items, 300)
    got = model.logits(users, items).cpu().numpy()
    sparse, dense = tm.row_features(spec, users, items)
    ref = tm.din_forward(w2, spec, users, items, seqs_u[users], lens_u[users], sparse, dense, dtype=np.float64)
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
