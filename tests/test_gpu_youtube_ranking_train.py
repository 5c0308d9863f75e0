"""GPU parity of the device YouTubeRanking training step (librecommender_b200/training.py::YouTubeRankingTrainer)
against oracle/youtube_ranking_train.py (torch float64 autograd): logits, loss and raw gradients of one batch
(including the sequence gradient scattered into the item table), parameters after 3 steps, exported weights in
the inference model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed, use_bn, hidden=(64, 32), R=1500, K=16, T=12, n_users=300, n_items=500):
    from oracle import tf_models as tm

    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, n_users, n_items, [7, 30, 12], [11, 5, 40], 1, 2)
    w = tm.make_seq_weights(rng, spec, K, hidden, use_bn, din=False)
    batches = []
    for _ in range(3):
        users, items = rng.integers(0, n_users, R), rng.integers(0, n_items, R)
        lens = rng.integers(0, T + 1, R)
        seqs = np.full((R, T), n_items, dtype=np.int32)
        for r in range(R):
            seqs[r, :lens[r]] = rng.integers(0, n_items, lens[r])
        batches.append((users, items, seqs, lens.astype(np.int32), (rng.random(R) < 0.35).astype(np.float32)))
    return spec, w, batches, n_items


def _to_trainer_layout(tr, k, ref):
    """oracle parameter (reference input order) -> trainer parameter (pooled block last, kernels transposed)."""
    if k == "W0":
        return "Wt0", ref[tr.perm].T
    if k.startswith("W"):
        return "Wt" + k[1:], ref.T
    if k.startswith("bn0_"):
        return k, ref[tr.perm]
    return k, ref


@pytest.mark.parametrize("use_bn,hidden", [(True, (64, 32)), (False, (48,)), (True, (128, 64, 32))])
def test_gradients_of_one_batch_match_oracle(use_bn, hidden):
    import torch

    from librecommender_b200.training import YouTubeRankingTrainer
    from oracle import tf_models as tm
    from oracle import youtube_ranking_train as yt

    spec, w, batches, n_items = _case(5, use_bn, hidden)
    users, items, seqs, lens, labels = batches[0]
    tr = YouTubeRankingTrainer(spec, w, use_bn=use_bn)
    st = yt.init_state(w, use_bn)
    sparse, dense = tm.row_features(spec, users, items)
    ref_loss, ref_out, ref_g, _ = yt.forward_backward(st, users, items, seqs, lens, n_items, sparse, dense, labels)
    cu = lambda a: torch.as_tensor(a).cuda()      # noqa: E731
    logits = tr.forward(cu(users), cu(items), cu(seqs), cu(lens))
    np.testing.assert_allclose(logits.cpu().numpy(), ref_out, rtol=3e-5, atol=3e-5)
    loss = tr.backward(cu(labels))
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss) < 2e-5
    gmax = max(np.abs(v).max() for v in ref_g.values())
    for k, ref in ref_g.items():
        name, ref_t = _to_trainer_layout(tr, k, ref)
        got = tr.grads[name].cpu().numpy().astype(np.float64).reshape(ref_t.shape)
        scale = np.abs(ref_t).max()
        assert np.abs(got - ref_t).max() <= 1e-3 * scale + 2e-5 * gmax, (k, float(np.abs(got - ref_t).max()), scale)


@pytest.mark.parametrize("use_bn", [True, False])
def test_training_steps_match_oracle_and_export(use_bn):
    import torch

    from librecommender_b200.feat_models import YouTubeRanking
    from librecommender_b200.training import YouTubeRankingTrainer
    from oracle import tf_models as tm
    from oracle import youtube_ranking_train as yt

    spec, w, batches, n_items = _case(11, use_bn)
    lr, eps = 1e-2, 1e-5
    tr = YouTubeRankingTrainer(spec, w, use_bn=use_bn, lr=lr, epsilon=eps)
    st = yt.init_state(w, use_bn)
    cu = lambda a: torch.as_tensor(a).cuda()      # noqa: E731
    for step, (users, items, seqs, lens, labels) in enumerate(batches):
        sparse, dense = tm.row_features(spec, users, items)
        ref_loss = yt.train_step(st, users, items, seqs, lens, n_items, sparse, dense, labels, lr, eps)
        loss = tr.step(cu(users), cu(items), cu(seqs), cu(lens), cu(labels))
        assert abs(float(loss) - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss)) * (step + 1), (step, float(loss), ref_loss)
        if step == 0:
            for k, ref in st["params"].items():
                name, ref_t = _to_trainer_layout(tr, k, ref)
                got = tr.params[name].cpu().numpy().astype(np.float64).reshape(ref_t.shape)
                assert np.abs(got - ref_t).max() <= 2e-2 * lr + 1e-6, (k, float(np.abs(got - ref_t).max()))
    # exported weights (reference input order again) in the inference model, fed with per-user sequences
    w2 = tr.export_weights()
    n_users = 300
    rng = np.random.default_rng(3)
    T = 12
    lens_u = rng.integers(0, T + 1, n_users + 1).astype(np.int32)
    seqs_u = np.full((n_users + 1, T), n_items, dtype=np.int32)
    for u in range(n_users + 1):
        seqs_u[u, :lens_u[u]] = rng.integers(0, n_items, lens_u[u])
    model = YouTubeRanking(spec, w2, seqs_u, lens_u)
    users, items = rng.integers(0, n_users, 400), rng.integers(0, n_items, 400)
    got = model.logits(users, items).cpu().numpy()
    sparse, dense = tm.row_features(spec, users, items)
    ref = tm.youtube_ranking_forward(w2, users, items, seqs_u[users], lens_u[users], n_items, sparse, dense, dtype=np.float64)
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
