"""oracle/youtube_ranking_train.py self-checks: with BN off the training forward equals the inference
restatement (oracle.tf_models.youtube_ranking_forward), the pad position carries no gradient, steps reduce the loss."""
import numpy as np

from oracle import tf_models as tm
from oracle import youtube_ranking_train as yt


def _case(seed, use_bn, R=120, K=8, T=6):
    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, 40, 60, [5, 9], [7, 4], 1, 1)
    w = tm.make_seq_weights(rng, spec, K, (16, 8), use_bn, din=False)
    users, items = rng.integers(0, 40, R), rng.integers(0, 60, R)
    lens = rng.integers(0, T + 1, R)
    seqs = np.full((R, T), 60, dtype=np.int64)
    for r in range(R):
        seqs[r, :lens[r]] = rng.integers(0, 60, lens[r])
    sparse, dense = tm.row_features(spec, users, items)
    labels = (rng.random(R) < 0.4).astype(np.float32)
    return spec, w, users, items, seqs, lens, sparse, dense, labels


def test_forward_equals_inference_restatement_without_bn():
    spec, w, users, items, seqs, lens, sparse, dense, labels = _case(1, False)
    st = yt.init_state(w, False)
    loss, out, g, _ = yt.forward_backward(st, users, items, seqs, lens, 60, sparse, dense, labels)
    ref = tm.youtube_ranking_forward(w, users, items, seqs, lens, 60, sparse, dense, dtype=np.float64)
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
    # the OOV / pad row of the item table only gets gradient from rows whose ITEM is the pad id (none here)
    assert np.abs(g["item_embeds"][60]).max() == 0.0
    assert np.isfinite(loss)


def test_steps_reduce_loss():
    spec, w, users, items, seqs, lens, sparse, dense, labels = _case(2, True)
    st = yt.init_state(w, True)
    losses = [yt.train_step(st, users, items, seqs, lens, 60, sparse, dense, labels, 1e-2) for _ in range(6)]
    assert losses[-1] < losses[0]
