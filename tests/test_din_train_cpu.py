"""oracle/din_train.py self-checks: with BN off the training forward equals the inference restatement
(oracle.tf_models.din_forward), the Dense(1) bias of the attention has an exactly zero gradient (softmax shift
invariance — the device trainer relies on it), steps reduce the loss."""
import numpy as np

from oracle import din_train as dt_
from oracle import tf_models as tm


def _case(seed, use_bn, R=90, K=8, T=6):
    rng = np.random.default_rng(seed)
    spec = tm.make_spec(rng, 40, 60, [5, 9], [7, 4], 1, 1)
    w = tm.make_seq_weights(rng, spec, K, (16, 8), use_bn, din=True)
    users, items = rng.integers(0, 40, R), rng.integers(0, 60, R)
    lens = rng.integers(1, T + 1, R)
    seqs = np.full((R, T), 60, dtype=np.int64)
    for r in range(R):
        seqs[r, :lens[r]] = rng.integers(0, 60, lens[r])
    sparse, dense = tm.row_features(spec, users, items)
    labels = (rng.random(R) < 0.4).astype(np.float32)
    return spec, w, users, items, seqs, lens, sparse, dense, labels


def test_forward_equals_inference_restatement_without_bn():
    spec, w, users, items, seqs, lens, sparse, dense, labels = _case(1, False)
    st = dt_.init_state(w, False)
    loss, out, g, _, _ = dt_.forward_backward(st, spec, users, items, seqs, lens, sparse, dense, labels)
    ref = tm.din_forward(w, spec, users, items, seqs, lens, sparse, dense, dtype=np.float64)
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-11)
    assert np.abs(g["att_b2"]).max() < 1e-15            # shift invariance of the softmax
    assert np.abs(g["att_k1"]).max() > 0 and np.isfinite(loss)


def test_steps_reduce_loss():
    spec, w, users, items, seqs, lens, sparse, dense, labels = _case(2, True)
    st = dt_.init_state(w, True)
    losses = [dt_.train_step(st, spec, users, items, seqs, lens, sparse, dense, labels, 1e-2) for _ in range(6)]
    assert losses[-1] < losses[0]
