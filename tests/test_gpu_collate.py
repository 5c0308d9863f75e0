"""GPU: device collators reproduce the batch layout of the reference's collators
(libreco/batch/collators.py:225-252,277-299; pinned by the reference's tests/test_collators.py:59-346:
label pattern 1,0,0…, negatives interleaved after their positive)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host_pointwise(users, items, negs, num_neg):
    # restatement of collators.py:226-232
    u = np.repeat(users, num_neg + 1)
    it = np.repeat(items, num_neg + 1)
    lab = np.zeros(len(it), dtype=np.float32)
    lab[:: num_neg + 1] = 1.0
    for i in range(num_neg):
        it[(i + 1):: num_neg + 1] = negs[i::num_neg]
    return u, it, lab


@pytest.mark.parametrize("num_neg", [1, 3])
def test_pointwise_layout(num_neg):
    import torch
    from librecommender_b200.collate import DevicePointwiseCollator, adjust_batch_size
    from librecommender_b200.sampling import DeviceNegativeSampler

    rng = np.random.default_rng(1)
    users, items = rng.integers(0, 50, 64), rng.integers(0, 300, 64)
    smp = DeviceNegativeSampler(300, seed=7)
    col = DevicePointwiseCollator(smp, num_neg)
    ud, idv = torch.as_tensor(users).cuda(), torch.as_tensor(items).cuda()
    negs = smp.sample(ud, idv, num_neg, "random", step=0)
    u, it, lab = col(ud, idv, negatives_d=negs)
    ru, rit, rlab = _host_pointwise(users, items, negs.cpu().numpy(), num_neg)
    np.testing.assert_array_equal(u.cpu().numpy(), ru)
    np.testing.assert_array_equal(it.cpu().numpy(), rit)
    np.testing.assert_array_equal(lab.cpu().numpy(), rlab)
    assert adjust_batch_size(8192, 5, False) == 1365 and adjust_batch_size(2048, 1, True) == 2048


def test_pairwise_layout():
    import torch
    from librecommender_b200.collate import DevicePairwiseCollator
    from librecommender_b200.sampling import DeviceNegativeSampler

    smp = DeviceNegativeSampler(100, seed=3)
    ud, idv = torch.arange(10).cuda(), torch.arange(10, 20).cuda()
    q, pos, neg = DevicePairwiseCollator(smp, 4)(ud, idv)
    np.testing.assert_array_equal(q.cpu().numpy(), np.repeat(np.arange(10), 4))
    np.testing.assert_array_equal(pos.cpu().numpy(), np.repeat(np.arange(10, 20), 4))
    assert neg.shape == (40,) and (neg.cpu().numpy() != pos.cpu().numpy()).all()
