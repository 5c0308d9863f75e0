"""GPU: the device sampler equals its oracle restatement bit-for-bit and satisfies the
reference's sampler invariants (tests/test_collators.py:399-414) and uniformity."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(seed=3, n_users=60, n_items=400, n=257):
    gen = np.random.default_rng(seed)
    consumed = {u: gen.choice(n_items, size=int(gen.integers(1, 80)), replace=False).tolist() for u in range(n_users)}
    users = gen.integers(0, n_users, size=n)
    pos = np.array([consumed[u][0] for u in users])
    return consumed, users, pos, n_users, n_items


@pytest.mark.parametrize("sampler,num_neg", [("random", 1), ("random", 4), ("unconsumed", 1), ("unconsumed", 5),
                                               ("popular", 2)])
def test_device_sampler_bit_exact_vs_oracle(sampler, num_neg):
    import torch
    from librecommender_b200.sampling import DeviceNegativeSampler, MODES
    from oracle import sampling as osm

    consumed, users, pos, n_users, n_items = _data()
    item_consumed = {i: [u for u, its in consumed.items() if i in its] + [0] for i in range(n_items)}
    probs = osm.neg_probs_from_frequency(item_consumed, n_items, 0.75)
    smp = DeviceNegativeSampler(n_items, consumed, n_users, neg_probs=probs, seed=42)
    got = smp.sample(torch.as_tensor(users).cuda(), torch.as_tensor(pos).cuda(), num_neg, sampler, step=7).cpu().numpy()
    cdf = smp.cdf.cpu().numpy()
    ref = osm.device_sampler_reference(users, pos, num_neg, n_items, MODES[sampler], 10, smp.seed, 7,
                                       {u: set(v) for u, v in consumed.items()}, cdf)
    np.testing.assert_array_equal(got, ref)
    osm.check_reference_invariants(got, users, pos, num_neg, n_items,
                                   consumed if sampler == "unconsumed" else None)


def test_device_sampler_statistics_and_steps_differ():
    import torch
    from librecommender_b200.sampling import DeviceNegativeSampler

    n_items = 1000
    smp = DeviceNegativeSampler(n_items, seed=1)
    pos = torch.zeros(200_000, dtype=torch.int64, device="cuda")
    a = smp.sample(None, pos, 1, "random").cpu().numpy()
    b = smp.sample(None, pos, 1, "random").cpu().numpy()
    assert (a != b).mean() > 0.99                      # consecutive steps give different streams
    assert (a != 0).all()                              # never the positive
    counts = np.bincount(a, minlength=n_items)[1:]
    expected = len(a) / (n_items - 1)
    chi2 = ((counts - expected) ** 2 / expected).sum()
    assert chi2 < 1.25 * (n_items - 1)                 # uniform over the other items
