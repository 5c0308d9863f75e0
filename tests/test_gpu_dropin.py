"""The drop-in seam of SURVEY.md 8b row 1, exercised through the reference's OWN classes: the
unmodified ``libreco.algorithms.LightGCN`` (``fit`` -> ``TorchTrainer`` -> ``set_embeddings`` ->
``assign_embedding_oov`` -> ``default_recs``; ``recommend_user``) runs once on the reference's
torch-CPU / numpy path and once with ``librecommender_b200.dropin.install()`` (CUDA K6 module, CUDA
losses, fused K4 scorer).  Data: C1's ``sample_movielens_rating.dat`` (reference
``examples/pure_ranking_example.py:30-40``).

* identical weights -> ``recommend_user(all users, 7 and 100)`` ids equal outside near-ties, and the
  ``default_recs`` of the OOV user likewise (``bases/embed_base.py:153-161,190-251``);
* same seed, one epoch of the reference's own training loop both ways -> embeddings agree to 1e-4.

The reference comes from ``/root/reference`` (build container) or from the byte-identical staged
copy ``oracle/_ref`` (GPU box; ``oracle/make_ref.py``)."""
import numpy as np
import pytest

from oracle.ref_loader import load_reference, reference_available, sample_data_path

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(), reason="reference neither mounted nor staged")]


def _data():
    import pandas as pd

    load_reference()
    from libreco.data import DatasetPure, split_by_ratio_chrono

    data = pd.read_csv(sample_data_path(), sep="::", names=["user", "item", "label", "time"], engine="python")
    train, _ = split_by_ratio_chrono(data, test_size=0.2)
    train_data, data_info = DatasetPure.build_trainset(train)
    return train_data, data_info


def _make(data_info, device, n_epochs=1):
    from libreco.algorithms import LightGCN

    return LightGCN("ranking", data_info, loss_type="bpr", embed_size=16, n_epochs=n_epochs, lr=1e-3,
                    batch_size=2048, num_neg=1, dropout_rate=0.0, n_layers=3, device=device, seed=42)


def _near_tie_ok(ref_ids, got_ids, U, I, users, n_items, tol=1e-5):
    from oracle import ranking as orc

    full = orc.embed_scores(U, I, users, n_items)
    return orc.near_tie_mask(np.asarray(ref_ids), np.asarray(got_ids), full, tol)


def test_reference_lightgcn_runs_on_the_dropin_and_matches():
    import torch

    from librecommender_b200 import dropin
    from librecommender_b200.engine import invalidate_scorers

    train_data, data_info = _data()
    # ---- unmodified reference, torch CPU + numpy
    ref = _make(data_info, "cpu")
    ref.fit(train_data, neg_sampling=True, verbose=1, shuffle=True)
    users = list(range(data_info.n_users))
    ref_rec7 = ref.recommend_user(users, 7, inner_id=True)
    ref_rec100 = ref.recommend_user(users, 100, inner_id=True)

    # ---- the same classes with the drop-in installed
    import libreco

    dropin.install(libreco)
    try:
        from libreco.algorithms import lightgcn as lg_mod

        assert lg_mod.LightGCNModel.__module__.startswith("librecommender_b200")
        gpu = _make(data_info, "cuda")
        gpu.fit(train_data, neg_sampling=True, verbose=1, shuffle=True)
        assert next(gpu.torch_model.parameters()).is_cuda
        # same seed, same sampler stream, one epoch of the reference's loop: embeddings agree
        np.testing.assert_allclose(gpu.user_embeds_np, ref.user_embeds_np, rtol=0, atol=1e-4)
        np.testing.assert_allclose(gpu.item_embeds_np, ref.item_embeds_np, rtol=0, atol=1e-4)

        # ---- identical weights: load the reference's trained parameters into the drop-in module
        with torch.no_grad():
            sd = {k: v.to("cuda") for k, v in ref.torch_model.state_dict().items()
                  if k in ("user_init_embeds.weight", "item_init_embeds.weight")}
            gpu.torch_model.load_state_dict(sd, strict=False)
        gpu.set_embeddings()                       # reference code, drop-in module underneath
        gpu.assign_embedding_oov()
        invalidate_scorers()
        np.testing.assert_allclose(gpu.user_embeds_np, ref.user_embeds_np, rtol=0, atol=2e-6)
        got7 = gpu.recommend_user(users, 7, inner_id=True)
        got100 = gpu.recommend_user(users, 100, inner_id=True)
        for ref_rec, got in ((ref_rec7, got7), (ref_rec100, got100)):
            r = np.stack([ref_rec[u] for u in users])
            g = np.stack([got[u] for u in users])
            assert g.dtype == r.dtype
            ok = _near_tie_ok(r, g, ref.user_embeds_np, ref.item_embeds_np, users, data_info.n_items)
            assert ok.all()
            assert (r == g).mean() > 0.98
            for u in users[:200]:                  # consumed items never recommended
                assert not set(g[u].tolist()) & set(data_info.user_consumed[u])
        # default_recs (OOV user, top-2000 without filter) through the patched recommend_from_embedding
        from libreco.recommendation import recommend_from_embedding

        dr = recommend_from_embedding(gpu, [gpu.n_users], min(2000, gpu.n_items), gpu.user_embeds_np,
                                      gpu.item_embeds_np, False, False).flatten()
        ok = _near_tie_ok(ref.default_recs[None, :], dr[None, :], ref.user_embeds_np, ref.item_embeds_np,
                          [ref.n_users], data_info.n_items)
        assert ok.all() and (dr == ref.default_recs).mean() > 0.98
        # original-id output + cold start keep working through the reference's own code
        some = [data_info.id2user[u] for u in users[:5]] + ["no-such-user"]
        out = gpu.recommend_user(some, 10)
        assert len(out) == 6 and all(len(v) == 10 for v in out.values())
    finally:
        dropin.uninstall()
    assert not dropin.installed()
