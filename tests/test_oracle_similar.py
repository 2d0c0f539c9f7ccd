"""Pins oracle/similar_oracle.py against golden vectors produced by the UNMODIFIED reference functions
(tools/find_similar_imgs.py, _3_label_images.diversity_ordered_image_files; tools/gen_golden.py gen_similar)."""
import numpy as np
import torch

from oracle.similar_oracle import (diversity_order_oracle, draw_samples, nearest_oracle, synthetic_clusters)


def _similar_case(g):
    n, E, n_ctx, seed = g["sim_meta"].tolist()
    emb = synthetic_clusters(n, E, seed)
    # context rows live in another directory; search samples without a .jpg are skipped (find_similar_imgs.py:106-110)
    skip = np.asarray([i < n_ctx or i % 17 == 0 for i in range(n)])
    return torch.from_numpy(emb), skip, n_ctx


def test_nearest_oracle_vs_reference_golden(golden):
    g = golden("similar_ref.npz")
    emb, skip, n_ctx = _similar_case(g)
    ctx = emb[:n_ctx].mean(dim=0)
    for measure in ("l2", "cosine"):
        np.testing.assert_allclose(ctx.numpy(), g[f"sim_{measure}_ctx"], rtol=0, atol=1e-7)
        idx, dist = nearest_oracle(torch.from_numpy(g[f"sim_{measure}_ctx"]), emb, 25, measure, skip)
        assert idx == g[f"sim_{measure}_idx"].tolist(), measure
        np.testing.assert_allclose(np.asarray(dist, np.float32), g[f"sim_{measure}_dist"], rtol=0, atol=1e-7)


def test_topn_oracle_replacement_rule():
    from oracle.similar_oracle import TopNOracle
    t = TopNOracle(2)
    for d, i in [(0.5, 0), (0.3, 1), (0.5, 2), (0.3, 3), (0.1, 4)]:
        t.update(d, i)
    # 0.5(2) does not replace 0.5(0) (strict <); 0.3(3) replaces the worst (0.5); 0.1 replaces the first worst 0.3
    assert sorted(zip(t.best_distances, t.best_ids)) == [(0.1, 4), (0.3, 1)] or sorted(zip(t.best_distances, t.best_ids)) == [(0.1, 4), (0.3, 3)]


def test_diversity_oracle_vs_reference_golden(golden):
    g = golden("similar_ref.npz")
    n, E, seed, steps, S, rseed = g["div_meta"].tolist()
    emb = torch.from_numpy(synthetic_clusters(n, E, seed, n_clusters=9))
    samples = draw_samples(n, steps, S, rseed)
    order = diversity_order_oracle(emb, samples)
    want = g["div_order"]
    assert order == want[:steps + 1].tolist()
    # the tail is every image not selected, in the original order (_3_label_images.py:175)
    chosen = set(order)
    assert want[steps + 1:].tolist() == [i for i in range(n) if i not in chosen]
