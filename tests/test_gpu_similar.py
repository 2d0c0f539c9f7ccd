"""-m gpu: K11 (context scores, exact top-k, greedy diversity ordering) through the C-ABI against the oracle and the
golden vectors the unmodified reference produced.  Distances are fp32 sums in a different
order than torch's: tolerance 2e-6 absolute; index sets must be identical except where two distances are closer
than that tolerance."""
import os
import types

import numpy as np
import pytest
import torch

from oracle.similar_oracle import (compute_distance, diversity_order_oracle, draw_samples, nearest_oracle,
                                   synthetic_clusters)

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return torch.device("cuda")


def _same_selection(got_idx, got_val, ref_idx, ref_val):
    if list(got_idx) == list(ref_idx):
        return True
    # allowed: swaps among near-ties
    extra = set(got_idx) ^ set(ref_idx)
    vals = {**dict(zip(ref_idx, ref_val)), **dict(zip(got_idx, got_val))}
    edge = max(ref_val)
    return all(abs(vals[i] - edge) <= TOL for i in extra) and np.allclose(sorted(got_val), sorted(ref_val), atol=TOL, rtol=0)


@pytest.mark.parametrize("measure", ["l2", "cosine"])
def test_nearest_vs_reference_golden(cuda, lib, golden, measure):
    from clip_assisted_data_labeling_b200.similar import context_scores, nearest
    g = golden("similar_ref.npz")
    n, E, n_ctx, seed = g["sim_meta"].tolist()
    emb = torch.from_numpy(synthetic_clusters(n, E, seed))
    skip = np.asarray([i < n_ctx or i % 17 == 0 for i in range(n)])
    ctx = torch.from_numpy(g[f"sim_{measure}_ctx"])
    d = context_scores(emb.cuda(), ctx, measure, skip).cpu()
    ref = compute_distance(ctx, emb, measure)
    assert torch.isinf(d[skip]).all()
    np.testing.assert_allclose(d[~skip].numpy(), ref[~skip].numpy(), rtol=0, atol=TOL)
    idx, val = nearest(emb.cuda(), ctx, 25, measure, skip)
    assert _same_selection(idx.tolist(), val.tolist(), g[f"sim_{measure}_idx"].tolist(), g[f"sim_{measure}_dist"].tolist())
    np.testing.assert_allclose(val, g[f"sim_{measure}_dist"], rtol=0, atol=TOL)


@pytest.mark.parametrize("n,E,dtype,strided", [(1, 64, torch.float32, False), (1000, 768, torch.float32, True),
                                               (5000, 1024, torch.float16, False), (333, 70, torch.float32, False),
                                               (257, 100, torch.float16, True)])
def test_context_scores_shapes_dtypes_strides(cuda, lib, n, E, dtype, strided):
    """Vector and scalar load paths, f16 input, and a crop column of a packed [N,4,E] block addressed in place."""
    from clip_assisted_data_labeling_b200.similar import context_scores
    g = torch.Generator().manual_seed(n + E)
    if strided:
        block = torch.randn(n, 4, E, generator=g).to(dtype)
        emb = block[:, 1, :]
    else:
        emb = torch.randn(n, E, generator=g).to(dtype)
    ctx = torch.randn(E, generator=g)
    dev = emb.cuda() if not strided else block.cuda()[:, 1, :]
    for measure in ("l2", "cosine", "cosine_sim"):
        got = context_scores(dev, ctx, measure).cpu()
        if measure == "cosine_sim":
            ref = torch.nn.functional.cosine_similarity(ctx, emb.float(), dim=-1)
        else:
            ref = compute_distance(ctx, emb.float(), measure)
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-6, atol=4e-6)


@pytest.mark.parametrize("n,k", [(1, 1), (10, 10), (1000, 30), (4097, 4096), (200_000, 100), (1_000_003, 1000)])
def test_topk_exact_with_ties(cuda, lib, n, k):
    from clip_assisted_data_labeling_b200.similar import topk_smallest
    g = torch.Generator().manual_seed(n)
    s = torch.randn(n, generator=g)
    s[torch.randint(0, n, (max(1, n // 7),), generator=g)] = 0.25  # many exact ties
    if n > 20:
        s[5] = float("inf")
        s[6] = -0.0
        s[7] = 0.0
    idx, val = topk_smallest(s.cuda(), k)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    order = np.lexsort((np.arange(n), (s + 0.0).numpy()))[:k]  # ascending by (value, index)
    assert np.array_equal(idx, order)
    assert np.array_equal(val, s.numpy()[order] + 0.0)


def test_diversity_vs_reference_golden(cuda, lib, golden):
    from clip_assisted_data_labeling_b200.similar import diversity_order
    g = golden("similar_ref.npz")
    n, E, seed, steps, S, rseed = g["div_meta"].tolist()
    emb = torch.from_numpy(synthetic_clusters(n, E, seed, n_clusters=9))
    samples = draw_samples(n, steps, S, rseed)
    got = diversity_order(emb.cuda(), samples)
    assert got.tolist() == g["div_order"][:steps + 1].tolist()


def test_diversity_vs_oracle_larger(cuda, lib):
    from clip_assisted_data_labeling_b200.similar import diversity_order
    n, E, steps, S = 20_000, 768, 120, 100
    emb = torch.from_numpy(synthetic_clusters(n, E, 5, n_clusters=40))
    samples = draw_samples(n, steps, S, 99)
    got = diversity_order(emb.cuda(), samples).tolist()
    ref = diversity_order_oracle(emb, samples)
    if got != ref:  # a different pick is only acceptable on a near-tie of the two smallest maxima
        first = next(i for i, (a, b) in enumerate(zip(got, ref)) if a != b)
        sims = torch.nn.functional.normalize(emb[ref[:first]], dim=1) @ torch.nn.functional.normalize(emb[[got[first], ref[first]]], dim=1).T
        m = sims.max(dim=0).values
        assert abs(float(m[0] - m[1])) < 1e-5, (first, m)


def test_reference_shaped_entry_points(cuda, lib, tmp_path):
    """create_context_embedding / find_similar_imgs / diversity_ordered_image_files over .pt directories."""
    import random
    from clip_assisted_data_labeling_b200 import similar
    n, E, n_ctx = 200, 96, 5
    emb = synthetic_clusters(n, E, 3)
    ctx_dir, search_dir = tmp_path / "ctx", tmp_path / "search"
    ctx_dir.mkdir()
    search_dir.mkdir()
    for i in range(n):
        d = ctx_dir if i < n_ctx else search_dir
        torch.save({"M/x": {"square_padded_crop": torch.from_numpy(emb[i:i + 1].copy())}}, d / f"{i:05d}.pt")
        if i >= n_ctx and i % 11 != 0:
            (d / f"{i:05d}.jpg").write_bytes(b"")
    args = types.SimpleNamespace(clip_models_to_use=["all"], crop_name_to_use="square_padded_crop", similarity_measure="cosine",
                                 top_n=12, search_dir=str(search_dir), output_dir=str(tmp_path))
    ctx, names = similar.create_context_embedding(args, str(ctx_dir))
    assert args.clip_models_to_use == ["M/x"] and len(names) == n_ctx
    top = similar.find_similar_imgs(args, ctx, names)
    skip = np.asarray([i < n_ctx or i % 11 == 0 for i in range(n)])
    ridx, rdist = nearest_oracle(ctx, torch.from_numpy(emb), 12, "cosine", skip)
    assert [int(os.path.basename(p)[:5]) for p in top.best_img_paths] == ridx
    np.testing.assert_allclose(top.best_distances, rdist, rtol=0, atol=TOL)
    files = [str(search_dir / f"{i:05d}.jpg") for i in range(n_ctx, n)]
    random.seed(7)
    ordered = similar.diversity_ordered_image_files(files, str(search_dir), total_n_ordered_imgs=20, sample_size=15)
    samples = draw_samples(len(files), 20, 15, 7)
    ref = diversity_order_oracle(torch.from_numpy(emb[n_ctx:]), samples)
    assert ordered[:21] == [files[i] for i in ref]
    assert sorted(set(ordered)) == sorted(files) and len(ordered) == len(files) + (21 - len(set(ref)))
