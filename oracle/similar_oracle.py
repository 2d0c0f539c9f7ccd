"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's similarity-search helpers, torch CPU fp32:

  * ``compute_distance``            tools/find_similar_imgs.py:88-94
  * ``TopNOracle``                  tools/find_similar_imgs.py:67-85 (the ``topN`` class: fill, then replace the
                                    current worst when strictly better)
  * ``nearest_oracle``              the per-sample loop of find_similar_imgs (:96-137) over an in-memory [N,E] array
  * ``diversity_order_oracle``      _3_label_images.py:128-177 — greedy farthest-point ordering with Python's
                                    ``random.sample`` stream

Pinned by tests/golden/similar_ref.npz, produced by tools/gen_golden.py from the UNMODIFIED reference functions
(imported from /root/reference with tkinter/natsort stubbed) on synthetic ``.pt`` directories.
"""
from __future__ import annotations

import random

import numpy as np
import torch


def compute_distance(context, sample, similarity_measure):
    """tools/find_similar_imgs.py:88-94."""
    if similarity_measure == "cosine":
        return (1 - torch.nn.functional.cosine_similarity(context, sample, dim=-1)) / 2
    if similarity_measure == "l2":
        return torch.nn.functional.pairwise_distance(context, sample, p=2, eps=1e-06)
    raise NotImplementedError(similarity_measure)


class TopNOracle:
    """tools/find_similar_imgs.py:67-85."""

    def __init__(self, top_n):
        self.top_n = top_n
        self.best_ids = []
        self.best_distances = []

    def update(self, distance, ident):
        if len(self.best_distances) < self.top_n:
            self.best_ids.append(ident)
            self.best_distances.append(distance)
        else:
            idx = int(torch.tensor(self.best_distances).argmax().item())
            if distance < self.best_distances[idx]:
                self.best_ids[idx] = ident
                self.best_distances[idx] = distance


def nearest_oracle(context: torch.Tensor, emb: torch.Tensor, top_n: int, measure: str, skip=None):
    """Rows of ``emb`` visited in order (the reference visits files in os.walk order); returns (indices, distances)
    sorted ascending by (distance, index) — the reference's list is in replacement order, the SET is what matters."""
    top = TopNOracle(top_n)
    for i in range(emb.shape[0]):
        if skip is not None and skip[i]:
            continue
        top.update(float(compute_distance(context, emb[i], measure)), i)
    order = sorted(range(len(top.best_ids)), key=lambda t: (top.best_distances[t], top.best_ids[t]))
    return [top.best_ids[t] for t in order], [top.best_distances[t] for t in order]


def cosine_similarity_matrix(a, b):
    """_3_label_images.py:129-132."""
    a_norm = a / a.norm(dim=1, keepdim=True)
    b_norm = b / b.norm(dim=1, keepdim=True)
    return torch.matmul(a_norm, b_norm.t())


def draw_samples(n_items: int, steps: int, sample_size: int, seed):
    """The index stream ``random.sample(image_files, sample_size)`` produces (_3_label_images.py:147): sampling a list
    of n items draws the same positions as sampling range(n) with the same generator state."""
    rng = random.Random(seed)
    return [rng.sample(range(n_items), sample_size) for _ in range(steps)]


def diversity_order_oracle(emb: torch.Tensor, samples, total_n_ordered_imgs=None):
    """_3_label_images.py:135-177 on an in-memory [N,E] array with the sample positions given (one list per step).
    Returns the selected row indices, first row first (duplicates possible, like the reference's list)."""
    order = [0]
    cur = emb[0:1].float()
    steps = len(samples) if total_n_ordered_imgs is None else min(total_n_ordered_imgs, len(samples))
    for s in range(steps):
        idx = list(samples[s])
        se = emb[idx].float()
        sims = cosine_similarity_matrix(cur, se)
        max_val, _ = torch.max(sims, dim=0)
        pick = idx[int(torch.argmin(max_val).item())]
        order.append(pick)
        cur = torch.cat((cur, emb[pick:pick + 1].float()), dim=0)
    return order


def synthetic_clusters(n, E, seed, n_clusters=12, spread=0.35):
    """Unit-norm embeddings drawn around a few cluster centres, so nearest / farthest structure is non-trivial."""
    g = np.random.default_rng(seed)
    centres = g.standard_normal((n_clusters, E)).astype(np.float32)
    which = g.integers(0, n_clusters, n)
    x = centres[which] + spread * g.standard_normal((n, E)).astype(np.float32) * np.sqrt(E).astype(np.float32) / np.float32(np.sqrt(E))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)
