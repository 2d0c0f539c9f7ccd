"""ORACLE (test infrastructure, never imported by the product path) — CPU restatement of the duplicate
search core, /root/reference/_2_remove_duplicates.py:63-80:

    E = stack(embeddings)                      (63)   fp16 [N, D]  (cast at :38)
    E = E / ||E||_2                            (67)   in fp16
    S = E @ E.T                                (69)   fp16 out
    idx = where(triu(S, 1) > threshold)        (74)   strict >, i < j, row-major order
    values = [S[i, j].item() ...]              (80)   read back from the fp16 matrix

Pinned by tests/test_oracle_dedup.py against the UNMODIFIED reference functions run on synthetic ``.pt``
directories (fixtures under tests/golden/, generator tools/gen_golden.py)."""
from __future__ import annotations

import numpy as np
import torch


def duplicate_pairs_oracle(embeddings: torch.Tensor, threshold: float, dtype=torch.float16):
    """Returns (pairs int64 [K,2] row-major, sims float32 [K], S float32 [N,N] computed in fp32 for band checks)."""
    e = embeddings.to(dtype)
    e = e / torch.norm(e, dim=1, keepdim=True)
    S = torch.matmul(e, e.T)
    idx = torch.where(torch.triu(S, diagonal=1) > threshold)
    pairs = np.stack([idx[0].numpy(), idx[1].numpy()], axis=1).astype(np.int64).reshape(-1, 2)
    sims = np.array([S[i, j].item() for i, j in pairs.tolist()], dtype=np.float32)
    e32 = torch.nn.functional.normalize(embeddings.float(), dim=1)
    return pairs, sims, (e32 @ e32.T).numpy()


def pair_sets_match(ref_pairs, got_pairs, S32: np.ndarray, threshold: float, band: float = 1e-3):
    """north_star tolerance: identical pair sets except pairs whose similarity lies within ``band`` of the threshold.
    Returns (ok, offending pairs)."""
    a = set(map(tuple, np.asarray(ref_pairs).reshape(-1, 2).tolist()))
    b = set(map(tuple, np.asarray(got_pairs).reshape(-1, 2).tolist()))
    bad = [(i, j) for (i, j) in (a ^ b) if abs(float(S32[i, j]) - threshold) > band]
    return len(bad) == 0, bad


def synthetic_embeddings(n: int, d: int = 768, seed: int = 0, dup_fraction: float = 0.02):
    """unit-norm N(0,I) rows with planted near-duplicates whose target cosine is
    U[0.90, 0.999] (so similarities straddle 0.96).  numpy Generator streams are stable across versions."""
    rng = np.random.default_rng(seed)
    e = rng.standard_normal((n, d)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    k = max(1, int(n * dup_fraction))
    dst = rng.permutation(n)[:k]
    src = rng.integers(0, n, k)
    c = rng.uniform(0.90, 0.999, k).astype(np.float32)
    sigma = np.sqrt(1.0 / c ** 2 - 1.0)
    noise = rng.standard_normal((k, d)).astype(np.float32) / np.sqrt(d)
    v = e[src] + sigma[:, None] * noise
    e[dst] = v / np.linalg.norm(v, axis=1, keepdims=True)
    return torch.from_numpy(e)
