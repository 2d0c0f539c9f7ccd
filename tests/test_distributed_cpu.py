"""world_size-2 gloo run (CPU) of the multi-GPU host logic: shard -> all-gather -> per-rank band ownership ->
variable-length pair exchange -> globally sorted result.  The CUDA pair finder is replaced by the oracle here
(test infrastructure standing in for the kernel); the real kernel is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_pair_fn(gathered, ranges, threshold, compare, capacity):
    e = gathered.float()
    S = e @ e.T
    S16 = S.to(torch.float16)
    pairs, sims = [], []
    thr16 = torch.tensor(threshold, dtype=torch.float16)
    for (r0, r1) in ranges:
        blk = torch.triu(S16, diagonal=1)[r0:r1] > thr16
        ii, jj = torch.where(blk)
        for i, j in zip((ii + r0).tolist(), jj.tolist()):
            pairs.append((i, j))
            sims.append(float(S[i, j]))
    return np.asarray(pairs, np.int64).reshape(-1, 2), np.asarray(sims, np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_assisted_data_labeling_b200 import dedup
    from oracle.dedup_oracle import synthetic_embeddings
    n, d = 600, 64
    e = synthetic_embeddings(n, d, seed=4, dup_fraction=0.05)
    local = e[rank * n // world:(rank + 1) * n // world]
    old = dedup.BAND_ROWS
    pairs, sims = dedup.duplicate_pairs_distributed(local, 0.96, _pair_fn=_oracle_pair_fn)
    # a finer band grid must give the same answer (exercises multi-band ownership on both ranks)
    dedup_bands = dedup.owned_bands(n, rank, world, band_rows=128)
    assert len(dedup_bands) >= 2
    q.put((rank, pairs.tolist(), sims.tolist(), dedup_bands))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dedup_host_logic():
    from oracle.dedup_oracle import duplicate_pairs_oracle, synthetic_embeddings
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    e = synthetic_embeddings(600, 64, seed=4, dup_fraction=0.05)
    ref_pairs, _, S32 = duplicate_pairs_oracle(e, 0.96)
    assert len(ref_pairs) > 3
    for rank, pairs, sims, bands in res:
        assert pairs == sorted(pairs), "result must be in row-major order"
        from oracle.dedup_oracle import pair_sets_match
        ok, bad = pair_sets_match(ref_pairs, pairs, S32, 0.96)
        assert ok, bad
    assert res[0][1] == res[1][1], "every rank returns the same global result"
    all_bands = sorted(res[0][3] + res[1][3])
    assert all_bands[0][0] == 0 and all_bands[-1][1] == 600 and all(a[1] == b[0] for a, b in zip(all_bands, all_bands[1:]))
