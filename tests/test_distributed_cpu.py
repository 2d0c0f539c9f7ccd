"""world_size-2 gloo run (CPU) of the multi-GPU host logic: shard -> all-gather -> per-rank band ownership ->
fixed-capacity pair exchange (counts + slots, re-run on overflow) -> globally sorted result.  The two CUDA kernel calls
are replaced by CPU stand-ins inside the TEST process (monkeypatched; the product function carries no hook and refuses
host tensors); the real kernels are covered by the -m gpu tests and tools/dist_check.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_normalize(emb, out=None):
    """Stands in for b2c_normalize_rows_f16 (the CUDA kernel) in this CPU-only host-logic test."""
    n = torch.nn.functional.normalize(emb.float(), dim=1).to(torch.float16)
    E_pad = (emb.shape[1] + 63) // 64 * 64
    if out is None:
        out = torch.zeros(emb.shape[0], E_pad, dtype=torch.float16)
    out.zero_()
    out[:, :emb.shape[1]] = n
    return out


def _oracle_search(gathered, ranges, threshold, compare, buf, cnt):
    """Stands in for b2c_dedup_pairs: same contract (append b2c_pair rows to buf, count every hit even past capacity)."""
    e = gathered.float()
    S = e @ e.T
    S16 = S.to(torch.float16)
    thr16 = torch.tensor(threshold, dtype=torch.float16)
    k = int(cnt[0])
    for blk_range in ranges:
        r0, r1, c0, c1 = blk_range if len(blk_range) == 4 else (*blk_range, 0, gathered.shape[0])
        blk = torch.triu(S16, diagonal=1)[r0:r1] > thr16
        blk[:, :c0] = False
        blk[:, c1:] = False
        ii, jj = torch.where(blk)
        for i, j in zip((ii + r0).tolist(), jj.tolist()):
            if k < buf.shape[0]:
                buf[k, 0], buf[k, 1] = i, j
                buf[k, 2] = int(np.float32(S[i, j]).view(np.int32))
            k += 1
    cnt[0] = k


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_assisted_data_labeling_b200 import _lib, dedup
    from oracle.dedup_oracle import synthetic_embeddings
    n, d = 600, 64
    e = synthetic_embeddings(n, d, seed=4, dup_fraction=0.05)
    local = e[rank * n // world:(rank + 1) * n // world]
    # the product function has no CPU path: it refuses host tensors
    try:
        dedup.duplicate_pairs_distributed(local, 0.96)
        raise AssertionError("expected B2CError for CPU tensors")
    except _lib.B2CError:
        pass
    dist.barrier()
    dedup.normalize_rows_f16 = _cpu_normalize      # the two kernel calls, replaced by their CPU stand-ins (test only)
    dedup.launch_pair_search = _oracle_search
    # bands are drawn from the job's store counter (whoever is faster takes more): the result must not depend on who
    # computes what.  The stream marker of the pacing loop has no CPU form.
    class _NoEvent:
        def synchronize(self):
            pass
    dedup.launch_marker = _NoEvent
    pairs, sims = dedup.duplicate_pairs_distributed(local, 0.96)
    # too small a pair buffer: every rank re-runs with room for the largest list and the answer is the same
    pairs2, sims2 = dedup.duplicate_pairs_distributed(local, 0.96, capacity=2)
    assert pairs2.tolist() == pairs.tolist() and np.array_equal(sims2, sims)
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
