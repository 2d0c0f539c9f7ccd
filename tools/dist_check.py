"""Multi-GPU check, launched under torchrun (one rank per GPU, NCCL):
  * duplicate search: shards -> normalise -> all_gather_into_tensor (NCCL) -> per-rank bands -> pair exchange must equal the
    single-GPU result on the concatenated embeddings (identical pair list, identical order);
  * store-backed duplicate search: every rank writes its shard of the same set into a packed store and
    find_near_duplicates_in_store_distributed must report the pairs of the device run, as paths;
  * embedding: each rank encodes its contiguous shard of a fixed image list; gathered, they must equal rank 0 encoding
    the whole list (bit-identical: no cross-image arithmetic on the path).
Prints one JSON line on rank 0; exit code 1 on mismatch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bench import synth_batch
    from clip_assisted_data_labeling_b200.dedup import duplicate_pairs, duplicate_pairs_distributed
    from clip_assisted_data_labeling_b200.embed_driver import shard_for_rank
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder

    n = int(os.environ.get("DIST_CHECK_N", 60000)) // world * world
    g = torch.Generator().manual_seed(11)
    e = torch.nn.functional.normalize(torch.randn(n, 768, generator=g), dim=1)
    k = n // 50
    src = torch.randint(0, n, (k,), generator=g)
    dst = torch.randperm(n, generator=g)[:k]
    c = torch.empty(k).uniform_(0.90, 0.999, generator=g)
    e[dst] = torch.nn.functional.normalize(e[src] + (1 / c ** 2 - 1).sqrt()[:, None] * torch.randn(k, 768, generator=g) / 768 ** 0.5, dim=1)
    e16 = e.to(torch.float16).cuda()
    shard = e16[rank * n // world:(rank + 1) * n // world].contiguous()
    pd, sd = duplicate_pairs_distributed(shard, 0.96)
    ok_dedup = True
    if rank == 0:
        p1, s1 = duplicate_pairs(e16, 0.96)
        ok_dedup = bool(np.array_equal(p1, pd) and np.array_equal(s1, sd)) and len(p1) > 0

    # the same search fed from a packed store on disk, one shard per rank (ragged: the last rank's shard is shorter)
    import shutil
    import tempfile
    from clip_assisted_data_labeling_b200.dedup import find_near_duplicates_in_store_distributed
    from clip_assisted_data_labeling_b200.store import PackedWriter
    tmp = os.path.join(tempfile.gettempdir(), "b2c_dist_check_store")
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
    dist.barrier()
    n_s = n - 7  # rows of the set that go into the store; the last shard is 7 rows short of the others
    per = n // world
    lo, hi = rank * per, min((rank + 1) * per, n_s)
    feats = np.zeros((hi - lo, 4, 768), np.float16)
    feats[:, 1] = e16[lo:hi].cpu().numpy()
    with PackedWriter(tmp, "M/x", 768, ["square_padded_crop"], shard=rank, dtype="float16") as w:
        w.append(feats, [f"/d/{i:08d}.jpg" for i in range(lo, hi)])
    dist.barrier()
    dups, vals = find_near_duplicates_in_store_distributed(tmp, 0.96, model_name="M/x")
    ok_store = True
    if rank == 0:
        want = [(int(i), int(j)) for i, j in pd.tolist() if i < n_s and j < n_s]
        got = [(int(os.path.basename(a)[:8]), int(os.path.basename(b)[:8])) for a, b in dups]
        ok_store = got == want and len(got) > 0 and len(vals) == len(got)
    dist.barrier()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        enc = CLIP_Encoder("ViT-B-32/openai", device="cuda", seed=0, allow_random_init=True)
    imgs = synth_batch(8 * world, 3)
    idx = shard_for_rank(list(range(len(imgs))), rank, world)
    mine = enc.encode_images_u8(imgs[idx].cuda())
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok_embed = True
    if rank == 0:
        full = enc.encode_images_u8(imgs.cuda())
        ok_embed = bool(torch.equal(torch.cat(gathered), full))
        print(json.dumps({"world": world, "dedup_n": n, "dedup_pairs": int(len(pd)), "dedup_identical": ok_dedup,
                          "store_dedup_identical": ok_store, "embed_images": len(imgs), "embed_identical": ok_embed}))
    flag = torch.tensor([int(ok_dedup and ok_embed and ok_store)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
