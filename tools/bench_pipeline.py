#!/usr/bin/env python
"""Files -> embeddings, the whole drop-in path of _1_embed_with_CLIP.py on one GPU: a directory of N synthetic 512x512
JPEG files -> DataLoader workers (Huffman stage of K14, or Pillow) -> device decode (K14) -> 4 crops + resize (K0) ->
ViT-L/14 (K1-K8) -> image statistics (K13) -> per-image .pt files and/or the packed store.  Wall-clock images/s of
Feature_Dataset.process(), second pass (page cache warm, weights resident).
    python tools/bench_pipeline.py [--n 8192] [--workers 16] [--batch 256]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_pipeline.py --images 4096 --only "packed store"
One JSON line per configuration.  Under torchrun every rank runs the pipeline on its OWN directory of --n files with its
share of the host cores as DataLoader workers (ranks start together after a file barrier; the box-wide throughput is
the sum of the ranks' lines): what the 16 cores of an 8-GPU box can feed."""
import argparse
import contextlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--images", dest="n", type=int, default=8192)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--model", default="ViT-L-14/openai")
    ap.add_argument("--progressive-every", type=int, default=4, help="every k-th file is a progressive JPEG (0 = none)")
    ap.add_argument("--only", default="", help="substring filter on the configuration names")
    a = ap.parse_args()
    world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if world > 1:
        torch.cuda.set_device(local)
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
        if a.workers == os.cpu_count():
            a.workers = per
    from PIL import Image
    from bench import synth_batch
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder

    root = tempfile.mkdtemp(prefix="b2c_pipe_")
    try:
        import concurrent.futures as cf
        imgs = synth_batch(64, 0).numpy()

        def write(i):
            Image.fromarray(imgs[i % 64]).save(os.path.join(root, f"{i:06d}.jpg"), quality=90, subsampling=2,
                                               progressive=(a.progressive_every > 0 and i % a.progressive_every == a.progressive_every - 1))
        with cf.ThreadPoolExecutor(os.cpu_count()) as ex:
            list(ex.map(write, range(a.n)))
        # ---- the host side alone: how many images per second the DataLoader workers can hand to the main process (what
        # bounds files -> embeddings once several GPUs share the box's cores), and the same plus the device decode
        from torch.utils.data import DataLoader
        from clip_assisted_data_labeling_b200.embed_driver import find_images
        from clip_assisted_data_labeling_b200.embedder import RawImageDataset, collate_raw, to_device_images
        paths = sorted(find_images(root))
        for name, kw in (("loader only: Pillow decode in the workers", dict(device_jpeg=False)),
                         ("loader only: host Huffman stage in the workers (packed coefficients)", dict(device_jpeg=True, device_huffman=False)),
                         ("loader only: marker parse in the workers (file bytes; Huffman stage on the device)", dict(device_jpeg=True, device_huffman=True))):
            if a.only and a.only not in name:
                continue
            for with_device in (False, True):
                best = None
                for rep in range(2):
                    dl = DataLoader(RawImageDataset(paths, **kw), batch_size=a.batch, shuffle=False, num_workers=a.workers,
                                    collate_fn=collate_raw, prefetch_factor=2, persistent_workers=False)
                    it = iter(dl)
                    first = next(it)  # worker start-up is not the steady state
                    if with_device:
                        to_device_images(first[0], "cuda")
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    cnt = 0
                    for items, _ in it:
                        if with_device:
                            to_device_images(items, "cuda")
                        cnt += len(items)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                    del it, dl
                print(json.dumps({"config": name + (" + decode to uint8 RGB in HBM" if with_device else ""), "images": cnt, "workers": a.workers,
                                  "batch": a.batch, "seconds": best, "images_per_s": cnt / best}), flush=True)
        if a.only == "loader only":
            return
        if world > 1:  # start together: wait until every rank has written its files
            flag = os.path.join(tempfile.gettempdir(), f"b2c_pipe_ready_{os.environ.get('MASTER_PORT', '0')}")
            os.makedirs(flag, exist_ok=True)
            open(os.path.join(flag, str(local)), "w").close()
            while len(os.listdir(flag)) < world:
                time.sleep(0.05)
        with contextlib.redirect_stdout(sys.stderr):
            enc = CLIP_Encoder(a.model, device="cuda", seed=0, allow_random_init=True)
        os.environ["B2C_DEVICE_HUFFMAN"] = "1"
        for name, kw in (("pillow decode, .pt files", dict(device_jpeg=False)),
                         ("device JPEG decode (K14, Huffman stage on the host), .pt files", dict(device_jpeg=True, _huff="0")),
                         ("device JPEG decode (K14 + K14b: Huffman stage on the device), .pt files", dict(device_jpeg=True)),
                         ("device JPEG decode (K14, Huffman stage on the host), packed store only",
                          dict(device_jpeg=True, write_pt=False, packed_dir=os.path.join(root, "_packed0"), _huff="0")),
                         ("device JPEG decode (K14 + K14b), packed store only", dict(device_jpeg=True, write_pt=False, packed_dir=os.path.join(root, "_packed")))):
            os.environ["B2C_DEVICE_HUFFMAN"] = kw.pop("_huff", "1")
            if a.only and a.only not in name:
                continue
            best = None
            for rep in range(2):
                with contextlib.redirect_stdout(sys.stderr):
                    ds = Feature_Dataset(root, a.model, batch_size=a.batch, num_workers=a.workers, shuffle_filenames=False,
                                         force_reencode=True, encoder=enc, rank=0, world_size=1, **kw)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    n, _ = ds.process()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                assert n == a.n and not ds.failed
                best = dt if best is None else min(best, dt)
            print(json.dumps({"config": name, "model": a.model, "images": a.n, "workers": a.workers, "batch": a.batch,
                              "seconds": best, "images_per_s": a.n / best, "rank": local, "ranks": world}), flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
