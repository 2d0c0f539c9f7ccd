"""ORACLE-side runner (baseline infrastructure, executed only by bench.py's reference arm / library bar, always in a
SUBPROCESS): time the reference's OWN code for the embedding path —

  loop    ``Feature_Dataset(root, model, batch_size, num_workers=...).process()`` of _1_embed_with_CLIP.py:36-184, imported
          verbatim (oracle/reference_shim.py supplies ``open_clip``: the random-init tower of the named architecture and
          the open_clip val transform), over directories of synthetic 512x512 PNG files: DataLoader workers running
          CustomImageDataset.__getitem__ (PIL crops + transform + ImageFeaturizer), encode_image, the .pt save loop;
  encode  ``CLIP_Encoder(model).encode_image(batch)`` of utils/embedder.py:94-100 on a resident batch.

``--device cpu`` hides the GPUs from the process before torch is imported (the reference picks 'cuda' whenever one is
visible, utils/embedder.py:20, _1:17) — that is the CPU baseline; ``--device cuda`` leaves it alone: the reference in
its native fp16 eager mode on cuBLAS / SDPA library kernels (SURVEY.md §2.1's "honest GPU bar").  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["loop", "encode"], default="loop")
    ap.add_argument("--model", default="ViT-L-14/openai")
    ap.add_argument("--device", choices=["cpu", "cuda"], default="cpu")
    ap.add_argument("--images", type=int, default=8, help="images per step")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--batch", type=int, default=8)      # _1_embed_with_CLIP.py:192
    ap.add_argument("--workers", type=int, default=4)    # _1_embed_with_CLIP.py:193
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if a.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    import contextlib
    import numpy as np
    import torch
    from PIL import Image
    from oracle import reference_shim as rs
    from oracle.preprocess_oracle import synthetic_image
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    out = {"mode": a.mode, "model": a.model, "device": a.device, "cores": cores, "torch_threads": torch.get_num_threads(),
           "reference_root": rs.REFERENCE_ROOT, "cuda_visible": torch.cuda.is_available()}
    if not rs.reference_available():
        print(json.dumps(dict(out, unavailable=f"no reference tree at {rs.REFERENCE_ROOT} (run python -m oracle.make_ref in the build container)")))
        return
    with contextlib.redirect_stdout(sys.stderr):
        if a.mode == "encode":
            emb = rs.import_reference("utils.embedder", seed=a.seed)
            enc = emb.CLIP_Encoder(a.model)
            R = enc.img_resolution
            x = torch.randn(4 * a.images, 3, R, R, device=enc.device)
            sync = torch.cuda.synchronize if str(enc.device).startswith("cuda") else (lambda: None)
            times = []
            with torch.no_grad():
                for s in range(a.warmup + a.steps):
                    sync()
                    t0 = time.perf_counter()
                    y = enc.encode_image(x)
                    sync()
                    if s >= a.warmup:
                        times.append(time.perf_counter() - t0)
            out.update(step_s=times, images_per_step=a.images, crops_per_step=4 * a.images, precision=enc.precision,
                       out_dtype=str(y.dtype), device_used=str(enc.device))
        else:
            ref = rs.import_reference("_1_embed_with_CLIP", seed=a.seed)
            tmp = tempfile.mkdtemp(prefix="b2c_refarm_")
            try:
                pool = os.path.join(tmp, "pool")
                os.makedirs(pool)
                for k in range(a.images):  # lossless files: the decoded pixels are the synthetic images
                    Image.fromarray(synthetic_image(1000 + k, 512, 512)).save(os.path.join(pool, f"{k:05d}.png"))
                times = []
                for s in range(a.warmup + a.steps):
                    root = os.path.join(tmp, f"step{s}")
                    os.makedirs(root)
                    for f in os.listdir(pool):  # fresh directory per step: no .pt files yet, nothing is skipped (_1:118-128)
                        os.link(os.path.join(pool, f), os.path.join(root, f))
                    ds = ref.Feature_Dataset(root, a.model, a.batch, num_workers=a.workers)  # model build is not timed
                    t0 = time.perf_counter()
                    ds.process()
                    dt = time.perf_counter() - t0
                    if s >= a.warmup:
                        times.append(dt)
                    shutil.rmtree(root)
                out.update(step_s=times, images_per_step=a.images, batch_size=a.batch, num_workers=a.workers,
                           device_used=str(ds.device))
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
