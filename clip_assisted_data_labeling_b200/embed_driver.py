"""Drop-in mirror of the reference's embedding driver (_1_embed_with_CLIP.py) on the B200 path.

``Feature_Dataset(root_dir, model_name, batch_size, model_path=None, force_reencode=False,
shuffle_filenames=True, num_workers=0, crop_names=[4 names])`` + ``.process()`` + ``__len__`` keep the
reference's names, positional order and defaults (_1_embed_with_CLIP.py:34-96).  What changes underneath:

  * DataLoader workers only decode (RawImageDataset); the 4 crops, the PIL-exact resize, normalisation and the
    ViT run fused on the GPU (CLIP_Encoder.encode_images_u8);
  * the per-image ``<img>.pt`` dict is written in the layout every consumer reads:
    ``{model_name: {img_stat_*: f32 0-d, crop_name: f32[1,E] CPU tensor}}`` (the 22 statistics come from the device,
    imgstats.image_stats), merged into an existing file unless force_reencode —
    for all B images of a batch and all 4 crops (the reference's collate transposition bug, SURVEY.md App. B1,
    is not reproduced);
  * resume is per image: an image whose ``.pt`` already holds ``model_name`` is skipped (_1:118-128 does this
    per batch);
  * one process per GPU: with RANK/WORLD_SIZE set (torchrun) the sorted path list is sharded contiguously,
    no collective is needed.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import random

import torch
from torch.utils.data import DataLoader

from .embedder import CLIP_Encoder, RawImageDataset, collate_raw, to_device_images
from .vit_arch import CROP_NAMES

IMG_EXTENSIONS = (".png", ".jpg", ".jpeg", ".JPEG", ".JPG", ".PNG")  # _1_embed_with_CLIP.py:47


def find_images(root_dir):
    out = []
    for root, _dirs, files in os.walk(root_dir):
        for name in files:
            if name.endswith(IMG_EXTENSIONS):
                out.append(os.path.join(root, name))
    return out


def shard_for_rank(paths, rank: int, world_size: int):
    """Contiguous slice [r*N/P, (r+1)*N/P) of the list for rank r."""
    n = len(paths)
    return paths[rank * n // world_size:(rank + 1) * n // world_size]


def build_feature_dict(features_per_image: torch.Tensor, crop_names, kept=None) -> dict:
    """features_per_image: f32 [4,E] on CPU in CROP_NAMES order -> {crop_name: f32[1,E]} (_1:146-161)."""
    d = {}
    for name in crop_names:
        i = CROP_NAMES.index(name)
        if kept is not None and not kept[i]:
            continue  # crop dropped as empty by extract_crops (utils/embedder.py:243-247)
        d[name] = features_per_image[i].unsqueeze(0).float().clone()
    return d


def feature_dict_from_flat(flat: torch.Tensor, n_stats: int, stat_names, crop_names, kept=None) -> dict:
    """flat: f32 [n_stats + 4*E] on CPU (the image's statistics, then its four crops in CROP_NAMES order) ->
    {img_stat_*: 0-d f32, crop_name: f32[1,E]} as VIEWS of that one storage: torch.save then writes a single storage
    entry instead of 26 (faster to write, and to read back in _2/_4/_5), the values and shapes are the reference's."""
    E = (flat.numel() - n_stats) // 4
    d = {n: flat[i] for i, n in enumerate(stat_names[:n_stats])}
    for name in crop_names:
        i = CROP_NAMES.index(name)
        if kept is not None and not kept[i]:
            continue  # crop dropped as empty by extract_crops (utils/embedder.py:243-247)
        d[name] = flat[n_stats + i * E:n_stats + (i + 1) * E].view(1, E)
    return d


def _write_feature_batch(model_name: str, force_reencode: bool, n_stats: int, stat_names, crop_names, rows, paths, kepts):
    """Writer-process task: rows f32 [b, n_stats + 4E] (numpy) -> b ``.pt`` files."""
    import numpy as np
    rows = torch.from_numpy(np.ascontiguousarray(rows))
    for r, path, kept in zip(rows, paths, kepts):
        save_feature_file(path, model_name, feature_dict_from_flat(r.clone(), n_stats, stat_names, crop_names, kept), force_reencode)
    return len(paths)


def save_feature_file(path: str, model_name: str, feature_dict: dict, force_reencode: bool) -> None:
    """Merge-and-save one ``<img>.pt`` (_1_embed_with_CLIP.py:138-170)."""
    final = {}
    if os.path.exists(path) and not force_reencode:
        try:
            final = torch.load(path, map_location="cpu")
        except Exception as e:  # noqa: BLE001
            print(f"Warning: Failed to load existing {path} for update: {e}")
    final[model_name] = feature_dict
    try:
        torch.save(final, path)
    except Exception as e:  # noqa: BLE001
        print(f"Error saving features to {path}: {e}")


def already_encoded(path: str, model_name: str) -> bool:
    if not os.path.exists(path):
        return False
    try:
        return model_name in torch.load(path, map_location="cpu").keys()
    except Exception as e:  # noqa: BLE001
        print(f"Warning: Could not load existing feature file {path}: {e}")
        return False


class Feature_Dataset:
    def __init__(self, root_dir, model_name, batch_size, model_path=None, force_reencode=False, shuffle_filenames=True,
                 num_workers=0, crop_names=("centre_crop", "square_padded_crop", "subcrop1", "subcrop2"),
                 rank=None, world_size=None, state_dict=None, encoder=None, writer_threads=4, packed_dir=None,
                 write_pt=True, img_stats=None, device_jpeg=None, writer_procs=None, allow_random_init=False):
        self.device = getattr(encoder, "device", "cuda") if encoder is not None else "cuda"
        self.root_dir = root_dir
        self.model_name = model_name
        self.force_reencode = force_reencode
        self.img_extensions = IMG_EXTENSIONS
        self.batch_size = batch_size
        self.crop_names = list(crop_names)
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size

        print("Searching images..")
        self.img_filepaths = find_images(root_dir)
        if shuffle_filenames and self.world_size == 1:
            random.shuffle(self.img_filepaths)
        else:  # ranks must agree on the order
            self.img_filepaths.sort()
        print(f"---> Found {len(self.img_filepaths)} images in {root_dir}")
        if self.world_size > 1:
            self.img_filepaths = shard_for_rank(self.img_filepaths, self.rank, self.world_size)
            print(f"---> rank {self.rank}/{self.world_size} owns {len(self.img_filepaths)} of them")

        if model_name.startswith("PE-"):
            raise ValueError("PE (perception_models) encoders are outside the B200 hot path; use an 'Arch/Dataset' CLIP name")
        elif "/" in model_name:
            self.encoder = encoder or CLIP_Encoder(model_name, model_path, device=self.device, state_dict=state_dict,
                                                   allow_random_init=allow_random_init)
        else:
            raise ValueError(f"Unknown model format: {model_name}. Expected 'PE-...' or 'Arch/Dataset'.")

        # baseline JPEGs: workers only Huffman-decode, the device finishes the decode (K14); everything else stays on Pillow
        self.device_jpeg = str(self.device).startswith("cuda") if device_jpeg is None else bool(device_jpeg)
        self.img_dataset = RawImageDataset(self.img_filepaths, device_jpeg=self.device_jpeg)
        kw = dict(batch_size=batch_size, shuffle=False, num_workers=num_workers, collate_fn=collate_raw)
        if num_workers > 0:
            kw["prefetch_factor"] = 2
        self._dl_kw = kw
        self.dataloader = DataLoader(self.img_dataset, **kw)
        self._writer_threads = writer_threads
        # .pt writing is pickling, i.e. GIL-bound (~1 ms per file): large jobs hand whole batches to writer PROCESSES
        # (spawned, never touch CUDA); small jobs keep the thread pool and skip the interpreter start-up cost
        if writer_procs is None:
            writer_procs = min(8, max(1, (os.cpu_count() or 2) // 2)) if len(self.img_filepaths) >= 2000 else 0
        self._writer_procs = writer_procs
        self.failed = []
        # with packed_dir set every rank appends its [B,4,E] blocks to one flat shard
        # (store.PackedWriter); write_pt=False skips the per-image pickles entirely (store.export_pt writes them later)
        self.packed_dir = packed_dir
        self.write_pt = write_pt or packed_dir is None
        # the 22 img_stat_* scalars the reference stores ahead of the crops (_1:149-152): computed on the device from the
        # same uint8 image (imgstats.image_stats); on by default wherever the device path runs
        self.img_stats = str(self.device).startswith("cuda") if img_stats is None else bool(img_stats)

    def __len__(self):
        return len(self.img_filepaths)

    def _pack_existing(self, packed, img_paths, stat_names):
        """Resume with a packed shard: rows of already-embedded images are read back from their ``.pt`` files (thread pool)
        and appended to the shard.  Returns the set of images whose file could NOT supply the row."""
        import numpy as np
        from .store import _load_one
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(32, 4 * (os.cpu_count() or 4))) as ex:
            loaded = list(ex.map(lambda p: _load_one(p, self.model_name, CROP_NAMES), img_paths))
        ok_paths, feats, kept, stats, bad = [], [], [], [], set()
        for p, r in zip(img_paths, loaded):
            if r is None or r[1].shape[1] != packed.embed or any(n not in r[3] for n in stat_names):
                bad.add(p)
                continue
            ok_paths.append(p)
            feats.append(r[1])
            kept.append([bool((r[2] >> ci) & 1) for ci in range(len(CROP_NAMES))])
            stats.append([r[3][n] for n in stat_names])
        if ok_paths:
            packed.append(np.stack(feats), ok_paths, kept, stats=np.asarray(stats, np.float32) if stat_names else None)
        return bad

    @torch.no_grad()
    def process(self):
        """One pass over the images.  The loop is software-pipelined: batch i's device work (decode finish, crops, tower,
        statistics, D2H into pinned memory) is enqueued before batch i-1's results are turned into files on the host, so
        the GPU computes while the host pickles and the DataLoader refills."""
        from . import _lib
        import ctypes as C
        import multiprocessing as mp
        lib = _lib.load()
        n_embedded, n_skipped = 0, 0
        R = self.encoder.img_resolution
        on_cuda = str(self.device).startswith("cuda")
        print(f"Embedding dataset of {len(self.img_filepaths)} images using {self.model_name}...")
        if self._writer_procs > 0 and self.write_pt:
            pool = concurrent.futures.ProcessPoolExecutor(max_workers=self._writer_procs, mp_context=mp.get_context("spawn"))
        else:
            pool = concurrent.futures.ThreadPoolExecutor(max_workers=self._writer_threads)
        pending = []
        stat_names = []
        if self.img_stats:
            from .imgstats import STAT_NAMES, image_stats
            stat_names = list(STAT_NAMES)
        n_stats = len(stat_names)
        packed = None
        if self.packed_dir is not None:
            from .store import PackedWriter
            packed = PackedWriter(self.packed_dir, self.model_name, self.encoder.embed_dim, CROP_NAMES, shard=self.rank,
                                  weights_source=getattr(self.encoder, "weights_source", None), stat_names=stat_names)
        slots = [None, None]  # pinned result buffers, alternating between consecutive batches
        import time as _time
        timing = os.environ.get("B2C_DRIVER_TIMING") is not None  # wall-clock seconds of the main thread per phase
        phase = {}

        def tick(name, t0):
            if timing:
                phase[name] = phase.get(name, 0.0) + _time.perf_counter() - t0
            return _time.perf_counter()

        def finish(job):
            """Host half of a batch: wait for its D2H copy, write files / append to the packed shard."""
            nonlocal pending
            ev, rows, save_paths, img_paths, shapes = job
            t0 = _time.perf_counter()
            if ev is not None:
                ev.synchronize()
            t0 = tick("wait for the device (previous batch)", t0)
            b = len(save_paths)
            rows = rows[:b]
            E = (rows.shape[1] - n_stats) // 4
            kept_all = []
            for (h, w) in shapes:
                g = (_lib.Crop * 4)()
                _lib.check(lib.b2c_crop_geometry(int(w), int(h), R, g), "b2c_crop_geometry")
                kept_all.append([g[i].cw > 0 for i in range(4)])
            if self.write_pt:
                if self._writer_procs > 0:
                    step = max(8, (b + self._writer_procs - 1) // self._writer_procs)
                    for lo in range(0, b, step):
                        pending.append(pool.submit(_write_feature_batch, self.model_name, self.force_reencode, n_stats, stat_names,
                                                   list(self.crop_names), rows[lo:lo + step].numpy().copy(), save_paths[lo:lo + step],
                                                   kept_all[lo:lo + step]))
                else:
                    for r, sp, kept in zip(rows, save_paths, kept_all):
                        fd = feature_dict_from_flat(r.clone(), n_stats, stat_names, self.crop_names, kept)
                        pending.append(pool.submit(save_feature_file, sp, self.model_name, fd, self.force_reencode))
            if packed is not None:
                # (append writes the block to the shard before it returns: views of the pinned slot are enough)
                packed.append(rows[:, n_stats:].reshape(b, 4, E), img_paths, kept_all,
                              stats=rows[:, :n_stats] if n_stats else None)
            if len(pending) > 4096:
                for fu in pending:
                    fu.result()
                pending = []
            tick("files / packed shard (host)", t0)

        # Resume (_1_embed_with_CLIP.py:118-128: skip an image whose .pt already holds this model's key): decided BEFORE
        # the images are read — the existing files are probed on a thread pool and only the rest goes to the DataLoader, so
        # a resumed run neither decodes what it will skip nor loads a .pt per image on the main thread inside the loop.
        dataloader = self.dataloader
        if self.write_pt and not self.force_reencode:
            t0 = _time.perf_counter()
            cand = [p for p in self.img_filepaths if os.path.exists(os.path.splitext(p)[0] + ".pt")]
            if cand:
                with concurrent.futures.ThreadPoolExecutor(max_workers=min(32, 4 * (os.cpu_count() or 4))) as ex:
                    flags = list(ex.map(lambda p: already_encoded(os.path.splitext(p)[0] + ".pt", self.model_name), cand))
                done = {p for p, f in zip(cand, flags) if f}
                if done and packed is not None:
                    # the shard is rewritten by every run, so what the resume skips has to come from the .pt files; an
                    # image whose file cannot supply a complete row (other embedding width, statistics missing while this
                    # run stores them) is embedded again instead
                    done -= self._pack_existing(packed, sorted(done), stat_names)
                if done:
                    n_skipped = len(done)
                    rest = [p for p in self.img_filepaths if p not in done]
                    dataloader = DataLoader(RawImageDataset(rest, device_jpeg=self.device_jpeg), **self._dl_kw) if rest else []
            tick("resume checks", t0)

        prev, k = None, 0
        t_it = _time.perf_counter()
        for images, img_paths in dataloader:
            t_it = tick("DataLoader (wait + unpickle)", t_it)
            todo_imgs, todo_paths, todo_img_paths = [], [], []
            for im, p in zip(images, img_paths):
                save_path = os.path.splitext(p)[0] + ".pt"
                if im is None:
                    self.failed.append(p)
                else:
                    todo_imgs.append(im)
                    todo_paths.append(save_path)
                    todo_img_paths.append(p)
            cur = None
            if todo_imgs:
                b = len(todo_imgs)
                dev_imgs = to_device_images(todo_imgs, self.device) if on_cuda else todo_imgs  # cpu: injected encoder (host-logic tests)
                if any(im is None for im in dev_imgs):  # a JPEG the device reported and neither the host stage nor Pillow decodes
                    keep = [j for j, im in enumerate(dev_imgs) if im is not None]
                    self.failed.extend(todo_img_paths[j] for j in range(b) if dev_imgs[j] is None)
                    dev_imgs = [dev_imgs[j] for j in keep]
                    todo_paths = [todo_paths[j] for j in keep]
                    todo_img_paths = [todo_img_paths[j] for j in keep]
                    b = len(keep)
                    if b == 0:
                        continue
                t_it = tick("gather + H2D + JPEG reconstruct launches", t_it)
                feats = self.encoder.encode_images_u8(dev_imgs)  # [B,4,E]
                t_it = tick("preprocess + tower launches", t_it)
                stats = image_stats(dev_imgs) if self.img_stats else None
                t_it = tick("image statistics launches", t_it)
                shapes = [(int(im.shape[0]), int(im.shape[1])) for im in dev_imgs]
                E = int(feats.shape[-1])
                if on_cuda:
                    slot = slots[k & 1]
                    if slot is None or slot.shape[0] < b:
                        slot = slots[k & 1] = torch.empty(max(b, self.batch_size), n_stats + 4 * E, dtype=torch.float32, pin_memory=True)
                    # rows are assembled on the device: a D2H copy into a strided slice of the pinned slot would go through
                    # a synchronous staging copy and stall the host until the whole batch has been computed
                    rows_dev = feats.reshape(b, 4 * E)
                    if stats is not None:
                        rows_dev = torch.cat([stats.to(torch.float32), rows_dev], dim=1)
                    slot[:b].copy_(rows_dev, non_blocking=True)
                    # blocking: the main thread sleeps while it waits for the batch instead of spinning on a core the
                    # DataLoader workers (and, on a multi-GPU box, the other ranks) need
                    ev = torch.cuda.Event(blocking=True)
                    ev.record()
                    cur = (ev, slot, todo_paths, todo_img_paths, shapes)
                    k += 1
                else:
                    rows = torch.empty(b, n_stats + 4 * E, dtype=torch.float32)
                    if stats is not None:
                        rows[:, :n_stats] = stats.to(torch.float32)
                    rows[:, n_stats:] = feats.reshape(b, 4 * E).float()
                    cur = (None, rows, todo_paths, todo_img_paths, shapes)
                n_embedded += b
                t_it = tick("D2H enqueue", t_it)
            if prev is not None:
                finish(prev)
            prev = cur
            t_it = _time.perf_counter()
        if prev is not None:
            finish(prev)
        for fu in pending:
            fu.result()
        pool.shutdown()
        if packed is not None:
            packed.close()
        if timing:
            from . import embedder as _emb, jpeg as _jpeg
            for st in (_jpeg._coef_staging, _jpeg._packed_staging, _jpeg._file_staging, _emb._pixel_staging):
                if st is not None:
                    for k2, v2 in st.timing.items():
                        phase[k2] = phase.get(k2, 0.0) + v2
                    st.timing = {}
            print("main-thread seconds per phase: " + ", ".join(f"{k}: {v:.2f}" for k, v in phase.items()), file=__import__("sys").stderr)
        print("\n--- Feature encoding done! ---\n")
        print(f"Embedded {n_embedded} images ({n_skipped} images were already embedded). "
              f"Features saved with model key '{self.model_name}'.")
        if self.failed:
            print(f"{len(self.failed)} images could not be decoded and were skipped: {self.failed[:5]}...")
        print(f"Feature vector dicts were saved alongside original images in {self.root_dir}")
        print(f"Crop names that were processed: {self.crop_names}")
        return n_embedded, n_skipped


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--root_dir", type=str, required=True, help="Root directory of the dataset (can contain subdirectories)")
    parser.add_argument("--models_to_use", type=str, nargs="+", default=["ViT-L-14-336/openai"],
                        help="Which CLIP models to use (Arch/pretrained)")
    parser.add_argument("--batch_size", type=int, default=8, help="Number of images to encode at once")
    parser.add_argument("--num_workers", type=int, default=4, help="Number of workers for the dataloader")
    parser.add_argument("--force_reencode", action="store_true", help="Force re-encoding of all images for the specified models")
    parser.add_argument("--model_path", type=str, default=None, help="Path to a local checkpoint file or directory (optional)")
    parser.add_argument("--packed_dir", type=str, default=None,
                        help="Also write one packed [N,4,E] shard per rank here (store.py; not a reference flag)")
    parser.add_argument("--no_pt", action="store_true", help="With --packed_dir: skip the per-image .pt files (export them later)")
    parser.add_argument("--allow_random_init", action="store_true",
                        help="Benchmarks only: seeded random weights when no checkpoint is found (never for real data)")
    args = parser.parse_args(argv)
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    print(f"Embedding all imgs with {len(args.models_to_use)} models: \n--> {args.models_to_use}")
    for model_name in args.models_to_use:
        print(f"\n--- Processing model: {model_name} ---")
        Feature_Dataset(args.root_dir, model_name, args.batch_size, model_path=args.model_path,
                        force_reencode=args.force_reencode, num_workers=args.num_workers, crop_names=CROP_NAMES,
                        packed_dir=args.packed_dir, write_pt=not args.no_pt, allow_random_init=args.allow_random_init).process()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    mp.set_start_method("spawn")
    main()
