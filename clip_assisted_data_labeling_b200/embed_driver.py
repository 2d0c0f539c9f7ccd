"""Drop-in mirror of the reference's embedding driver (_1_embed_with_CLIP.py) on the B200 path.

``Feature_Dataset(root_dir, model_name, batch_size, model_path=None, force_reencode=False,
shuffle_filenames=True, num_workers=0, crop_names=[4 names])`` + ``.process()`` + ``__len__`` keep the
reference's names, positional order and defaults (_1_embed_with_CLIP.py:34-96).  What changes underneath:

  * DataLoader workers only decode (RawImageDataset); the 4 crops, the PIL-exact resize, normalisation and the
    ViT run fused on the GPU (CLIP_Encoder.encode_images_u8);
  * the per-image ``<img>.pt`` dict is written in the layout every consumer reads:
    ``{model_name: {crop_name: f32[1,E] CPU tensor}}``, merged into an existing file unless force_reencode —
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
                 write_pt=True, img_stats=None, device_jpeg=None):
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
            self.encoder = encoder or CLIP_Encoder(model_name, model_path, device=self.device, state_dict=state_dict)
        else:
            raise ValueError(f"Unknown model format: {model_name}. Expected 'PE-...' or 'Arch/Dataset'.")

        # baseline JPEGs: workers only Huffman-decode, the device finishes the decode (K14); everything else stays on Pillow
        self.device_jpeg = str(self.device).startswith("cuda") if device_jpeg is None else bool(device_jpeg)
        self.img_dataset = RawImageDataset(self.img_filepaths, device_jpeg=self.device_jpeg)
        kw = dict(batch_size=batch_size, shuffle=False, num_workers=num_workers, collate_fn=collate_raw)
        if num_workers > 0:
            kw["prefetch_factor"] = 2
        self.dataloader = DataLoader(self.img_dataset, **kw)
        self._writer_threads = writer_threads
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

    @torch.no_grad()
    def process(self):
        from . import _lib
        import ctypes as C
        lib = _lib.load()
        n_embedded, n_skipped = 0, 0
        R = self.encoder.img_resolution
        print(f"Embedding dataset of {len(self.img_filepaths)} images using {self.model_name}...")
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=self._writer_threads)
        pending = []
        packed = None
        if self.packed_dir is not None:
            from .store import PackedWriter
            packed = PackedWriter(self.packed_dir, self.model_name, self.encoder.embed_dim, CROP_NAMES, shard=self.rank)
        for images, img_paths in self.dataloader:
            todo_imgs, todo_paths, todo_img_paths = [], [], []
            for im, p in zip(images, img_paths):
                save_path = os.path.splitext(p)[0] + ".pt"
                if im is None:
                    self.failed.append(p)
                elif self.write_pt and not self.force_reencode and already_encoded(save_path, self.model_name):
                    n_skipped += 1
                else:
                    todo_imgs.append(im)
                    todo_paths.append(save_path)
                    todo_img_paths.append(p)
            if todo_imgs:
                if str(self.device).startswith("cuda"):
                    dev_imgs = to_device_images(todo_imgs, self.device)
                else:  # only reachable with an injected encoder (host-logic tests)
                    dev_imgs = todo_imgs
                feats = self.encoder.encode_images_u8(dev_imgs).cpu()  # [B,4,E], one D2H per batch
                stats = None
                if self.img_stats:
                    from .imgstats import image_stats, stats_dict
                    stats = image_stats(dev_imgs).cpu()
                kept_all = []
                for bi, (im, f, sp) in enumerate(zip(dev_imgs, feats, todo_paths)):
                    g = (_lib.Crop * 4)()
                    _lib.check(lib.b2c_crop_geometry(int(im.shape[1]), int(im.shape[0]), R, g), "b2c_crop_geometry")
                    kept = [g[i].cw > 0 for i in range(4)]
                    kept_all.append(kept)
                    if self.write_pt:
                        fd = build_feature_dict(f, self.crop_names, kept)
                        if stats is not None:
                            fd = {**stats_dict(stats[bi]), **fd}
                        pending.append(pool.submit(save_feature_file, sp, self.model_name, fd, self.force_reencode))
                if packed is not None:
                    packed.append(feats, todo_img_paths, kept_all)
                n_embedded += len(todo_imgs)
            if len(pending) > 4096:
                for fu in pending:
                    fu.result()
                pending = []
        for fu in pending:
            fu.result()
        pool.shutdown()
        if packed is not None:
            packed.close()
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
    args = parser.parse_args(argv)
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    print(f"Embedding all imgs with {len(args.models_to_use)} models: \n--> {args.models_to_use}")
    for model_name in args.models_to_use:
        print(f"\n--- Processing model: {model_name} ---")
        Feature_Dataset(args.root_dir, model_name, args.batch_size, model_path=args.model_path,
                        force_reencode=args.force_reencode, num_workers=args.num_workers, crop_names=CROP_NAMES,
                        packed_dir=args.packed_dir, write_pt=not args.no_pt).process()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    mp.set_start_method("spawn")
    main()
