"""Drop-in mirror of the reference's embedder call surface (utils/embedder.py) on the B200 path.

Kept from the reference (names, positional order, defaults, attributes):
  * ``CLIP_Encoder(model_name, model_path=None, device=None)`` with ``device, precision, model_name,
    model_architecture, pretrained_dataset, model, preprocess, img_resolution`` (utils/embedder.py:59-86),
    ``get_preprocess_transform()`` (90-92), ``encode_image(Tensor[n,3,R,R]) -> Tensor[n,E]`` unit-norm (94-100);
  * ``CustomImageDataset(image_paths, crop_names, preprocess_transform)`` with ``__len__``,
    ``__getitem__ -> (Tensor[k,3,R,R], list[str], str, dict)`` and ``extract_crops(pil)`` (153-251).
Added (SURVEY.md §8b "additive fast entry"):
  * ``CLIP_Encoder.encode_images_u8(uint8 [B,H,W,3] | list of ragged HWC) -> f32 [B,4,E]`` — crops,
    PIL-exact resize, normalise and the tower fused on the device;
  * ``RawImageDataset`` — decodes to uint8 HWC only, leaving crops/resize to the GPU.
Deliberate deviations, both documented in DESIGN.md: outputs are float32 (the reference returns fp16 on
CUDA and every consumer casts with ``.float()``); a failed image is skipped and reported instead of being
silently replaced by a random other image (utils/embedder.py:176-181).
"""
from __future__ import annotations

import os

import torch
from torch.utils.data import Dataset

from .vit_arch import ARCHS, CROP_NAMES, OPENAI_MEAN, OPENAI_STD, activation_for, random_state_dict, split_model_name


def _open_clip_val_transform(image_size: int):
    """The open_clip validation transform (third-party): Resize(R, BICUBIC) ->
    CenterCrop(R) -> RGB -> ToTensor -> Normalize(OPENAI mean/std)."""
    from torchvision import transforms

    return transforms.Compose([
        transforms.Resize(image_size, interpolation=transforms.InterpolationMode.BICUBIC),
        transforms.CenterCrop(image_size),
        _ConvertRGB(),
        transforms.ToTensor(),
        transforms.Normalize(mean=OPENAI_MEAN, std=OPENAI_STD),
    ])


class _ConvertRGB:  # picklable (DataLoader workers use spawn, _1_embed_with_CLIP.py:202)
    def __call__(self, im):
        return im.convert("RGB")


# open_clip's own download locations (third-party; the reference passes cache_dir=model_path, utils/embedder.py:66-73):
# OpenAI weights are TorchScript archives under ~/.cache/clip, LAION weights live in the Hugging Face hub cache.
_OPENAI_FILES = {"ViT-B-32": "ViT-B-32.pt", "ViT-L-14": "ViT-L-14.pt", "ViT-L-14-336": "ViT-L-14-336px.pt"}
_HF_OPENAI_REPOS = {"ViT-B-32": "clip-vit-base-patch32", "ViT-L-14": "clip-vit-large-patch14",
                    "ViT-L-14-336": "clip-vit-large-patch14-336"}
_CKPT_SUFFIXES = (".pt", ".pth", ".bin", ".safetensors")
_HUB_FILES = ("open_clip_model.safetensors", "open_clip_pytorch_model.bin", "model.safetensors", "pytorch_model.bin")


def _hub_snapshot_files(hub_dir: str, arch: str, pretrained: str):
    """Checkpoint files of hub repositories whose name matches the architecture (and, for LAION tags, the tag)."""
    out = []
    if not os.path.isdir(hub_dir):
        return out
    want = [arch.lower()] if pretrained != "openai" else [_HF_OPENAI_REPOS.get(arch, "\0").lower()]
    tag = pretrained.lower().split("_")[0]  # 'laion2b_s32b_b79k' -> 'laion2b'
    for repo in sorted(os.listdir(hub_dir)):
        low = repo.lower()
        if not low.startswith("models--") or not any(w in low for w in want):
            continue
        if pretrained == "openai":
            if not low.endswith(want[0]):  # 'clip-vit-large-patch14' must not match '...patch14-336'
                continue
        elif tag not in low:
            continue
        snaps = os.path.join(hub_dir, repo, "snapshots")
        for snap in sorted(os.listdir(snaps)) if os.path.isdir(snaps) else []:
            for f in _HUB_FILES:
                if os.path.isfile(os.path.join(snaps, snap, f)):
                    out.append(os.path.join(snaps, snap, f))
    return out


def _find_checkpoint(model_path, arch: str, pretrained: str, search_caches: bool = True):
    """Where the weights of ``arch/pretrained`` are on this machine, or None.  ``model_path`` may be the file itself or a
    directory (the reference hands it to open_clip as cache_dir); then open_clip's default download locations."""
    if model_path is not None and os.path.isfile(model_path):
        return model_path
    dirs = [model_path] if model_path is not None and os.path.isdir(model_path) else []
    if search_caches:
        if os.environ.get("B2C_CLIP_CACHE"):
            dirs.append(os.environ["B2C_CLIP_CACHE"])
        dirs.append(os.path.expanduser("~/.cache/clip"))
    stems = [f"{arch}_{pretrained}", f"{arch}-{pretrained}", arch]
    for d in dirs:
        if not os.path.isdir(d):
            continue
        files = sorted(os.listdir(d))
        if pretrained == "openai" and _OPENAI_FILES.get(arch) in files:
            return os.path.join(d, _OPENAI_FILES[arch])
        for stem in stems:
            for f in files:
                base = os.path.splitext(f)[0]
                # 'ViT-L-14' must not pick up 'ViT-L-14-336...': the stem has to end the name or be followed by a separator
                if f.endswith(_CKPT_SUFFIXES) and (base == stem or base.startswith(stem + "_") or base.startswith(stem + ".")):
                    return os.path.join(d, f)
        hub = _hub_snapshot_files(d, arch, pretrained) + _hub_snapshot_files(os.path.join(d, "hub"), arch, pretrained)
        if hub:
            return hub[0]
    if search_caches:
        hub_dir = os.path.join(os.environ.get("HF_HOME", os.path.expanduser("~/.cache/huggingface")), "hub")
        hub = _hub_snapshot_files(hub_dir, arch, pretrained)
        if hub:
            return hub[0]
    return None


def _hf_clip_to_open_clip(sd: dict) -> dict:
    """transformers' CLIP(Vision)Model parameter names -> open_clip ``visual.*`` names."""
    pre = "vision_model."
    out = {}
    direct = {"embeddings.class_embedding": "class_embedding", "embeddings.patch_embedding.weight": "conv1.weight",
              "embeddings.position_embedding.weight": "positional_embedding", "pre_layrnorm.weight": "ln_pre.weight",
              "pre_layrnorm.bias": "ln_pre.bias", "post_layernorm.weight": "ln_post.weight", "post_layernorm.bias": "ln_post.bias"}
    per_layer = {"layer_norm1": "ln_1", "layer_norm2": "ln_2", "self_attn.out_proj": "attn.out_proj", "mlp.fc1": "mlp.c_fc",
                 "mlp.fc2": "mlp.c_proj"}
    qkv = {}
    for k, v in sd.items():
        if k == "visual_projection.weight":
            out["proj"] = v.t().contiguous()
            continue
        if not k.startswith(pre):
            continue
        k = k[len(pre):]
        if k in direct:
            out[direct[k]] = v
        elif k.startswith("encoder.layers."):
            i, rest = k[len("encoder.layers."):].split(".", 1)
            mod, leaf = rest.rsplit(".", 1)
            if mod in per_layer:
                out[f"transformer.resblocks.{i}.{per_layer[mod]}.{leaf}"] = v
            elif mod in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"):
                qkv.setdefault((i, leaf), {})[mod[-6]] = v
    for (i, leaf), parts in qkv.items():
        if set(parts) == {"q", "k", "v"}:
            out[f"transformer.resblocks.{i}.attn.in_proj_{leaf}"] = torch.cat([parts["q"], parts["k"], parts["v"]], dim=0)
    return out


def _load_checkpoint(path: str) -> dict:
    """State dict of a checkpoint file: safetensors, a pickled state dict (optionally wrapped in {'state_dict': ...}), or
    one of OpenAI's TorchScript archives; transformers-format CLIP checkpoints are renamed to open_clip's keys."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    else:
        try:
            sd = torch.load(path, map_location="cpu", weights_only=True)
        except Exception:  # noqa: BLE001  (OpenAI's .pt files are TorchScript archives)
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        if not isinstance(sd, dict) and hasattr(sd, "state_dict"):  # torch.load may hand back the scripted module itself
            sd = sd.state_dict()
        if isinstance(sd, dict) and isinstance(sd.get("state_dict"), dict):
            sd = sd["state_dict"]
    if not isinstance(sd, dict):
        raise ValueError(f"{path}: not a state dict")
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items() if isinstance(v, torch.Tensor)}
    if any(k.startswith("vision_model.") for k in sd):
        sd = _hf_clip_to_open_clip(sd)
    return sd


class CLIP_Encoder:
    """utils/embedder.py:58-100 on sm_100a kernels.

    Weights: ``state_dict`` (open_clip ``visual.*`` names) if given, else the checkpoint ``_find_checkpoint`` locates
    (``model_path`` file / directory, then open_clip's download caches).  The reference never returns a model without
    pretrained weights (open_clip downloads them, utils/embedder.py:66-73); there is no network here, so a missing
    checkpoint is an error — embeddings of random weights written under the real model key would be consumed silently
    by _2/_4/_5.  ``allow_random_init=True`` (benches, tests: BASELINE.json asks for random-init weights of the named
    architecture) opts into a seeded random initialisation; ``weights_source`` says which of the three happened."""

    def __init__(self, model_name, model_path=None, device=None, state_dict=None, seed=0, allow_random_init=False):
        from .vit import VisionTower, _require_cuda

        self.device = device if device else "cuda"
        if not str(self.device).startswith("cuda"):
            raise RuntimeError("CLIP_Encoder (B200 path) runs on CUDA only; there is no CPU fallback")
        _require_cuda(self.device)  # fail for the real reason (no sm_100a device) before looking for weights
        # the reference's values are 'fp16' (CUDA) | 'fp32' (CPU) (utils/embedder.py:61); this path runs bf16 tensor-core
        # GEMMs with an fp32 residual stream and returns fp32 — documented in INTEGRATION.md
        self.precision = "bf16"
        self.model_name = model_name
        self.model_architecture, self.pretrained_dataset = split_model_name(model_name)
        if self.model_architecture not in ARCHS:
            raise ValueError(f"unknown architecture {self.model_architecture}; known: {sorted(ARCHS)}")
        cfg = ARCHS[self.model_architecture]
        print(f"Loading CLIP model {self.model_name}...")
        self.weights_source = "state_dict"
        if state_dict is None:
            ckpt = _find_checkpoint(model_path, self.model_architecture, self.pretrained_dataset)
            if ckpt is not None:
                state_dict = _load_checkpoint(ckpt)
                self.weights_source = ckpt
            elif allow_random_init:
                state_dict = random_state_dict(cfg, seed=seed)
                self.weights_source = f"random-init(seed={seed})"
                print(f"NOTE: {model_name}: {self.weights_source} was requested (allow_random_init=True); these are not pretrained weights")
            else:
                raise FileNotFoundError(
                    f"no checkpoint for {model_name} (model_path={model_path!r}, $B2C_CLIP_CACHE, ~/.cache/clip, the Hugging Face hub "
                    "cache): open_clip would download it, this machine cannot.  Pass model_path=<file or directory> or "
                    "state_dict=...; random weights need the explicit allow_random_init=True")
        self.model = VisionTower(cfg, activation_for(self.pretrained_dataset), self.device)
        self.model.load_state_dict(state_dict)
        self.preprocess = _open_clip_val_transform(cfg["image"])
        self.img_resolution = cfg["image"]
        self.embed_dim = cfg["embed"]
        print(f"CLIP model {self.model_name} with img_resolution {self.img_resolution} loaded on {self.device} (weights: {self.weights_source})!")

    def get_preprocess_transform(self):
        return self.preprocess

    @torch.no_grad()
    def encode_image(self, preprocessed_images: torch.Tensor) -> torch.Tensor:
        """[n,3,R,R] normalised crops -> [n,E] unit-norm (utils/embedder.py:94-100)."""
        return self.model.forward_pixels(preprocessed_images)

    @torch.no_grad()
    def encode_images_u8(self, images) -> torch.Tensor:
        """uint8 [B,H,W,3] (or a list of ragged HWC uint8 tensors) -> f32 [B,4,E]: the 4 crops in the order
        [centre_crop, square_padded_crop, subcrop1, subcrop2] (_1_embed_with_CLIP.py:200)."""
        return self.model.encode_u8(images)


    @torch.no_grad()
    def encode_host_batches(self, batches):
        """Bulk form of ``encode_images_u8`` for data that lives on the host: ``batches`` is an iterable of uint8
        [B,H,W,3] host tensors (pinned for full speed); yields one f32 [B,4,E] pinned host tensor per batch, in order.
        Double-buffered: the H2D copy of batch i+1 runs on a copy stream under the compute of batch i, the D2H copy of
        the embeddings follows the compute on its stream, and batch i is handed out once its copy has landed.  A yielded
        tensor is a view of a pinned ring of four buffers: it stays valid until two more batches have been yielded (the
        D2H copy of batch i+1 is already in flight when batch i is handed out, so two slots would not do)."""
        dev = torch.device(self.device)
        with torch.cuda.device(dev):
            compute = torch.cuda.current_stream()
            copy = torch.cuda.Stream()
            bufs, freed = [None, None], [None, None]   # device input ring + "its consumer has run" events
            n_out = 4
            outs = [None] * n_out                       # pinned output ring (see the docstring for its depth)
            pending = None  # (slot, event, rows) of the batch whose embeddings are on their way to the host
            it = iter(batches)

            def stage(host, slot):
                """H2D of one batch into ring slot `slot` on the copy stream; waits only for that slot's last consumer."""
                if host.dim() != 4 or host.shape[-1] != 3 or host.dtype != torch.uint8:
                    raise ValueError(f"expected uint8 [B,H,W,3], got {host.dtype} {tuple(host.shape)}")
                if bufs[slot] is None or bufs[slot].shape[1:] != host.shape[1:] or bufs[slot].shape[0] < host.shape[0]:
                    if freed[slot] is not None:
                        freed[slot].synchronize()
                    bufs[slot] = torch.empty(tuple(host.shape), dtype=torch.uint8, device=dev)
                    copy.wait_stream(compute)  # the fresh block may be memory that work queued on `compute` still uses
                with torch.cuda.stream(copy):
                    if freed[slot] is not None:
                        copy.wait_event(freed[slot])
                    d = bufs[slot][:host.shape[0]]
                    d.copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                return d, ev

            k = 0
            nxt = next(it, None)
            staged = stage(nxt, 0) if nxt is not None else None
            while staged is not None:
                d, ev = staged
                slot = k & 1
                nxt = next(it, None)
                compute.wait_event(ev)
                feats = self.model.encode_u8(d)
                freed[slot] = torch.cuda.Event()
                freed[slot].record(compute)
                staged = stage(nxt, slot ^ 1) if nxt is not None else None  # copies under the kernels just enqueued
                b = feats.shape[0]
                oslot = k % n_out
                if outs[oslot] is None or outs[oslot].shape[0] < b:
                    outs[oslot] = torch.empty(b, 4, feats.shape[-1], dtype=torch.float32, pin_memory=True)
                outs[oslot][:b].copy_(feats, non_blocking=True)
                done = torch.cuda.Event()
                done.record(compute)
                if pending is not None:
                    ps, pe, pb = pending
                    pe.synchronize()
                    yield outs[ps][:pb]
                pending = (oslot, done, b)
                k += 1
            if pending is not None:
                ps, pe, pb = pending
                pe.synchronize()
                yield outs[ps][:pb]
            copy.synchronize()


class CustomImageDataset(Dataset):
    """utils/embedder.py:153-251 — host (PIL) implementation kept for the reference-compatible
    ``encode_image`` path and as the CPU side of parity tests.  The fourth element is an EMPTY dict where the reference
    returns the 22 ``img_stat_*`` scalars (:170,175): those are computed on the device (imgstats.image_stats, K13) by
    ``Feature_Dataset`` for the images already in HBM — DataLoader workers never touch CUDA.  A file
    that fails to decode raises here instead of being replaced by a random other image (:176-181)."""

    def __init__(self, image_paths, crop_names, preprocess_transform):
        self.image_paths = image_paths
        self.crop_names = crop_names
        self.preprocess_transform = preprocess_transform

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, idx):
        from PIL import Image
        img_path = self.image_paths[idx]
        pil_img = Image.open(img_path).convert("RGB")
        raw_crops, crop_names_list = self.extract_crops(pil_img)
        processed = torch.stack([self.preprocess_transform(c) for c in raw_crops])
        return processed, crop_names_list, img_path, {}

    def extract_crops(self, pil_img):
        """Same crop rectangles as utils/embedder.py:184-251 (centre / black-padded square / 2 sub-crops)."""
        from PIL import Image
        W, H = pil_img.width, pil_img.height
        crops, names = [], []
        if "centre_crop" in self.crop_names:
            s = min(W, H)
            top, left = int(round((H - s) / 2.0)), int(round((W - s) / 2.0))  # torchvision CenterCrop
            crops.append(pil_img.crop((left, top, left + s, top + s)))
            names.append("centre_crop")
        if "square_padded_crop" in self.crop_names:
            S = max(W, H)
            canvas = Image.new("RGB", (S, S), (0, 0, 0))
            canvas.paste(pil_img, ((S - W) // 2, (S - H) // 2))
            crops.append(canvas)
            names.append("square_padded_crop")
        if any("subcrop1" in n for n in self.crop_names) or any("subcrop2" in n for n in self.crop_names):
            sizes = [int((W * H * 0.15) ** 0.5), int((W * H * 0.1) ** 0.5)]
            if W >= H:
                centers = [(W // 4, H // 2), (W // 4 * 3, H // 2)]
            else:
                centers = [(W // 2, H // 4), (W // 2, H // 4 * 3)]
            for name, (cx, cy), sz in zip(["subcrop1", "subcrop2"], centers, sizes):
                if name not in self.crop_names:
                    continue
                left, top = max(0, cx - sz // 2), max(0, cy - sz // 2)
                right, bottom = min(W, left + sz), min(H, top + sz)
                if right - left > 0 and bottom - top > 0:
                    crops.append(pil_img.crop((left, top, right, bottom)))
                    names.append(name)
                else:
                    print(f"Warning: {name} resulted in zero size.")
        return crops, names


class RawImageDataset(Dataset):
    """Decode only: returns (item, path).  Crops, resize and normalisation run on the GPU (b2c_preprocess_4crop).
    item is
      * ``("jpegf", info_bytes, huff_bytes, uint8 file bytes)`` for a single-scan sequential JPEG when ``device_jpeg`` is
        on: the worker only parses the markers (jpeg.prepare_file) and the main process decodes the file ENTIRELY on the
        device — Huffman stage (b2c_jpeg_huff_decode) and reconstruction (jpeg.decode_device), bit-exact with Pillow,
        SURVEY.md §8f row 2; a stream the device reports (damaged, truncated) is retried on the host there;
      * ``("jpegp", info_bytes, uint8 packed coefficients)`` for the JPEGs whose Huffman stage stays on the host
        (progressive, multi-scan; or every JPEG when ``device_huffman`` is off): the worker does the serial part
        (jpeg.entropy_decode_packed) and the main process finishes the decode on the device (jpeg.reconstruct_packed)
        (``("jpeg", info_bytes, int16 dense coefficients)`` items are accepted too);
      * a uint8 HWC tensor decoded with Pillow (`Image.open(path).convert('RGB')`, utils/embedder.py:167) for every
        other format and for JPEG streams the device path does not cover;
      * ``None`` for a file that fails to decode — reported by the driver, never silently substituted."""

    def __init__(self, image_paths, device_jpeg: bool = False, device_huffman: bool = None):
        import os
        self.image_paths = image_paths
        self.device_jpeg = device_jpeg
        self.device_huffman = (os.environ.get("B2C_DEVICE_HUFFMAN", "1") != "0") if device_huffman is None else bool(device_huffman)

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, idx):
        import io
        import numpy as np
        from PIL import Image
        path = self.image_paths[idx]
        try:
            src = path
            if self.device_jpeg and path.lower().endswith((".jpg", ".jpeg")):
                from . import _lib, jpeg
                with open(path, "rb") as fh:
                    data = fh.read()
                try:
                    if self.device_huffman:
                        try:
                            info, huff, raw = jpeg.prepare_file(data)
                            return ("jpegf", bytes(info), bytes(huff), raw), path
                        except jpeg.UnsupportedJPEG:
                            pass  # progressive / multi-scan: the Huffman stage of these stays on the host
                    info, packed = jpeg.entropy_decode_packed(data)
                    return ("jpegp", bytes(info), packed), path
                except (jpeg.UnsupportedJPEG, _lib.B2CError):
                    # CMYK, arithmetic-coded, ... or not a JPEG stream at all (a .jpg that is really a PNG, a damaged
                    # file): Pillow gets the bytes already read and has the last word, exactly like the reference
                    src = io.BytesIO(data)
            with Image.open(src) as im:
                arr = np.array(im.convert("RGB"))  # (a writable copy: torch.from_numpy shares it)
            return torch.from_numpy(arr), path
        except Exception as e:  # noqa: BLE001
            print(f"Error loading image {path}: {e}")
            return None, path


_pixel_staging = None


def to_device_images(items, device):
    """RawImageDataset items (no ``None``) -> uint8 [H,W,3] device tensors, same order.  Pillow-decoded tensors go through
    one pinned gather + one H2D copy (the device tensors are views of that buffer); entropy-decoded JPEGs likewise and are
    then reconstructed on the device in one batched call; whole JPEG files are decoded on the device (an entry is None
    only for a file that neither the device, nor the host stage, nor Pillow can decode: the caller reports it)."""
    global _pixel_staging
    from . import jpeg
    out = [None] * len(items)
    jobs, where, pjobs, pjwhere, pix, pwhere, fjobs, fwhere = [], [], [], [], [], [], [], []
    for i, it in enumerate(items):
        if isinstance(it, tuple) and it[0] == "jpegf":
            fjobs.append((jpeg.JpegInfo.from_buffer_copy(it[1]), jpeg.JpegHuff.from_buffer_copy(it[2]), it[3]))
            fwhere.append(i)
        elif isinstance(it, tuple) and it[0] == "jpegp":
            pjobs.append((jpeg.JpegInfo.from_buffer_copy(it[1]), it[2]))
            pjwhere.append(i)
        elif isinstance(it, tuple) and it[0] == "jpeg":
            jobs.append((jpeg.JpegInfo.from_buffer_copy(it[1]), it[2]))
            where.append(i)
        else:
            pix.append(it.contiguous())
            pwhere.append(i)
    if pix:
        if _pixel_staging is None:
            _pixel_staging = jpeg.Staging(torch.uint8)
        with torch.cuda.device(torch.device(device)):
            dflat, offs = _pixel_staging.gather(pix, torch.device(device))
        for i, t, o in zip(pwhere, pix, offs[:-1]):
            out[i] = dflat[int(o):int(o) + t.numel()].view(t.shape)
    for i, t in zip(fwhere, jpeg.decode_device(fjobs, device)[0]):
        out[i] = t
    for i, t in zip(pjwhere, jpeg.reconstruct_packed(pjobs, device)):
        out[i] = t
    for i, t in zip(where, jpeg.reconstruct(jobs, device)):
        out[i] = t
    return out


def collate_raw(batch):
    """Keep ragged images as a list (default_collate would try to stack them).  Runs in the DataLoader worker: the
    coefficient tensors of the batch's JPEG items are concatenated into ONE tensor there, and each item keeps a view of it —
    the main process then receives one shared-memory segment per batch instead of one per image (unpickling 256 segments
    costs it ~55 ms per batch), and the views gather into pinned memory as before."""
    items = [b[0] for b in batch]
    for kind in ("jpegf", "jpegp", "jpeg"):
        idx = [i for i, it in enumerate(items) if isinstance(it, tuple) and it[0] == kind]
        if len(idx) > 1:
            parts = [items[i][-1] for i in idx]  # buffers are multiples of 16 bytes: views stay aligned
            total = sum(t.numel() for t in parts)
            if torch.utils.data.get_worker_info() is not None:
                # concatenate straight into shared memory (what default_collate does for stacked batches): handing the
                # batch to the main process then moves a handle, not another copy of every file
                storage = parts[0]._typed_storage()._new_shared(total, device=parts[0].device)
                pack = torch.cat(parts, out=parts[0].new(storage).resize_(total))
            else:
                pack = torch.cat(parts)
            off = 0
            for i in idx:
                n = items[i][-1].numel()
                items[i] = items[i][:-1] + (pack[off:off + n],)
                off += n
    return items, [b[1] for b in batch]
