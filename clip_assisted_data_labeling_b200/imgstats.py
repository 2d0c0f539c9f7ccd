"""The 22 ``img_stat_*`` scalars of the reference's ``ImageFeaturizer`` (utils/image_features.py:52-94) on the device
(SURVEY.md §8f row 2): ``image_stats(images)`` runs ``b2c_image_stats`` over the same device-resident uint8 images the
4-crop preprocess reads and returns float64 ``[B, 22]`` in the reference's dict order; ``stats_dict`` turns one row into
the ``{name: f32 0-d tensor}`` entries the reference stores in the ``.pt`` file ahead of the crop embeddings
(_1_embed_with_CLIP.py:149-161)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

STAT_NAMES = ["img_stat_width", "img_stat_height", "img_stat_aspect_ratio", "img_stat_mean_color", "img_stat_std_color",
              "img_stat_mean_red", "img_stat_mean_green", "img_stat_mean_blue", "img_stat_std_red", "img_stat_std_green",
              "img_stat_std_blue", "img_stat_mean_gray", "img_stat_std_gray", "img_stat_mean_hue", "img_stat_mean_sat",
              "img_stat_mean_val", "img_stat_std_hue", "img_stat_std_sat", "img_stat_std_val", "img_stat_colorfulness",
              "img_stat_image_entropy", "img_stat_laplacian_variance"]  # utils/image_features.py:63-86


def target_size(W: int, H: int):
    """(new_w, new_h) the reference resizes a W x H image to (utils/image_features.py:57-58)."""
    nw, nh = C.c_int(), C.c_int()
    _lib.check(_lib.load().b2c_image_stats_target_size(int(W), int(H), C.byref(nw), C.byref(nh)), "b2c_image_stats_target_size")
    return nw.value, nh.value


@torch.no_grad()
def image_stats(images, device=None) -> torch.Tensor:
    """images: uint8 [B,H,W,3] tensor or list of (ragged) uint8 [H,W,3] tensors, RGB as ``np.array(pil_img)`` gives them
    (utils/embedder.py:170).  Returns float64 [B, 22] on the device."""
    if not torch.cuda.is_available():
        raise _lib.B2CError("image_stats needs a CUDA device (sm_100a); there is no CPU fallback")
    imgs = list(images) if not isinstance(images, torch.Tensor) else [images[i] for i in range(images.shape[0])]
    B = len(imgs)
    dev = torch.device(device) if device is not None else (imgs[0].device if B and imgs[0].is_cuda else torch.device("cuda"))
    keep = []
    for im in imgs:
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError(f"expected uint8 [H,W,3], got {im.dtype} {tuple(im.shape)}")
        t = im.to(dev)
        if t.stride(2) != 1 or t.stride(1) != 3:
            t = t.contiguous()
        keep.append(t)
    out = torch.empty(B, _lib.IMG_STATS, dtype=torch.float64, device=dev)
    if B == 0:
        return out
    lib = _lib.load()
    ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in keep])
    Hs = (C.c_int * B)(*[t.shape[0] for t in keep])
    Ws = (C.c_int * B)(*[t.shape[1] for t in keep])
    Ps = (C.c_int * B)(*[t.stride(0) for t in keep])
    need = C.c_size_t()
    _lib.check(lib.b2c_image_stats_workspace_bytes(Hs, Ws, B, C.byref(need)), "b2c_image_stats_workspace_bytes")
    ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    with torch.cuda.device(dev):
        _lib.check(lib.b2c_image_stats(ptrs, Hs, Ws, Ps, B, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr() + off), need.value,
                                       C.c_void_p(_lib.current_stream_ptr())), "b2c_image_stats")
        # no synchronisation: `keep` / `ws` were allocated on the current stream, so torch's caching allocator hands
        # their memory out again only to work that is ordered after the launches above
    return out


def stats_dict(row) -> dict:
    """One row of ``image_stats`` -> {img_stat_*: f32 0-d CPU tensor} as stored in the .pt file (_1:149-161)."""
    r = row.detach().cpu().tolist() if isinstance(row, torch.Tensor) else list(row)
    return {n: torch.tensor(v, dtype=torch.float64).float() for n, v in zip(STAT_NAMES, r)}
