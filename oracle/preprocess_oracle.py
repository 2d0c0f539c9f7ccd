"""ORACLE (test infrastructure, never imported by the product path) — CPU restatement of the
reference's 4-crop extraction and per-crop preprocessing.

Follows:
  * crop geometry ............ /root/reference/utils/embedder.py:184-251 (extract_crops) with
                               torchvision.transforms.CenterCrop rounding (int(round((H - s) / 2.0)))
  * per-crop transform ....... /root/reference/utils/embedder.py:90-92,173 = open_clip's val transform
                               (third-party, un-vendored, version unpinned by the reference):
                               Resize(R, BICUBIC) -> CenterCrop(R) -> ToTensor -> Normalize(mean, std)
                               (constants restated by the reference at utils/embedder.py:122-123)
  * Resize on a PIL image .... Pillow 12.2 ImagingResample (src/libImaging/Resample.c): separable
                               a=-0.5 bicubic, support scaled by the down-scale factor, coefficients
                               quantised to 22-bit fixed point, horizontal pass first with a uint8
                               intermediate, then the vertical pass.

Pinned by tests/test_oracle_preprocess.py against (a) PIL.Image.resize itself, (b) the reference's own
CustomImageDataset + torchvision transform imported from /root/reference (fixtures under tests/golden/).
Pure numpy: float64 arithmetic in the same operation order as the C code (no FMA contraction).
"""
from __future__ import annotations

import math

import numpy as np

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)
CROP_NAMES = ["centre_crop", "square_padded_crop", "subcrop1", "subcrop2"]
PRECISION_BITS = 32 - 8 - 2


def _round_half_even_div2(a: int) -> int:
    """int(round(a / 2.0)) for a >= 0 (Python's round is banker's rounding)."""
    return int(round(a / 2.0))


def resized_size(w: int, h: int, R: int) -> tuple[int, int]:
    """torchvision Resize(R) with an int size: shorter side -> R, longer = int(R * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = R, int(R * long / short)
    return (new_short, new_long) if w <= h else (new_long, new_short)


def crop_geometry(W: int, H: int, R: int) -> list[dict]:
    """The 4 crops of a W x H image as canvases: pixel (x, y) of a cw x ch canvas is image pixel
    (x + dx, y + dy) when inside the image, black otherwise.  embedder.py:196-247."""
    crops = []
    # centre_crop: CenterCrop(min(W, H))                                   embedder.py:196-202
    s = min(W, H)
    top = _round_half_even_div2(H - s)
    left = _round_half_even_div2(W - s)
    crops.append(dict(cw=s, ch=s, dx=left, dy=top))
    # square_padded_crop: black max(W,H)^2 canvas, image pasted centred     embedder.py:204-212
    S = max(W, H)
    start_h = (S - H) // 2
    start_w = (S - W) // 2
    crops.append(dict(cw=S, ch=S, dx=-start_w, dy=-start_h))
    # subcrops                                                              embedder.py:215-247
    s1 = int((W * H * 0.15) ** 0.5)
    s2 = int((W * H * 0.1) ** 0.5)
    if W >= H:
        centers = [(W // 4, H // 2), (W // 4 * 3, H // 2)]
    else:
        centers = [(W // 2, H // 4), (W // 2, H // 4 * 3)]
    for (cx, cy), sz in zip(centers, (s1, s2)):
        left = max(0, cx - sz // 2)
        top = max(0, cy - sz // 2)
        right = min(W, left + sz)
        bottom = min(H, top + sz)
        cw, ch = right - left, bottom - top
        if cw > 0 and ch > 0:
            crops.append(dict(cw=cw, ch=ch, dx=left, dy=top))
        else:
            crops.append(dict(cw=0, ch=0, dx=0, dy=0))  # dropped by the reference (243-247)
    for c in crops:
        if c["cw"] > 0:
            ow, oh = resized_size(c["cw"], c["ch"], R)
            c.update(out_w=ow, out_h=oh, off_x=_round_half_even_div2(ow - R), off_y=_round_half_even_div2(oh - R))
        else:
            c.update(out_w=0, out_h=0, off_x=0, off_y=0)
    return crops


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for the full-image box.
    Returns (xmin[out], count[out], coef[out, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        x = np.arange(n, dtype=np.float64)
        w = _bicubic((x + lo - center + 0.5) * ss)
        ww = 0.0
        for v in w:  # sequential sum like the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        q = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64),
                     (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))  # C (int) cast truncates toward 0
        xmin[xx], cnt[xx] = lo, n
        coef[xx, :n] = q
    return xmin, cnt, coef


def _resample_axis_last(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample the last-but-one axis... img: [rows, in, C] uint8 -> [rows, out, C] uint8 along axis 1."""
    in_size = img.shape[1]
    xmin, cnt, coef = pil_bicubic_coeffs(in_size, out_size)
    out = np.empty((img.shape[0], out_size, img.shape[2]), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = cnt[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, xmin[xx]:xmin[xx] + n, :], coef[xx, :n].astype(np.int64),
                                                         axes=([1], [0]))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BICUBIC) on an HWC uint8 array: horizontal pass, then vertical."""
    h, w = img.shape[:2]
    if (w, h) == (out_w, out_h):
        return img.copy()
    tmp = _resample_axis_last(img, out_w) if out_w != w else img
    if out_h != h:
        tmp = _resample_axis_last(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(tmp)


def render_canvas(img: np.ndarray, crop: dict) -> np.ndarray:
    """Materialise the crop canvas (black where it leaves the image)."""
    H, W = img.shape[:2]
    cw, ch, dx, dy = crop["cw"], crop["ch"], crop["dx"], crop["dy"]
    canvas = np.zeros((ch, cw, 3), np.uint8)
    x0, y0 = max(0, -dx), max(0, -dy)
    x1, y1 = min(cw, W - dx), min(ch, H - dy)
    canvas[y0:y1, x0:x1] = img[y0 + dy:y1 + dy, x0 + dx:x1 + dx]
    return canvas


def four_crop_u8(img: np.ndarray, R: int) -> np.ndarray:
    """uint8 [4, R, R, 3]: the 4 crops after Resize(R) + CenterCrop(R), before ToTensor/Normalize."""
    H, W = img.shape[:2]
    out = np.zeros((4, R, R, 3), np.uint8)
    for i, c in enumerate(crop_geometry(W, H, R)):
        if c["cw"] == 0:
            continue
        rs = pil_resize_bicubic(render_canvas(img, c), c["out_w"], c["out_h"])
        out[i] = rs[c["off_y"]:c["off_y"] + R, c["off_x"]:c["off_x"] + R]
    return out


def normalize_f32(u8: np.ndarray, mean=OPENAI_MEAN, std=OPENAI_STD) -> np.ndarray:
    """ToTensor (u8 -> f32 / 255, HWC -> CHW) + Normalize((x - mean) / std), all in float32."""
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(np.moveaxis(x, -1, -3))


def four_crop_preprocess(img: np.ndarray, R: int, mean=OPENAI_MEAN, std=OPENAI_STD) -> np.ndarray:
    """f32 [4, 3, R, R] == torch.stack([preprocess(c) for c in raw_crops]) (embedder.py:173).
    A crop the reference cannot produce (zero area: W*H < 10, where embedder.py:247 itself raises) is all zeros."""
    out = normalize_f32(four_crop_u8(img, R), mean, std)
    for i, c in enumerate(crop_geometry(img.shape[1], img.shape[0], R)):
        if c["cw"] == 0:
            out[i] = 0.0
    return out


def synthetic_image(k: int, H: int = 512, W: int = 512) -> np.ndarray:
    """SURVEY.md §8d config-1 image k: low-frequency 3-channel sinusoid field + N(0, 20^2) noise."""
    rng = np.random.default_rng(1000 + k)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.empty((H, W, 3), np.float64)
    for c in range(3):
        fx, fy = rng.uniform(0.5, 4.0, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        img[..., c] = 128 + 90 * np.sin(2 * np.pi * fx * xx / W + ph[0]) * np.cos(2 * np.pi * fy * yy / H + ph[1])
    img += rng.normal(0, 20, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
