"""TEST INFRASTRUCTURE — never imported by the product path (clip_assisted_data_labeling_b200/).

CPU restatement (numpy, integer arithmetic) of what `Image.open(path).convert('RGB')` (utils/embedder.py:167) does to a
baseline JPEG's coefficients inside Pillow's libjpeg-turbo at its defaults:
  * dequantise + 8x8 inverse DCT, the 13-bit fixed-point "islow" transform (jidctint.c: jpeg_idct_islow), post-IDCT
    range limiting through the 1024-entry wrap-around table (jdmaster.c: prepare_range_limit_table);
  * "fancy" triangle-filter chroma upsampling h2v1 / h2v2 with the first/last sample row replicated at the image edges
    (jdsample.c: h2v1_fancy_upsample, h2v2_fancy_upsample; jdmainct.c context rows);
  * YCbCr -> RGB with the SCALEBITS = 16 tables of jdcolor.c (build_ycc_rgb_table / ycc_rgb_convert); grey -> R=G=B.
Pinned against Pillow itself (tests/test_jpeg.py): the real implementation is importable on both boxes, so every test
compares with `PIL.Image.open(...).convert('RGB')` directly; this file exists so that the host-side Huffman decoder can
be checked without a GPU and so that a device mismatch can be localised to a stage."""
import numpy as np

F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
         f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(v):
    """v: int64 [..., 8] -> un-descaled outputs [..., 8] (LL&M, as jidctint.c orders it)."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F["f0_541"]
    tmp2 = z1 + z3 * (-F["f1_847"])
    tmp3 = z1 + z2 * F["f0_765"]
    tmp0 = (v[..., 0] + v[..., 4]) << 13
    tmp1 = (v[..., 0] - v[..., 4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F["f1_175"]
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F["f0_298"], tmp1 * F["f2_053"], tmp2 * F["f3_072"], tmp3 * F["f1_501"]
    z1, z2 = z1 * -F["f0_899"], z2 * -F["f2_562"]
    z3, z4 = z3 * -F["f1_961"] + z5, z4 * -F["f0_390"] + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                     tmp10 - tmp3], axis=-1)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _range_limit(x):
    i = x & 1023
    return np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)


def idct_plane(coefs, qt, bh, bw):
    """coefs int16 [bh*bw*64] (natural order), qt uint16 [64] -> uint8 plane [bh*8, bw*8]."""
    blk = coefs.reshape(bh, bw, 8, 8).astype(np.int64) * qt.reshape(8, 8).astype(np.int64)
    ws = _descale(_idct_1d(blk.swapaxes(-1, -2)), 11).swapaxes(-1, -2)  # pass 1 over columns
    out = _range_limit(_descale(_idct_1d(ws), 18))                      # pass 2 over rows
    return out.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def _h2v1(pl):
    p = pl.astype(np.int32)
    n = p.shape[1]
    out = np.empty((p.shape[0], 2 * n), np.int32)
    out[:, 0] = p[:, 0]
    out[:, 2::2] = (3 * p[:, 1:] + p[:, :-1] + 1) >> 2
    out[:, 1:-1:2] = (3 * p[:, :-1] + p[:, 1:] + 2) >> 2
    out[:, -1] = p[:, -1]
    return out


def _h2v2(pl):
    p = pl.astype(np.int32)
    h, n = p.shape
    up = np.concatenate([p[:1], p[:-1]])   # row above, first row replicated
    dn = np.concatenate([p[1:], p[-1:]])   # row below, last row replicated
    out = np.empty((2 * h, 2 * n), np.int32)
    for par, far in ((0, up), (1, dn)):
        cs = 3 * p + far
        o = np.empty((h, 2 * n), np.int32)
        o[:, 0] = (cs[:, 0] * 4 + 8) >> 4
        o[:, 2::2] = (cs[:, 1:] * 3 + cs[:, :-1] + 8) >> 4
        o[:, 1:-1:2] = (cs[:, :-1] * 3 + cs[:, 1:] + 7) >> 4
        o[:, -1] = (cs[:, -1] * 4 + 7) >> 4
        out[par::2] = o
    return out


def reconstruct(info: dict, coefs: np.ndarray) -> np.ndarray:
    """info: fields of b2c_jpeg_info as a dict; coefs: int16 [coef_count] -> uint8 [H, W, 3]."""
    W, H, nc = info["width"], info["height"], info["ncomp"]
    planes = []
    for c in range(nc):
        bw, bh = info["blocks_w"][c], info["blocks_h"][c]
        off = info["coef_offset"][c]
        pl = idct_plane(coefs[off:off + bw * bh * 64], np.asarray(info["qt"][c], np.uint16), bh, bw)
        planes.append(pl[:info["comp_h"][c], :info["comp_w"][c]])
    y = planes[0].astype(np.int32)
    if nc == 1:
        return np.repeat(planes[0][:, :, None], 3, axis=2)
    hs, vs = info["hs"][0], info["vs"][0]
    ch = []
    for pl in planes[1:]:
        if (hs, vs) == (1, 1):
            u = pl.astype(np.int32)
        elif (hs, vs) == (2, 1):
            u = _h2v1(pl)
        elif (hs, vs) == (2, 2):
            u = _h2v2(pl)
        else:
            raise ValueError("unsupported sampling")
        ch.append(u[:H, :W])
    cb, cr = ch[0] - 128, ch[1] - 128
    r = y + ((91881 * cr + 32768) >> 16)
    g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=2), 0, 255).astype(np.uint8)
