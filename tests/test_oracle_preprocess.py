"""Pins oracle/preprocess_oracle.py: against golden vectors produced by the reference's own
CustomImageDataset + transform (tools/gen_golden.py), against Pillow itself, and — in the build
container — against the live reference."""
import hashlib

import numpy as np
import pytest
from PIL import Image

from conftest import reference_present
from oracle.preprocess_oracle import (CROP_NAMES, crop_geometry, four_crop_preprocess, four_crop_u8, pil_resize_bicubic,
                                      synthetic_image)


def test_golden_digests(golden):
    g = golden("preprocess_ref.npz")
    for k, (W, H) in enumerate(g["sizes"].tolist()):
        img = synthetic_image(k, H, W)
        assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["in_sha256"][k]), "synthetic image generator drifted"
        out = four_crop_preprocess(img, 224)
        assert out.dtype == np.float32 and out.shape == (4, 3, 224, 224)
        assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["sha256"][k]), f"oracle != reference for {(W, H)}"
        geo = crop_geometry(W, H, 224)
        assert [[c["cw"], c["ch"]] for c in geo] == g["crop_sizes"][k].tolist()


def test_golden_full_example(golden):
    g = golden("preprocess_ref.npz")
    u8 = four_crop_u8(g["full_img"], 224)  # [4,R,R,3]
    assert np.array_equal(np.moveaxis(u8[2:4], -1, 1), g["full_crops_u8"])


def test_crop_geometry_golden(golden):
    rows = golden("crop_geometry_ref.npz")["rows"]
    for r in rows.tolist():
        W, H = r[0], r[1]
        geo = crop_geometry(W, H, 224)
        assert [v for c in geo for v in (c["cw"], c["ch"])] == r[2:], (W, H)


@pytest.mark.parametrize("w,h,ow,oh", [(512, 512, 224, 224), (198, 198, 224, 224), (161, 161, 224, 224), (768, 512, 336, 224),
                                       (100, 122, 224, 273), (37, 53, 224, 320), (1024, 256, 896, 224), (224, 224, 224, 224),
                                       (224, 300, 224, 300)])
def test_resize_matches_pillow(w, h, ow, oh):
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(pil_resize_bicubic(img, ow, oh), want)


def test_matches_torchvision_pipeline():
    """The open_clip val transform rebuilt from torchvision (installed on both boxes) on PIL crops."""
    import torch
    from clip_assisted_data_labeling_b200.embedder import CustomImageDataset, _open_clip_val_transform
    tf = _open_clip_val_transform(224)
    ds = CustomImageDataset([], CROP_NAMES, tf)
    rng = np.random.default_rng(3)
    for (W, H) in [(320, 200), (150, 333), (224, 224), (90, 90)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        crops, names = ds.extract_crops(Image.fromarray(img))
        assert names == CROP_NAMES
        want = torch.stack([tf(c) for c in crops]).numpy()
        assert np.array_equal(four_crop_preprocess(img, 224), want)


@pytest.mark.skipif(not reference_present(), reason="reference tree only exists in the build container")
def test_live_reference():
    import torch
    from oracle import reference_shim as rs
    emb = rs.import_reference("utils.embedder")
    enc = emb.CLIP_Encoder("ViT-B-32/openai", device="cpu")
    tf = enc.get_preprocess_transform()
    ds = emb.CustomImageDataset([], CROP_NAMES, tf)
    rng = np.random.default_rng(1)
    for (W, H) in [(257, 131), (64, 700), (512, 512)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        crops, names = ds.extract_crops(Image.fromarray(img))
        want = torch.stack([tf(c) for c in crops]).numpy()
        assert np.array_equal(four_crop_preprocess(img, 224), want)
