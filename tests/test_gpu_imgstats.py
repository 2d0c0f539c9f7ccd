"""-m gpu: K13 (img_stat_* scalars) through the C-ABI against the oracle and the reference's own ImageFeaturizer output
(tests/golden/imgstats_ref.npz).  The integer stages are bit-exact, so the float64 results agree to round-off: 1e-10."""
import numpy as np
import pytest
import torch

from oracle.imgstats_oracle import STAT_NAMES, image_stats_oracle, target_size
from oracle.preprocess_oracle import synthetic_image

pytestmark = pytest.mark.gpu
ATOL = 1e-10


def test_stats_vs_reference_golden(lib, golden):
    from clip_assisted_data_labeling_b200 import imgstats
    g = golden("imgstats_ref.npz")
    assert imgstats.STAT_NAMES == STAT_NAMES == g["names"].tolist()
    imgs = [torch.from_numpy(synthetic_image(k, H, W)) for k, (W, H) in enumerate(g["sizes"].tolist())]
    got = imgstats.image_stats(imgs).cpu().numpy()  # one ragged batch through every resize branch
    np.testing.assert_allclose(got, g["stats"], rtol=0, atol=ATOL)
    for (W, H) in g["sizes"].tolist():
        assert imgstats.target_size(W, H) == target_size(H, W)


def test_batch_tensor_chunking_and_pitch(lib):
    """uint8 [B,H,W,3] batches larger than one internal chunk (64), a non-contiguous view (row pitch > 3W), repeats."""
    from clip_assisted_data_labeling_b200.imgstats import image_stats
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (70, 40, 56, 3), dtype=np.uint8)
    got = image_stats(torch.from_numpy(base).cuda()).cpu().numpy()
    for b in (0, 1, 63, 64, 69):
        np.testing.assert_allclose(got[b], image_stats_oracle(base[b]), rtol=0, atol=ATOL)
    wide = torch.from_numpy(rng.integers(0, 256, (90, 200, 3), dtype=np.uint8)).cuda()
    view = wide[:, 20:150, :]
    got = image_stats([view, view]).cpu().numpy()
    ref = image_stats_oracle(view.cpu().numpy())
    np.testing.assert_allclose(got[0], ref, rtol=0, atol=ATOL)
    assert np.array_equal(got[0], got[1])  # integer accumulation: run-to-run / slot-to-slot identical


@pytest.mark.parametrize("W,H", [(1024, 1024), (1920, 1080), (333, 517), (768, 768), (3000, 2000), (20, 900)])
def test_random_sizes_vs_oracle(lib, W, H):
    from clip_assisted_data_labeling_b200.imgstats import image_stats
    img = synthetic_image(W + H, H, W)
    got = image_stats([torch.from_numpy(img)]).cpu().numpy()[0]
    np.testing.assert_allclose(got, image_stats_oracle(img), rtol=0, atol=ATOL)


def test_flat_and_extreme_images(lib):
    from clip_assisted_data_labeling_b200.imgstats import image_stats
    imgs = [np.zeros((64, 64, 3), np.uint8), np.full((64, 80, 3), 255, np.uint8), np.full((30, 30, 3), 128, np.uint8)]
    imgs[2][::2, ::2] = (255, 0, 0)
    got = image_stats([torch.from_numpy(i) for i in imgs]).cpu().numpy()
    for g_, im in zip(got, imgs):
        np.testing.assert_allclose(g_, image_stats_oracle(im), rtol=0, atol=ATOL)
    assert got[0][20] == pytest.approx(0.0, abs=1e-12) and got[0][21] == 0.0  # entropy / Laplacian variance of a flat image


def test_pt_files_carry_img_stats(lib, tmp_path):
    """Feature_Dataset on the device path stores the 22 scalars ahead of the crops, as f32 0-d tensors (_1:149-161)."""
    import os
    from PIL import Image
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    root = tmp_path / "d"
    root.mkdir()
    img = synthetic_image(3, 120, 160)
    Image.fromarray(img).save(root / "a.png")
    Feature_Dataset(str(root), "ViT-B-32/openai", 2, shuffle_filenames=False, allow_random_init=True).process()
    d = torch.load(os.path.join(root, "a.pt"))["ViT-B-32/openai"]
    keys = list(d.keys())
    assert keys[:22] == STAT_NAMES and keys[22:] == ["centre_crop", "square_padded_crop", "subcrop1", "subcrop2"]
    ref = image_stats_oracle(img)
    for i, n in enumerate(STAT_NAMES):
        assert d[n].dtype == torch.float32 and d[n].dim() == 0 and abs(float(d[n]) - ref[i]) < 1e-6
