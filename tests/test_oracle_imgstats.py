"""Pins oracle/imgstats_oracle.py: bit-exact against cv2 itself (resize INTER_AREA in all its branches, cvtColor,
Laplacian; with IPP on and off) and to 1e-12 against the 22 statistics the reference's own ImageFeaturizer produced
(tests/golden/imgstats_ref.npz, tools/gen_golden.py gen_imgstats)."""
import numpy as np
import pytest

from oracle.imgstats_oracle import (STAT_NAMES, bgr2gray, bgr2hsv, image_stats_oracle, laplacian_cross, resize_inter_area,
                                    resize_mode, target_size)
from oracle.preprocess_oracle import synthetic_image

cv2 = pytest.importorskip("cv2")


def test_stats_vs_reference_golden(golden):
    g = golden("imgstats_ref.npz")
    assert g["names"].tolist() == STAT_NAMES
    modes = set()
    for k, (W, H) in enumerate(g["sizes"].tolist()):
        nw, nh = target_size(H, W)
        modes.add(resize_mode(W, H, nw, nh)[0])
        got = image_stats_oracle(synthetic_image(k, H, W))
        np.testing.assert_allclose(got, g["stats"][k], rtol=0, atol=1e-12, err_msg=f"{W}x{H}")
    assert modes == {"fast", "area", "linear"}


@pytest.mark.parametrize("ipp", [False, True])
def test_resize_bit_exact_vs_cv2(ipp):
    cv2.ipp.setUseIPP(ipp)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(5)
    sizes = [(640, 480, 665, 886), (100, 1000, 2428, 242), (512, 512, 768, 768), (1536, 1536, 768, 768), (2304, 2304, 768, 768),
             (1000, 1000, 768, 768), (1200, 900, 665, 886), (300, 200, 627, 940), (1537, 1536, 767, 768), (64, 64, 768, 768)]
    for _ in range(6):
        sw, sh = (int(v) for v in rng.integers(20, 1800, 2))
        sizes.append((sw, sh) + target_size(sh, sw))
    for sw, sh, dw, dh in sizes:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(resize_inter_area(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)), (sw, sh, dw, dh)
    cv2.ipp.setUseIPP(True)


def test_colour_and_laplacian_bit_exact_vs_cv2():
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    img[:8, :8] = 0
    img[8:16, :8] = 255
    img[16:24, :8, 0] = img[16:24, :8, 1]  # ties between channels
    assert np.array_equal(bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(bgr2hsv(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV))
    g = bgr2gray(img)
    assert np.array_equal(laplacian_cross(g).astype(np.float64), cv2.Laplacian(g, cv2.CV_64F))


def test_target_size_swaps_like_the_reference():
    assert target_size(512, 512) == (768, 768)
    assert target_size(512, 768) == (627, 940)  # H=512, W=768: the reference's `w,h = shape[:2]` makes the WIDTH smaller
