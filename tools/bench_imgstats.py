#!/usr/bin/env python
"""K13 measurement: images/s and achieved HBM GB/s of the img_stat_* pass on 512x512 uint8 images (algorithmic bytes per
image: the 786432 source bytes, read once), next to the reference's ImageFeaturizer on the host
CPU (one core, bounded sample).  One JSON line."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_assisted_data_labeling_b200.imgstats import image_stats  # noqa: E402
from oracle.preprocess_oracle import synthetic_image  # noqa: E402

B = 256
imgs = torch.from_numpy(np.stack([synthetic_image(k % 16, 512, 512) for k in range(B)])).cuda()
for _ in range(3):
    image_stats(imgs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    image_stats(imgs)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
bytes_per_img = 512 * 512 * 3  # the fused kernel reads the source once; the resized image lives in shared memory only
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
out = {"batch": B, "image": "512x512x3 uint8", "ms_per_batch": ms, "images_per_s": B / ms * 1e3,
       "algorithmic_bytes_per_image": bytes_per_img, "GBps": B * bytes_per_img / ms / 1e6,
       "frac_of_hbm_peak": B * bytes_per_img / ms / 1e6 / peaks["hbm_gbs"], "timing": "CUDA events incl. plan upload and the wrapper's stream sync"}
try:
    import cv2  # noqa: F401
    from oracle.imgstats_oracle import image_stats_oracle  # noqa: F401
    sys.path.insert(0, "/root/reference")
    cpu_imgs = imgs[:8].cpu().numpy()
    import cv2 as _cv
    t0 = time.perf_counter()
    for im in cpu_imgs:  # the cv2 calls of utils/image_features.py:60-86 (restated; the reference tree is absent on the GPU box)
        small = _cv.resize(im, (768, 768), interpolation=_cv.INTER_AREA)
        g = _cv.cvtColor(small, _cv.COLOR_BGR2GRAY)
        hsv = _cv.cvtColor(small, _cv.COLOR_BGR2HSV)
        _ = [np.mean(small), np.std(small)] + [f(small[:, :, c]) for c in range(3) for f in (np.mean, np.std)] + [np.mean(g), np.std(g)] + \
            [f(hsv[:, :, c]) for c in range(3) for f in (np.mean, np.std)]
        f64 = small.astype("float")
        rg, yb = np.abs(f64[..., 2] - f64[..., 1]), np.abs(0.5 * (f64[..., 2] + f64[..., 1]) - f64[..., 0])
        _ = [np.mean(rg), np.std(rg), np.mean(yb), np.std(yb)]
        h = _cv.calcHist([g], [0], None, [256], [0, 256]); h /= h.sum(); _ = -np.sum(h * np.log2(h + np.finfo(float).eps))
        _ = np.var(_cv.Laplacian(g, _cv.CV_64F))
    out["cpu_images_per_s_cv2_numpy"] = len(cpu_imgs) / (time.perf_counter() - t0)
    out["cpu_threads_cv2"] = _cv.getNumThreads()
except Exception as e:  # noqa: BLE001
    out["cpu_error"] = str(e)
print(json.dumps(out))
