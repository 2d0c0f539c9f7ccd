#!/usr/bin/env python
"""K14 measurement: decoding N synthetic 512x512 baseline JPEGs (quality 90, 4:2:0) to uint8 RGB in HBM.
  pillow_host      Image.open(...).convert('RGB') on a thread pool over all host cores (+ H2D not included)
  entropy_host     the host stage of the device path alone (marker parse + Huffman decode, C-ABI) on the same pool
  reconstruct_dev  the device stage alone (dequantise + IDCT + upsample + colour), CUDA events, coefficients resident
  hybrid_e2e       host stage (packed coefficients) -> pinned gather -> H2D -> device stage, wall clock
  huffman_dev      K14b alone: the Huffman stage on the device, file bytes resident, CUDA events
  device_e2e       marker parse on the host -> pinned gather of the FILES -> H2D -> device Huffman + reconstruction, wall clock
One JSON line.  python tools/bench_jpeg.py [N] [threads]"""
import concurrent.futures as cf
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from bench import synth_batch  # noqa: E402
from clip_assisted_data_labeling_b200 import jpeg  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    noisy = synth_batch(min(n, 64), 0).numpy()
    # photo-like content: the same colour fields at quarter resolution, upsampled, with light sensor-like noise
    rng = np.random.default_rng(0)
    photo = [np.clip(np.asarray(Image.fromarray(im).resize((128, 128)).resize((512, 512), Image.BICUBIC)).astype(np.int16) +
                     rng.normal(0, 3, im.shape), 0, 255).astype(np.uint8) for im in noisy]
    pool = cf.ThreadPoolExecutor(threads)
    for label, imgs, q in (("photo-like (smooth fields + sigma-3 noise), quality 85", photo, 85),
                           ("noise-heavy (the embedding bench's images: sigma-20 noise), quality 90", noisy, 90)):
        run(n, threads, pool, label, imgs, q)


def run(n, threads, pool, label, imgs, q):
    datas = []
    for i in range(n):
        buf = io.BytesIO()
        Image.fromarray(imgs[i % len(imgs)]).save(buf, "JPEG", quality=q, subsampling=2)
        datas.append(buf.getvalue())

    def pil_decode(d):
        return np.asarray(Image.open(io.BytesIO(d)).convert("RGB"))

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            best = min(best, time.perf_counter() - t0)
        return best, out

    t_pil, ref = timed(lambda: list(pool.map(pil_decode, datas)))
    t_ent, items = timed(lambda: list(pool.map(jpeg.entropy_decode, datas)))
    t_pk, pitems = timed(lambda: list(pool.map(jpeg.entropy_decode_packed, datas)))
    t_pil1, _ = timed(lambda: [pil_decode(d) for d in datas[:32]], 2)
    t_ent1, _ = timed(lambda: [jpeg.entropy_decode(d) for d in datas[:32]], 2)
    outs = jpeg.reconstruct(items)
    torch.cuda.synchronize()
    ok = all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(outs[:16], ref[:16]))
    # device stage alone: coefficients resident in HBM, CUDA events around the two launches of a batch
    infos = [it[0] for it in items]
    dflat = torch.cat([it[1] for it in items]).cuda()
    for _ in range(3):
        jpeg.reconstruct_device(infos, dflat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        jpeg.reconstruct_device(infos, dflat)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 10 / 1e3
    # host coefficients -> pinned gather -> H2D -> device stage
    for _ in range(2):
        jpeg.reconstruct(items)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        jpeg.reconstruct(items)
    torch.cuda.synchronize()
    t_dev_h2d = (time.perf_counter() - t0) / 5

    def hybrid():
        its = list(pool.map(jpeg.entropy_decode_packed, datas))
        o = jpeg.reconstruct_packed(its)
        torch.cuda.synchronize()
        return o

    pouts = jpeg.reconstruct_packed(pitems)
    torch.cuda.synchronize()
    ok = ok and all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(pouts[:16], ref[:16]))
    for _ in range(2):
        jpeg.reconstruct_packed(pitems)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        jpeg.reconstruct_packed(pitems)
    torch.cuda.synchronize()
    t_pk_h2d = (time.perf_counter() - t0) / 5

    t_hyb, _ = timed(hybrid)

    # ---- K14b: both stages on the device
    prep = jpeg.prepare_file
    t_prep, hitems = timed(lambda: list(pool.map(prep, datas)))
    hcoefs, hstatus = jpeg.huffman_device(hitems)
    torch.cuda.synchronize()
    ok_h = hstatus.cpu().tolist() == [0] * n and torch.equal(hcoefs.cpu(), torch.cat([it[1] for it in items]))
    for _ in range(2):
        jpeg.huffman_device(hitems)
    torch.cuda.synchronize()
    from clip_assisted_data_labeling_b200 import _lib
    lib = _lib.load()
    _lib.prof_enable(True)
    for _ in range(5):
        jpeg.huffman_device(hitems)
    torch.cuda.synchronize()
    t_huff = prof_other_ms(lib) / 5 / 1e3
    _lib.prof_enable(False)

    def device_e2e():
        its = list(pool.map(prep, datas))
        o, st = jpeg.decode_device(its)
        torch.cuda.synchronize()
        return o

    douts = device_e2e()
    ok_h = ok_h and all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(douts[:16], ref[:16]))
    t_dev_e2e, _ = timed(device_e2e)
    coef_bytes = sum(int(it[0].coef_count) * 2 for it in items)
    out_bytes = sum(int(it[0].width) * int(it[0].height) * 3 for it in items)
    plane_bytes = sum(int(it[0].blocks_w[c]) * int(it[0].blocks_h[c]) * 64 for it in items for c in range(it[0].ncomp))
    print(json.dumps({
        "workload": f"{n} synthetic 512x512 baseline JPEGs, 4:2:0, {label}, {sum(map(len, datas)) / n / 1e3:.0f} KB each",
        "host_threads": threads, "bit_exact_vs_pillow": bool(ok),
        "pillow_host_images_per_s": n / t_pil, "entropy_host_images_per_s": n / t_ent,
        "pillow_1thread_ms_per_image": 1e3 * t_pil1 / 32, "entropy_1thread_ms_per_image": 1e3 * t_ent1 / 32,
        "reconstruct_dev_images_per_s": n / t_dev, "reconstruct_dev_ms_per_batch": 1e3 * t_dev,
        "reconstruct_dev_gb_per_s": (coef_bytes + out_bytes + 2 * plane_bytes) / t_dev / 1e9,
        "h2d_plus_reconstruct_images_per_s": n / t_dev_h2d, "h2d_plus_reconstruct_ms_per_batch": 1e3 * t_dev_h2d,
        "entropy_host_packed_images_per_s": n / t_pk, "packed_h2d_plus_reconstruct_images_per_s": n / t_pk_h2d,
        "hybrid_e2e_images_per_s": n / t_hyb,
        "huffman_dev": {"bit_exact_vs_host_stage_and_pillow": bool(ok_h), "images_per_s": n / t_huff, "ms_per_batch": 1e3 * t_huff,
                        "entropy_bytes_per_s": sum(map(len, datas)) / t_huff, "host_prepare_images_per_s": n / t_prep,
                        "device_e2e_images_per_s": n / t_dev_e2e, "file_bytes_h2d_per_image": sum(map(len, datas)) / n},
        "bytes_per_image": {"coefficients_h2d": coef_bytes / n, "rgb_out": out_bytes / n, "planes_write_then_read": plane_bytes / n,
                            "packed_coefficients_h2d": sum(int(it[1].numel()) for it in pitems) / n},
    }), flush=True)


def prof_other_ms(lib):
    """Milliseconds the library's stage timer recorded (CUDA events around the launches) since prof_enable(True)."""
    from clip_assisted_data_labeling_b200 import _lib
    return float(sum(ms for ms, _ in _lib.prof_read().values()))


if __name__ == "__main__":
    main()
