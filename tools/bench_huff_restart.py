"""K14b alone with and without restart markers: 256 synthetic 512x512 4:2:0 files per variant, CUDA-event time of the
device Huffman stage (library stage timer) and a bit-exact check against the host stage.  python tools/bench_huff_restart.py"""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from PIL import Image
from bench import synth_batch
from clip_assisted_data_labeling_b200 import jpeg, _lib
imgs = synth_batch(64, 0).numpy()
for label, kw in (("no restarts", {}), ("restart every MCU row", {"restart_marker_rows": 1}), ("restart every 4 MCUs", {"restart_marker_blocks": 4}),
                  ("restart every MCU", {"restart_marker_blocks": 1})):
    datas = []
    for i in range(256):
        buf = io.BytesIO(); Image.fromarray(imgs[i % 64]).save(buf, "JPEG", quality=90, subsampling=2, **kw); datas.append(buf.getvalue())
    items = [jpeg.prepare_file(d) for d in datas]
    coefs, status = jpeg.huffman_device(items)
    want = torch.cat([jpeg.entropy_decode(d)[1] for d in datas[:8]])
    ok = status.cpu().tolist() == [0] * 256 and torch.equal(coefs[:want.numel()].cpu(), want)
    for _ in range(2):
        jpeg.huffman_device(items)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(5):
        jpeg.huffman_device(items)
    torch.cuda.synchronize()
    ms = sum(v for v, _ in _lib.prof_read().values()) / 5
    _lib.prof_enable(False)
    print(f"{label:24s} ok={ok} {ms:8.3f} ms per 256 images  ({256 / ms * 1e3:9.0f} images/s), {sum(map(len, datas)) / 256 / 1e3:.0f} KB per file", flush=True)
