import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from PIL import Image
from bench import synth_batch
from clip_assisted_data_labeling_b200 import jpeg
imgs = synth_batch(64, 0).numpy()
datas = []
for i in range(256):
    buf = io.BytesIO(); Image.fromarray(imgs[i % 64]).save(buf, "JPEG", quality=90, subsampling=2); datas.append(buf.getvalue())
items = [jpeg.prepare_file(d) for d in datas]
for _ in range(3):
    jpeg.decode_device(items); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    jpeg.decode_device(items)
torch.cuda.synchronize()
print("decode_device from prepared items: %.2f ms per 256" % ((time.perf_counter() - t0) * 100))
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    jpeg.decode_device(items)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:5000])
