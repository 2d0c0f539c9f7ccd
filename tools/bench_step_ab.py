"""Whole-step A/B of attention kernel variants in ONE process (alternating, so box-to-box and thermal drift cancel):
    python tools/bench_step_ab.py [--model ViT-L-14/openai] [--batch 256] [--rounds 4] [--steps 4] --vars -1,0,1,5
-1 = v4, n >= 0 = v5 variant n (b2c_debug_set_attn5)."""
import argparse, contextlib, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ViT-L-14/openai")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--vars", default="-1,0,1")
a = ap.parse_args()
from bench import synth_batch
from clip_assisted_data_labeling_b200 import _lib
from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
lib = _lib.load()
with contextlib.redirect_stdout(sys.stderr):
    enc = CLIP_Encoder(a.model, device="cuda", seed=0, allow_random_init=True)
pool = [synth_batch(a.batch, i, device="cuda") for i in range(3)]
vs = [int(v) for v in a.vars.split(",")]
ref = None
res = {v: [] for v in vs}
err = {}
for rnd in range(a.rounds + 1):
    for v in vs:
        lib.b2c_debug_set_attn5(v)
        out = enc.encode_images_u8(pool[0])
        if ref is None:
            ref = out.clone()
        err[v] = float((out - ref).abs().max())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            enc.encode_images_u8(pool[i % 3])
        e1.record()
        torch.cuda.synchronize()
        if rnd > 0:
            res[v].append(e0.elapsed_time(e1) / a.steps)
for v in vs:
    ms = sorted(res[v])
    print(json.dumps({"model": a.model, "batch": a.batch, "attn": "v4" if v < 0 else f"v5:{v}", "ms_per_step_median": ms[len(ms) // 2],
                      "ms_min": ms[0], "ms_max": ms[-1], "images_per_s": a.batch / ms[len(ms) // 2] * 1e3, "max_abs_vs_first": err[v]}), flush=True)
