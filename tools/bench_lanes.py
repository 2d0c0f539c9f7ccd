#!/usr/bin/env python
"""Device-resident images/s of the ViT-L/14 4-crop step for lanes in {1,2,3,4} (b2c_vit_set_lanes) and a few batch
sizes: which split of a pass into independent sub-batches on separate streams hides the most non-GEMM time.
    python tools/bench_lanes.py [--batches 256,512] [--lanes 1,2,3,4] [--steps 6]
One JSON line per (batch, lanes)."""
import argparse
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256")
    ap.add_argument("--lanes", default="1,2,3,4")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--model", default="ViT-L-14/openai")
    ap.add_argument("--fused", default="1", help="comma list of 0/1: LayerNorm fused into the GEMMs (b2c_vit_set_fused_ln)")
    a = ap.parse_args()
    import torch
    from bench import synth_batch
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    with contextlib.redirect_stdout(sys.stderr):
        enc = CLIP_Encoder(a.model, device="cuda", seed=0, allow_random_init=True)
    for B in [int(x) for x in a.batches.split(",")]:
        pool = [synth_batch(B, i).cuda() for i in range(3)]
        base = None
        for fused, lanes in [(int(f), int(x)) for f in a.fused.split(",") for x in a.lanes.split(",")]:
            enc.model.set_fused_ln(bool(fused))
            enc.model.set_lanes(lanes)
            for i in range(3):
                out = enc.encode_images_u8(pool[i % 3])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                out = enc.encode_images_u8(pool[i % 3])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            if base is None:
                base = out.clone()
            print(json.dumps({"model": a.model, "batch": B, "fused_ln": fused, "lanes": lanes, "ms_per_step": ms, "images_per_s": B / ms * 1e3,
                              "max_abs_vs_first": float((out - base).abs().max())}), flush=True)


if __name__ == "__main__":
    main()
