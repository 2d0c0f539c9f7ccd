#!/usr/bin/env python
"""K11 measurement: achieved HBM bandwidth of the context-score kernel (algorithmic bytes = N*E*sizeof(elem) + 4N per
launch), exact top-k time, and one greedy diversity ordering — CUDA events on the launching stream, inputs larger than L2.
One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_assisted_data_labeling_b200.similar import context_scores, diversity_order, topk_smallest  # noqa: E402


def timed(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def main():
    n, E = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 768
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out = {"n": n, "E": E, "hbm_peak_gbs": peaks["hbm_gbs"]}
    ctx = torch.randn(E, device="cuda")
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        emb = torch.randn(n, E, device="cuda").to(dt)
        sc = torch.empty(n, device="cuda")
        for measure in ("l2", "cosine"):
            ms = timed(lambda: context_scores(emb, ctx, measure, out=sc))
            bytes_ = n * E * emb.element_size() + 4 * n
            out[f"scores_{name}_{measure}"] = {"ms": ms, "GBps": bytes_ / ms / 1e6, "frac_of_hbm_peak": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}
        if name == "f32":
            for k in (30, 1000):
                out[f"topk_{k}_ms"] = timed(lambda: topk_smallest(sc, k))
    nd, steps, S = min(n, 100_000), 500, 100
    smp = torch.randint(0, nd, (steps, S)).tolist()
    e2 = torch.nn.functional.normalize(torch.randn(nd, E, device="cuda"), dim=1)
    out["diversity_100k_500steps_ms"] = timed(lambda: diversity_order(e2, smp), warm=1, it=2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
