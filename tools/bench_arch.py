"""Secondary configs of BASELINE.json on one GPU (not the headline bench line): 4-crop embedding throughput of any
supported architecture, optionally with the FC regressor scored in the same pass (config 5), with stage shares from the
library's stage timer.  One JSON line per invocation.
    python tools/bench_arch.py ViT-H-14/laion2b_s32b_b79k --fc --batch 128 --steps 4
    python tools/bench_arch.py ViT-L-14-336/openai --batch 64
"""
import argparse
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--fc", action="store_true", help="also score SimpleFC(4*E,[264,128,64],1) on the embeddings (config 5)")
    a = ap.parse_args()
    from bench import load_peaks, synth_batch
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    from clip_assisted_data_labeling_b200.scorer import FCScorer, SimpleFC, embed_and_score
    from clip_assisted_data_labeling_b200.vit_arch import ARCHS, flops_per_crop
    arch = a.model.split("/")[0]
    cfg = ARCHS[arch]
    with contextlib.redirect_stdout(sys.stderr):
        enc = CLIP_Encoder(a.model, device="cuda", seed=0, allow_random_init=True)
    scorer = None
    if a.fc:
        torch.manual_seed(0)
        scorer = FCScorer(SimpleFC(4 * cfg["embed"], [264, 128, 64], 1, [a.model]).eval(), "cuda")
    pool = [synth_batch(a.batch, i).cuda() for i in range(3)]

    def step(x):
        return embed_and_score(enc, scorer, x) if scorer else (enc.encode_images_u8(x), None)

    for i in range(a.warmup):
        step(pool[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        emb, sc = step(pool[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    _lib.prof_enable(True)
    step(pool[0])
    torch.cuda.synchronize()
    rec = _lib.prof_read()
    _lib.prof_enable(False)
    tot = sum(v[0] for v in rec.values())
    peaks = load_peaks()
    ips = a.batch / ms * 1e3
    tf = ips * 4 * flops_per_crop(cfg) / 1e12
    print(json.dumps({"model": a.model, "fc_scoring": bool(scorer), "batch": a.batch, "images_per_s": ips, "ms_per_step": ms,
                      "tflops": tf, "frac_of_sustained_peak": tf / peaks["tf_sustained"],
                      "attn_mode": os.environ.get("B2C_ATTN", "default"),
                      "stage_ms": {k: round(v[0], 3) for k, v in rec.items()},
                      "stage_share": {k: round(v[0] / tot, 4) for k, v in rec.items()},
                      "unit_norm": bool(torch.allclose(emb.norm(dim=-1), torch.ones_like(emb[..., 0]), atol=1e-4)),
                      "score_range": None if sc is None else [float(sc.min()), float(sc.max())]}))


if __name__ == "__main__":
    main()
