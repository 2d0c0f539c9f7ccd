#!/usr/bin/env python
"""Fused-LayerNorm layer loop vs the stand-alone-LayerNorm loop vs the fp32 oracle, per architecture.
    python tools/dev_check_ln.py [arch ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from clip_assisted_data_labeling_b200.vit import VisionTower  # noqa: E402
from oracle import vit_oracle  # noqa: E402

CASES = {"ViT-B-32": ("openai", 9), "ViT-L-14": ("openai", 5), "ViT-H-14": ("laion2b_s32b_b79k", 3), "ViT-L-14-336": ("openai", 1)}


def main():
    archs = sys.argv[1:] or ["ViT-B-32", "ViT-L-14", "ViT-H-14"]
    for arch in archs:
        tag, n = CASES[arch]
        m = vit_oracle.build_visual(arch, tag, seed=0)
        tower = VisionTower(vit_oracle.ARCHS[arch], m.cfg["act"], "cuda")
        tower.load_state_dict(vit_oracle.visual_state_dict(m))
        R = m.cfg["image"]
        px = torch.randn(n, 3, R, R, generator=torch.Generator().manual_seed(1))
        ref = vit_oracle.encode_image_oracle(m, px)
        res = {}
        for fused in (False, True):
            tower.set_fused_ln(fused)
            try:
                got = tower.forward_pixels(px.cuda()).cpu()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(arch, "fused" if fused else "plain", "FAILED:", e, flush=True)
                raise
            cos = torch.nn.functional.cosine_similarity(ref, got, dim=-1).min().item()
            mx = (ref - got).abs().max().item()
            res[fused] = got
            print(f"{arch:14s} {'fused' if fused else 'plain'}: min cos {cos:.6f}  max abs {mx:.2e}  finite {bool(torch.isfinite(got).all())}", flush=True)
        print(f"{arch:14s} fused vs plain: max abs {(res[True] - res[False]).abs().max().item():.2e}", flush=True)


if __name__ == "__main__":
    main()
