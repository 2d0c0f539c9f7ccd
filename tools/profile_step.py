"""One warm-up + one timed-shape pass of the hot path for ncu (launch list / full capture).  Not a benchmark:
numbers printed under a profiler are never reported.
    python tools/profile_step.py embed [B]     # ViT-L/14 4-crop step on B images
    python tools/profile_step.py dedup [N]     # duplicate search over N x 768
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def embed(B):
    from bench import synth_batch
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    enc = CLIP_Encoder("ViT-L-14/openai", device="cuda", seed=0, allow_random_init=True)
    imgs = synth_batch(B, 0).cuda()
    for _ in range(2):
        out = enc.encode_images_u8(imgs)
    torch.cuda.synchronize()
    print("ok", tuple(out.shape))


def dedup(N):
    from clip_assisted_data_labeling_b200.dedup import duplicate_pairs
    e = torch.nn.functional.normalize(torch.randn(N, 768, device="cuda"), dim=1).half()
    e[1] = e[0]
    for _ in range(2):
        p, _ = duplicate_pairs(e, 0.96)
    print("ok", len(p))


if __name__ == "__main__":
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    embed(n or 256) if what == "embed" else dedup(n or 200_000)
