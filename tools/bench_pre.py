"""K0 alone: time b2c_preprocess_4crop over a batch of synthetic images (CUDA events, inputs larger than L2) and
check it bit for bit against the numpy oracle on a few of them.
    python tools/bench_pre.py [B] [side] [R] [patch]
Environment knobs read by the library: B2C_PRE_TR (output rows per band), B2C_PRE_SMEM_KB (tile budget)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from clip_assisted_data_labeling_b200.vit import preprocess_u8
    from oracle.preprocess_oracle import four_crop_preprocess
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    R = int(sys.argv[3]) if len(sys.argv) > 3 else 224
    patch = int(sys.argv[4]) if len(sys.argv) > 4 else 14
    g = torch.Generator(device="cuda").manual_seed(0)
    pools = [torch.randint(0, 256, (B, side, side, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3)]

    class Owner:  # keeps the workspace between calls like VisionTower does
        pass
    own = Owner()
    for p in pools:
        out = preprocess_u8(p, R, patch, "patch", cache_owner=own)
    torch.cuda.synchronize()
    reps = 12
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        out = preprocess_u8(pools[i % 3], R, patch, "patch", cache_owner=own)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    med = ms[len(ms) // 2]
    nchw = preprocess_u8(pools[0][:3], R, patch, "nchw").cpu().numpy()
    ok = all(np.array_equal(nchw[k], four_crop_preprocess(pools[0][k].cpu().numpy(), R)) for k in range(3))
    gsz = R // patch
    Kp = (3 * patch * patch + 63) // 64 * 64
    alg = B * (side * side * 3 + 4 * gsz * gsz * Kp * 2)
    print(json.dumps({"B": B, "side": side, "R": R, "patch": patch, "ms_median": med, "ms_min": ms[0],
                      "us_per_image": med * 1e3 / B, "algorithmic_GBps": alg / med / 1e6, "bit_exact_vs_oracle": ok,
                      "TR": os.environ.get("B2C_PRE_TR"), "smem_kb": os.environ.get("B2C_PRE_SMEM_KB")}))


if __name__ == "__main__":
    main()
