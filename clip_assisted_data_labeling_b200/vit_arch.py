"""Architecture table of the open_clip vision towers the reference can name on its CLI
(`_1_embed_with_CLIP.py`:190 ``--models_to_use Arch/pretrained``), and a seeded random initialiser
(no checkpoints are reachable offline; BASELINE.json asks for random-init weights of the named
architecture).  SURVEY.md App. A."""
from __future__ import annotations

import math

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)  # restated by the reference at utils/embedder.py:122
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)  # utils/embedder.py:123
CROP_NAMES = ["centre_crop", "square_padded_crop", "subcrop1", "subcrop2"]  # _1_embed_with_CLIP.py:200

ARCHS = {
    "ViT-B-32": dict(image=224, patch=32, width=768, layers=12, heads=12, mlp=3072, embed=512),
    "ViT-L-14": dict(image=224, patch=14, width=1024, layers=24, heads=16, mlp=4096, embed=768),
    "ViT-L-14-336": dict(image=336, patch=14, width=1024, layers=24, heads=16, mlp=4096, embed=768),
    "ViT-H-14": dict(image=224, patch=14, width=1280, layers=32, heads=16, mlp=5120, embed=1024),
}


def split_model_name(model_name: str) -> tuple[str, str]:
    """'ViT-L-14/openai' -> ('ViT-L-14', 'openai')  (utils/embedder.py:63)."""
    arch, pretrained = model_name.split("/", 2)
    return arch, pretrained


def activation_for(pretrained: str) -> str:
    """open_clip forces QuickGELU for the 'openai' weights; LAION tags use exact GELU."""
    return "quick_gelu" if pretrained == "openai" else "gelu"


def tokens(cfg: dict) -> int:
    g = cfg["image"] // cfg["patch"]
    return g * g + 1


def flops_per_crop(cfg: dict) -> float:
    """Algorithmic FLOPs of one crop: conv1 + L (24 T d^2 + 4 T^2 d) + 2 d E."""
    g = cfg["image"] // cfg["patch"]
    T, d = g * g + 1, cfg["width"]
    return 2.0 * g * g * 3 * cfg["patch"] ** 2 * d + cfg["layers"] * (24.0 * T * d * d + 4.0 * T * T * d) + 2.0 * d * cfg["embed"]


def state_dict_shapes(cfg: dict) -> dict:
    """open_clip ``visual.*`` key -> shape."""
    d, p, T = cfg["width"], cfg["patch"], tokens(cfg)
    s = {
        "class_embedding": (d,), "positional_embedding": (T, d), "proj": (d, cfg["embed"]),
        "conv1.weight": (d, 3, p, p), "ln_pre.weight": (d,), "ln_pre.bias": (d,),
        "ln_post.weight": (d,), "ln_post.bias": (d,),
    }
    for i in range(cfg["layers"]):
        b = f"transformer.resblocks.{i}."
        s.update({
            b + "ln_1.weight": (d,), b + "ln_1.bias": (d,), b + "ln_2.weight": (d,), b + "ln_2.bias": (d,),
            b + "attn.in_proj_weight": (3 * d, d), b + "attn.in_proj_bias": (3 * d,),
            b + "attn.out_proj.weight": (d, d), b + "attn.out_proj.bias": (d,),
            b + "mlp.c_fc.weight": (cfg["mlp"], d), b + "mlp.c_fc.bias": (cfg["mlp"],),
            b + "mlp.c_proj.weight": (d, cfg["mlp"]), b + "mlp.c_proj.bias": (d,),
        })
    return s


def random_state_dict(cfg: dict, seed: int = 0, device="cpu"):
    """Seeded random weights with open_clip's init scales (width^-0.5 embeddings, ~1/sqrt(fan_in)
    linears) and non-trivial LayerNorm affine / biases so every code path carries signal."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = cfg["width"]
    out = {}
    for k, shape in state_dict_shapes(cfg).items():
        if k.endswith("ln_1.weight") or k.endswith("ln_2.weight") or k in ("ln_pre.weight", "ln_post.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif k in ("class_embedding", "positional_embedding", "proj"):
            t = d ** -0.5 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            t = fan_in ** -0.5 * torch.randn(shape, generator=g)
        out[k] = t.to(device)
    return out
