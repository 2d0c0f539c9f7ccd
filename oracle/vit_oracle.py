"""ORACLE (test infrastructure, never imported by the product path) — plain-PyTorch fp32 restatement of
the vision tower the reference calls through ``open_clip`` at /root/reference/utils/embedder.py:66-73,98.

``open_clip`` (PyPI ``open_clip_torch``) is a third-party dependency the reference neither vendors nor
pins (there is no requirements file; README.md:96 lists it as a TODO) and it is not installed here, so
its published architecture (``open_clip.transformer.VisionTransformer`` / ``ResidualAttentionBlock``,
SURVEY.md App. A) is restated below and anchored on
  * the reference's own call sites (utils/embedder.py:59-100: precision rule, ``.half()``, in-place L2
    normalise),
  * an independent implementation of the same architecture that IS installed:
    ``transformers.CLIPVisionModelWithProjection`` with mapped weights (tests/test_oracle_vit.py).
PARITY UNPINNED: the reference ships no golden embeddings, tests or checkpoints for this path; weights
are seeded random-init of the named architecture (no checkpoints are available offline).
"""
from __future__ import annotations

import math

import torch
from torch import nn

# name -> (image, patch, width, layers, heads, mlp, embed)
ARCHS = {
    "ViT-B-32": dict(image=224, patch=32, width=768, layers=12, heads=12, mlp=3072, embed=512),
    "ViT-L-14": dict(image=224, patch=14, width=1024, layers=24, heads=16, mlp=4096, embed=768),
    "ViT-L-14-336": dict(image=336, patch=14, width=1024, layers=24, heads=16, mlp=4096, embed=768),
    "ViT-H-14": dict(image=224, patch=14, width=1280, layers=32, heads=16, mlp=5120, embed=1024),
}


def act_for(pretrained: str) -> str:
    """open_clip forces QuickGELU for the 'openai' tag; LAION tags use exact GELU."""
    return "quick_gelu" if pretrained == "openai" else "gelu"


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d, heads, mlp, act):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d)
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_2 = nn.LayerNorm(d)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d, mlp))
        self.mlp.add_module("gelu", QuickGELU() if act == "quick_gelu" else nn.GELU())
        self.mlp.add_module("c_proj", nn.Linear(mlp, d))

    def forward(self, x):  # x: [T, n, d] (sequence first, like open_clip's default)
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, d, layers, heads, mlp, act):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(d, heads, mlp, act) for _ in range(layers)])

    def forward(self, x):
        for blk in self.resblocks:
            x = blk(x)
        return x


class VisionTransformer(nn.Module):
    """State-dict keys equal open_clip's ``visual.*`` names."""

    def __init__(self, image, patch, width, layers, heads, mlp, embed, act="quick_gelu"):
        super().__init__()
        self.cfg = dict(image=image, patch=patch, width=width, layers=layers, heads=heads, mlp=mlp, embed=embed, act=act)
        g = image // patch
        scale = width ** -0.5
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(g * g + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, mlp, act)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, embed))

    def forward(self, x):
        x = self.conv1(x)                                   # [n, d, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [n, g*g, d]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIPVisualOnly(nn.Module):
    """Just enough of open_clip.CLIP for the reference wrapper: ``.visual`` + ``encode_image``."""

    def __init__(self, visual):
        super().__init__()
        self.visual = visual

    def encode_image(self, image, normalize: bool = False):
        f = self.visual(image)
        return torch.nn.functional.normalize(f, dim=-1) if normalize else f


def build_visual(arch: str, pretrained: str = "openai", seed: int = 0, perturb: bool = True) -> VisionTransformer:
    """Seeded random init of the named architecture.  With ``perturb`` every LayerNorm gamma/beta and
    every bias gets a seeded N(0, 0.05..0.1) offset so affine and bias code paths are exercised (PyTorch's
    default init has gamma=1, beta=0, MHA biases 0) — SURVEY.md §8c."""
    cfg = ARCHS[arch]
    gen = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        m = VisionTransformer(act=act_for(pretrained), **cfg)
    if perturb:
        with torch.no_grad():
            for name, p in m.named_parameters():
                if name.endswith("bias") or ".ln_" in name or name.startswith("ln_"):
                    sd = 0.1 if "ln_" in name else 0.05
                    p.add_(sd * torch.randn(p.shape, generator=gen))
    return m.eval()


def visual_state_dict(m: VisionTransformer) -> dict:
    """fp32 CPU state dict with open_clip's visual key names (no 'visual.' prefix)."""
    return {k: v.detach().float().cpu().contiguous() for k, v in m.state_dict().items()}


@torch.no_grad()
def encode_image_oracle(m: VisionTransformer, pixels: torch.Tensor) -> torch.Tensor:
    """Reference CPU semantics of CLIP_Encoder.encode_image (embedder.py:94-100, precision 'fp32'):
    model.encode_image then in-place division by the L2 norm over the last dim (no eps)."""
    f = m(pixels.float())
    f /= f.norm(dim=-1, keepdim=True)
    return f


def flops_per_crop(arch: str) -> float:
    """Algorithmic FLOPs per crop: conv1 + L*(24 T d^2 + 4 T^2 d) + 2 d E."""
    c = ARCHS[arch]
    g = c["image"] // c["patch"]
    T = g * g + 1
    d = c["width"]
    return 2.0 * g * g * 3 * c["patch"] ** 2 * d + c["layers"] * (24.0 * T * d * d + 4.0 * T * T * d) + 2.0 * d * c["embed"]


def to_hf_clip(m: VisionTransformer):
    """Independent second statement: the same weights loaded into transformers' CLIP vision tower."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    c = m.cfg
    cfg = CLIPVisionConfig(hidden_size=c["width"], intermediate_size=c["mlp"], num_hidden_layers=c["layers"],
                           num_attention_heads=c["heads"], image_size=c["image"], patch_size=c["patch"],
                           projection_dim=c["embed"], hidden_act="quick_gelu" if c["act"] == "quick_gelu" else "gelu",
                           layer_norm_eps=1e-5, attn_implementation="eager")
    hf = CLIPVisionModelWithProjection(cfg).eval()
    sd = m.state_dict()
    out = {}
    out["vision_model.embeddings.patch_embedding.weight"] = sd["conv1.weight"]
    out["vision_model.embeddings.class_embedding"] = sd["class_embedding"]
    out["vision_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    out["vision_model.pre_layrnorm.weight"] = sd["ln_pre.weight"]
    out["vision_model.pre_layrnorm.bias"] = sd["ln_pre.bias"]
    out["vision_model.post_layernorm.weight"] = sd["ln_post.weight"]
    out["vision_model.post_layernorm.bias"] = sd["ln_post.bias"]
    out["visual_projection.weight"] = sd["proj"].t().contiguous()
    for i in range(c["layers"]):
        p = f"transformer.resblocks.{i}."
        q = f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[p + "attn.in_proj_weight"].chunk(3)
        bq, bk, bv = sd[p + "attn.in_proj_bias"].chunk(3)
        for nm, w, b in (("q_proj", wq, bq), ("k_proj", wk, bk), ("v_proj", wv, bv)):
            out[q + f"self_attn.{nm}.weight"] = w
            out[q + f"self_attn.{nm}.bias"] = b
        out[q + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        out[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        out[q + "layer_norm1.weight"] = sd[p + "ln_1.weight"]
        out[q + "layer_norm1.bias"] = sd[p + "ln_1.bias"]
        out[q + "layer_norm2.weight"] = sd[p + "ln_2.weight"]
        out[q + "layer_norm2.bias"] = sd[p + "ln_2.bias"]
        out[q + "mlp.fc1.weight"] = sd[p + "mlp.c_fc.weight"]
        out[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.bias"]
        out[q + "mlp.fc2.weight"] = sd[p + "mlp.c_proj.weight"]
        out[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.bias"]
    missing, unexpected = hf.load_state_dict(out, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return hf
