"""Host-side wrapper of the b2c_vit handle (include/b2c.h): owns the handle, the workspace tensors
(torch allocations — "caller owns all tensors") and marshals pointers to the C-ABI on torch's current
stream.  This is the object CLIP_Encoder.model holds where the reference holds an open_clip model
(utils/embedder.py:66-74)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .vit_arch import OPENAI_MEAN, OPENAI_STD, tokens

_DTYPES = {torch.float32: _lib.B2C_F32, torch.float16: _lib.B2C_F16, torch.bfloat16: _lib.B2C_BF16}


def _require_cuda(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.B2CError("the B200 path needs a CUDA device (sm_100a); there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class VisionTower:
    """open_clip VisionTransformer forward + L2 normalise on sm_100a kernels."""

    def __init__(self, cfg: dict, act: str = "quick_gelu", device=None):
        self.cfg = dict(cfg)
        self.act = act
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        c = _lib.VitCfg(cfg["image"], cfg["patch"], cfg["width"], cfg["layers"], cfg["heads"], cfg["mlp"], cfg["embed"],
                        _lib.ACT_GELU if act == "gelu" else _lib.ACT_QUICK_GELU)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2c_vit_create(C.byref(c), C.byref(h)), "b2c_vit_create")
        self._h = h
        self.T = tokens(cfg)
        self.G2 = self.T - 1
        self.Kp = (3 * cfg["patch"] ** 2 + 63) // 64 * 64
        self._ws = None
        self._pre_ws = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b2c_vit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: dict) -> None:
        """sd: open_clip visual state dict (keys with or without the 'visual.' prefix; text-tower keys ignored)."""
        from .vit_arch import state_dict_shapes
        expected = state_dict_shapes(self.cfg)
        if any(k.startswith("visual.") for k in sd):  # a full CLIP checkpoint: keep the tower only
            sd = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
        with torch.cuda.device(self.device):
            for k, v in sd.items():
                if k not in expected:
                    continue
                if tuple(v.shape) != tuple(expected[k]):
                    raise ValueError(f"weight {k}: shape {tuple(v.shape)} != expected {tuple(expected[k])}")
                t = v.detach()
                if t.dtype not in _DTYPES:
                    t = t.float()
                t = t.to(self.device).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
                _lib.check(self.lib.b2c_vit_set_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), _DTYPES[t.dtype], shape,
                                                       max(t.dim(), 1)), f"b2c_vit_set_weight({k})")
            _lib.check(self.lib.b2c_vit_ready(self._h), "b2c_vit_ready")

    def set_lanes(self, lanes: int) -> None:
        """Number of independent sub-batches (own stream each) a pass is split into; see include/b2c.h."""
        _lib.check(self.lib.b2c_vit_set_lanes(self._h, int(lanes)), "b2c_vit_set_lanes")

    def set_cls_only_last_block(self, on: bool) -> None:
        """Opt-in: the last block evaluates only the class-token row (what ln_post / proj read); see include/b2c.h."""
        _lib.check(self.lib.b2c_vit_set_cls_only_last_block(self._h, 1 if on else 0), "b2c_vit_set_cls_only_last_block")

    def set_graph(self, on: bool) -> None:
        """Opt-in: replay repeated forward calls (same buffers / size / switches) as one CUDA graph; see include/b2c.h."""
        _lib.check(self.lib.b2c_vit_set_graph(self._h, 1 if on else 0), "b2c_vit_set_graph")

    def set_fused_ln(self, on: bool) -> None:
        """LayerNorm folded into the GEMMs on either side of it (default) or stand-alone kernels; see include/b2c.h."""
        _lib.check(self.lib.b2c_vit_set_fused_ln(self._h, 1 if on else 0), "b2c_vit_set_fused_ln")

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, n: int) -> torch.Tensor:
        need = C.c_size_t()
        _lib.check(self.lib.b2c_vit_workspace_bytes(self._h, int(n), C.byref(need)), "b2c_vit_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _aligned_ptr(t: torch.Tensor, a: int = 1024) -> int:
        return (t.data_ptr() + a - 1) // a * a

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward_pixels(self, pixels: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """pixels [n,3,R,R] (f32/f16/bf16, normalised) -> f32 [n,E] unit-norm. utils/embedder.py:94-100.
        ``out``: optional destination f32 [n,E] on the device (stable buffers let set_graph(True) replay the pass)."""
        R = self.cfg["image"]
        if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, R, R):
            raise ValueError(f"expected [n,3,{R},{R}], got {tuple(pixels.shape)}")
        if pixels.dtype not in _DTYPES:
            pixels = pixels.float()
        pixels = pixels.to(self.device).contiguous()
        n = pixels.shape[0]
        if out is None:
            out = torch.empty(n, self.cfg["embed"], dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (n, self.cfg["embed"]) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != pixels.device:
            raise ValueError("out must be a contiguous f32 [n, E] tensor on the tower's device")
        if n == 0:
            return out
        with torch.cuda.device(self.device):
            ws = self._workspace(n)
            p = self._aligned_ptr(ws)
            _lib.check(self.lib.b2c_vit_forward_pixels(self._h, C.c_void_p(pixels.data_ptr()), _DTYPES[pixels.dtype], n,
                                                       C.c_void_p(out.data_ptr()), C.c_void_p(p),
                                                       ws.numel() - (p - ws.data_ptr()), C.c_void_p(_lib.current_stream_ptr())),
                       "b2c_vit_forward_pixels")
        return out

    @torch.no_grad()
    def forward_patches(self, patches: torch.Tensor) -> torch.Tensor:
        """patches bf16 [n, g*g, Kp] (from preprocess_u8) -> f32 [n,E] unit-norm."""
        n = patches.shape[0]
        assert patches.dtype == torch.bfloat16 and patches.is_contiguous() and tuple(patches.shape[1:]) == (self.G2, self.Kp)
        out = torch.empty(n, self.cfg["embed"], dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        with torch.cuda.device(self.device):
            ws = self._workspace(n)
            p = self._aligned_ptr(ws)
            _lib.check(self.lib.b2c_vit_forward_patches(self._h, C.c_void_p(patches.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                        C.c_void_p(p), ws.numel() - (p - ws.data_ptr()),
                                                        C.c_void_p(_lib.current_stream_ptr())),
                       "b2c_vit_forward_patches")
        return out

    # ------------------------------------------------------------------ fused u8 path
    @torch.no_grad()
    def preprocess_u8(self, images, layout: str = "patch", mean=OPENAI_MEAN, std=OPENAI_STD) -> torch.Tensor:
        """images: uint8 [B,H,W,3] tensor or list of uint8 [H_i,W_i,3] tensors (device).  Returns the 4 crops
        [centre_crop, square_padded_crop, subcrop1, subcrop2] of every image after Resize/CenterCrop/ToTensor/
        Normalize: layout 'nchw' -> f32 [B,4,3,R,R] (bit-identical to the reference's CPU pipeline),
        'patch' -> bf16 [B*4, g*g, Kp] (the patch-embed GEMM operand)."""
        return preprocess_u8(images, self.cfg["image"], self.cfg["patch"], layout, mean, std, self.device, self)

    @torch.no_grad()
    def encode_u8(self, images, mean=OPENAI_MEAN, std=OPENAI_STD) -> torch.Tensor:
        """uint8 images -> f32 [B,4,E] unit-norm embeddings of the 4 crops (extract_crops + preprocess + encode_image)."""
        patches = self.preprocess_u8(images, "patch", mean, std)
        out = self.forward_patches(patches)
        return out.view(-1, 4, self.cfg["embed"])


def preprocess_u8(images, R: int, patch: int, layout: str = "nchw", mean=OPENAI_MEAN, std=OPENAI_STD, device=None,
                  cache_owner=None) -> torch.Tensor:
    """Functional form of VisionTower.preprocess_u8 (no tower needed for the 'nchw' layout)."""
    lib = _lib.load()
    dev = _require_cuda(device)
    if isinstance(images, torch.Tensor):
        if images.dim() != 4 or images.shape[-1] != 3 or images.dtype != torch.uint8:
            raise ValueError("expected uint8 [B,H,W,3]")
        images = images.to(dev)
        if images.stride(-1) != 1 or images.stride(-2) != 3:
            images = images.contiguous()
        B = images.shape[0]
        ptrs = [images.data_ptr() + i * images.stride(0) for i in range(B)]
        Hs, Ws, pitches = [images.shape[1]] * B, [images.shape[2]] * B, [images.stride(1)] * B
        keep = images
    else:
        keep = [im.to(dev).contiguous() for im in images]
        for im in keep:
            if im.dim() != 3 or im.shape[-1] != 3 or im.dtype != torch.uint8:
                raise ValueError("expected a list of uint8 [H,W,3] tensors")
        B = len(keep)
        ptrs = [im.data_ptr() for im in keep]
        Hs, Ws, pitches = [im.shape[0] for im in keep], [im.shape[1] for im in keep], [im.shape[1] * 3 for im in keep]
    g = R // patch
    Kp = (3 * patch * patch + 63) // 64 * 64
    if layout == "nchw":
        out = torch.empty(B, 4, 3, R, R, dtype=torch.float32, device=dev)
        code = _lib.OUT_NCHW_F32
    elif layout == "patch":
        out = torch.empty(B * 4, g * g, Kp, dtype=torch.bfloat16, device=dev)
        code = _lib.OUT_PATCH_BF16
    else:
        raise ValueError(layout)
    if B == 0:
        return out
    max_side = max(max(Hs), max(Ws))
    need = C.c_size_t()
    _lib.check(lib.b2c_preprocess_workspace_bytes(B, max_side, R, C.byref(need)), "b2c_preprocess_workspace_bytes")
    ws = getattr(cache_owner, "_pre_ws", None) if cache_owner is not None else None
    if ws is None or ws.numel() < need.value + 256:
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
        if cache_owner is not None:
            cache_owner._pre_ws = ws
    wp = (ws.data_ptr() + 255) // 256 * 256
    a_ptrs = (C.c_void_p * B)(*ptrs)
    a_h = (C.c_int * B)(*Hs)
    a_w = (C.c_int * B)(*Ws)
    a_p = (C.c_int * B)(*pitches)
    a_mean = (C.c_float * 3)(*mean)
    a_std = (C.c_float * 3)(*std)
    with torch.cuda.device(dev):
        _lib.check(lib.b2c_preprocess_4crop(a_ptrs, a_h, a_w, a_p, B, R, patch, a_mean, a_std, code, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(wp), ws.numel() - (wp - ws.data_ptr()),
                                            C.c_void_p(_lib.current_stream_ptr())), "b2c_preprocess_4crop")
    del keep
    return out
