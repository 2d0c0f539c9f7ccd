"""ctypes binding of libb2c.so (include/b2c.h).

The library is loaded lazily (DataLoader workers re-import this package under ``spawn`` and must not
touch CUDA — reference `_1_embed_with_CLIP.py`:202) and loudly: if ``libb2c.so`` is missing or a call
fails there is no fallback, a ``B2CError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# B2C_LIB: development override (A/B builds of the same sources with different compile-time constants)
LIB_PATH = os.environ.get("B2C_LIB") or os.path.join(PKG_DIR, "libb2c.so")

B2C_F32, B2C_F16, B2C_BF16, B2C_U8 = 0, 1, 2, 3
OUT_NCHW_F32, OUT_PATCH_BF16 = 0, 1
ACT_QUICK_GELU, ACT_GELU = 0, 1
EPI_BIAS_BF16, EPI_BIAS_QGELU_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32 = 0, 1, 2, 3
CMP_FP32, CMP_REF_FP16, CMP_EUCLID = 0, 1, 2
MEASURE_COSINE_DIST, MEASURE_L2, MEASURE_COSINE_SIM = 0, 1, 2
COMBINE_STORE, COMBINE_MAX = 0, 1
TOPK_MAX = 4096
MLP_MAX_LAYERS = 8
PROF_KINDS = 11
IMG_STATS = 22


class B2CError(RuntimeError):
    pass


class Crop(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("cw", "ch", "dx", "dy", "out_w", "out_h", "off_x", "off_y")]


class VitCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("image", "patch", "width", "layers", "heads", "mlp", "embed", "act")]


class Pair(C.Structure):
    _fields_ = [("i", C.c_int32), ("j", C.c_int32), ("sim", C.c_float)]


class MlpWeights(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("dims", C.c_int32 * (MLP_MAX_LAYERS + 1)),
        ("weight", C.c_void_p * MLP_MAX_LAYERS),
        ("bias", C.c_void_p * MLP_MAX_LAYERS),
        ("leaky_slope", C.c_float),
    ]


class TrainerCfg(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MLP_MAX_LAYERS + 1)), ("max_batch", C.c_int32),
                ("leaky_slope", C.c_float), ("dropout_p", C.c_float), ("seed", C.c_uint64)]


class Adam(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("lr", "beta1", "beta2", "eps", "weight_decay")]


_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
_u64 = C.c_ulonglong

# name -> (restype, argtypes); every symbol include/b2c.h declares
SIGNATURES = {
    "b2c_last_error": (C.c_char_p, []),
    "b2c_version": (_i, []),
    "b2c_launch_count": (_u64, []),
    "b2c_prof_enable": (_i, [_i]),
    "b2c_prof_read": (_i, [C.POINTER(C.c_double), C.POINTER(_u64)]),
    "b2c_prof_kind_name": (C.c_char_p, [_i]),
    "b2c_crop_geometry": (_i, [_i, _i, _i, C.POINTER(Crop)]),
    "b2c_preprocess_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "b2c_preprocess_4crop": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _i, _i,
                                  C.POINTER(_f), C.POINTER(_f), _i, _vp, _vp, _sz, _vp]),
    "b2c_vit_create": (_i, [C.POINTER(VitCfg), C.POINTER(_vp)]),
    "b2c_vit_destroy": (_i, [_vp]),
    "b2c_vit_set_weight": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i64), _i]),
    "b2c_vit_ready": (_i, [_vp]),
    "b2c_vit_set_lanes": (_i, [_vp, _i]),
    "b2c_vit_set_fused_ln": (_i, [_vp, _i]),
    "b2c_vit_set_graph": (_i, [_vp, _i]),
    "b2c_vit_set_cls_only_last_block": (_i, [_vp, _i]),
    "b2c_vit_workspace_bytes": (_i, [_vp, _i, C.POINTER(_sz)]),
    "b2c_vit_forward_pixels": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "b2c_vit_forward_patches": (_i, [_vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "b2c_jpeg_parse": (_i, [_vp, _sz, _vp]),
    "b2c_jpeg_decode_coefs": (_i, [_vp, _sz, _vp, _vp, _sz]),
    "b2c_jpeg_decode_packed": (_i, [_vp, _sz, _vp, _vp, _sz, _vp]),
    "b2c_jpeg_reconstruct_packed": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "b2c_jpeg_workspace_bytes": (_i, [_vp, _i, C.POINTER(_sz)]),
    "b2c_jpeg_reconstruct": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "b2c_jpeg_huff_prepare": (_i, [_vp, _sz, _vp, _vp]),
    "b2c_jpeg_huff_workspace_bytes": (_i, [_vp, _i, C.POINTER(_sz)]),
    "b2c_jpeg_huff_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "b2c_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "b2c_layernorm_bf16": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _f, _vp]),
    "b2c_attention_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "b2c_normalize_rows_f16": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "b2c_dedup_pairs": (_i, [_vp, _i64, _i, _i64, _i64, _f, _i, _vp, _u64, _vp, _vp]),
    "b2c_dedup_pairs_block": (_i, [_vp, _i64, _i, _i64, _i64, _i64, _i64, _f, _i, _vp, _u64, _vp, _vp]),
    "b2c_mlp_score": (_i, [_vp, _i64, C.POINTER(MlpWeights), _vp, _vp]),
    "b2c_image_stats_target_size": (_i, [_i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "b2c_image_stats_workspace_bytes": (_i, [C.POINTER(_i), C.POINTER(_i), _i, C.POINTER(_sz)]),
    "b2c_image_stats": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp, _sz, _vp]),
    "b2c_trainer_create": (_i, [C.POINTER(TrainerCfg), C.POINTER(_vp)]),
    "b2c_trainer_destroy": (_i, [_vp]),
    "b2c_trainer_set_layer": (_i, [_vp, _i, _vp, _vp, _vp]),
    "b2c_trainer_get_layer": (_i, [_vp, _i, _vp, _vp, _vp]),
    "b2c_trainer_weights": (_i, [_vp, C.POINTER(MlpWeights)]),
    "b2c_trainer_steps": (_u64, [_vp]),
    "b2c_trainer_epoch": (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _i, C.POINTER(Adam), _vp, _vp]),
    "b2c_context_scores": (_i, [_vp, _i, _i64, _i, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "b2c_topk_workspace_bytes": (_i, [_i, C.POINTER(_sz)]),
    "b2c_topk_smallest": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _sz, _vp]),
    "b2c_diversity_order": (_i, [_vp, _i, _i64, _i, _i64, C.c_int32, _vp, _i, _i, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libb2c.so (once) and attach the prototypes. Raises B2CError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B2CError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b2c_last_error().decode("utf-8", "replace")
        raise B2CError(f"{what or 'libb2c call'} failed (code {rc}): {msg}")


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise B2CError on a non-zero status."""
    check(getattr(load(), name)(*args), name)


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().b2c_launch_count())


def prof_enable(on: bool) -> None:
    """Start (and clear) or stop the library's per-stage CUDA-event timer (include/b2c.h: b2c_prof_enable)."""
    call("b2c_prof_enable", 1 if on else 0)


def prof_read() -> dict:
    """{stage name: (milliseconds, stages)} accumulated since prof_enable(True); synchronises the recorded events."""
    lib = load()
    ms = (C.c_double * PROF_KINDS)()
    cnt = (_u64 * PROF_KINDS)()
    check(lib.b2c_prof_read(ms, cnt), "b2c_prof_read")
    return {lib.b2c_prof_kind_name(k).decode(): (ms[k], int(cnt[k])) for k in range(PROF_KINDS) if cnt[k]}
