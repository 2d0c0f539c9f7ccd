"""Similarity-search variants on the device.

Mirrors, by name and argument meaning:
  * tools/find_similar_imgs.py — ``create_context_embedding`` (:19-62, mean of the context directory's embeddings),
    ``compute_distance`` (:88-94), ``topN`` (:67-85) and ``find_similar_imgs`` (:96-137);
  * _3_label_images.py:128-177 — ``diversity_ordered_image_files``.

What changes underneath: the reference computes one distance per ``torch.load``-ed sample in a Python loop and keeps
the top N with an O(N·top_n) list scan; here all stored embeddings of a directory are scored against the context
vector by one HBM-streaming kernel (``b2c_context_scores``) and the N best are selected exactly on the device
(``b2c_topk_smallest``).  The greedy diversity ordering keeps, for every image, its maximum similarity to the set
selected so far and updates it with one streaming pass per step (``b2c_diversity_order``), with the per-step random
samples drawn on the host from Python's ``random`` exactly as the reference draws them.
"""
from __future__ import annotations

import ctypes as C
import os
import random
from pathlib import Path

import numpy as np
import torch

from . import _lib

_MEASURES = {"cosine": _lib.MEASURE_COSINE_DIST, "l2": _lib.MEASURE_L2, "cosine_sim": _lib.MEASURE_COSINE_SIM}


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.B2C_F32
    if t.dtype == torch.float16:
        return _lib.B2C_F16
    raise TypeError(f"embeddings must be float32 or float16, got {t.dtype}")


def _as_rows(emb: torch.Tensor):
    """Accept a [N,E] tensor whose rows are contiguous (row stride >= E, e.g. one crop column of a packed [N,C,E]
    block) without copying it."""
    if emb.dim() != 2:
        raise ValueError(f"expected [N,E], got {tuple(emb.shape)}")
    if not emb.is_cuda:
        raise _lib.B2CError("similarity search runs on a CUDA device (sm_100a); there is no CPU fallback")
    if emb.shape[0] > 1 and emb.stride(1) != 1:
        emb = emb.contiguous()
    stride = emb.stride(0) if emb.shape[0] > 1 else emb.shape[1]
    if stride < emb.shape[1]:
        emb = emb.contiguous()
        stride = emb.shape[1]
    return emb, stride


def context_scores(emb: torch.Tensor, context: torch.Tensor, similarity_measure: str = "l2", skip=None,
                   out: torch.Tensor | None = None, combine_max: bool = False) -> torch.Tensor:
    """``compute_distance(context, emb[i], similarity_measure)`` for every row i (tools/find_similar_imgs.py:88-94).
    emb: device [N,E] f32/f16; context: [E]; skip: optional bool/u8 [N] (rows that get +inf).  Returns f32 [N]."""
    if similarity_measure not in _MEASURES:
        raise NotImplementedError(f"Similarity measure {similarity_measure} not implemented!")
    emb, stride = _as_rows(emb)
    n, E = emb.shape
    ctx = context.to(emb.device, torch.float32).contiguous().view(-1)
    if ctx.numel() != E:
        raise ValueError(f"context has {ctx.numel()} elements, embeddings have {E}")
    if out is None:
        if combine_max:
            raise ValueError("combine_max needs an existing `out`")
        out = torch.empty(n, dtype=torch.float32, device=emb.device)
    sk = None
    if skip is not None:
        sk = torch.as_tensor(skip).to(emb.device).to(torch.uint8).contiguous()
        if sk.numel() != n:
            raise ValueError("skip mask length mismatch")
    if n:
        with torch.cuda.device(emb.device):
            _lib.check(_lib.load().b2c_context_scores(
                C.c_void_p(emb.data_ptr()), _dtype_code(emb), n, E, stride, C.c_void_p(ctx.data_ptr()), None,
                _MEASURES[similarity_measure], _lib.COMBINE_MAX if combine_max else _lib.COMBINE_STORE,
                C.c_void_p(sk.data_ptr()) if sk is not None else None, C.c_void_p(out.data_ptr()),
                C.c_void_p(_lib.current_stream_ptr())), "b2c_context_scores")
    return out


def topk_smallest(scores: torch.Tensor, k: int):
    """The k smallest entries of a device f32 vector, ascending by (value, index).  Returns (idx int32[k], val f32[k])
    on the device.  Exact; k <= min(n, 4096)."""
    scores = scores.contiguous()
    n = scores.numel()
    k = min(int(k), n)
    idx = torch.empty(k, dtype=torch.int32, device=scores.device)
    val = torch.empty(k, dtype=torch.float32, device=scores.device)
    if k == 0:
        return idx, val
    lib = _lib.load()
    need = C.c_size_t()
    _lib.check(lib.b2c_topk_workspace_bytes(k, C.byref(need)), "b2c_topk_workspace_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(lib.b2c_topk_smallest(C.c_void_p(scores.data_ptr()), n, k, C.c_void_p(idx.data_ptr()),
                                         C.c_void_p(val.data_ptr()), C.c_void_p(ws.data_ptr()), need.value,
                                         C.c_void_p(_lib.current_stream_ptr())), "b2c_topk_smallest")
    return idx, val


def nearest(emb: torch.Tensor, context: torch.Tensor, top_n: int = 30, similarity_measure: str = "l2", skip=None):
    """Indices and distances of the ``top_n`` stored embeddings closest to ``context`` (ascending), skipped rows and
    non-finite distances left out — the result set of tools/find_similar_imgs.py:96-137."""
    d = context_scores(emb, context, similarity_measure, skip)
    idx, val = topk_smallest(d, top_n)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    keep = np.isfinite(val)
    return idx[keep].astype(np.int64), val[keep]


def diversity_order(emb: torch.Tensor, samples, first_row: int = 0) -> np.ndarray:
    """Greedy diversity ordering of _3_label_images.py:135-177 on device rows.  ``samples``: int array [steps, S] of
    row indices (the positions ``random.sample`` drew at each step).  Returns int64 [steps+1] selected rows,
    ``first_row`` first."""
    emb, stride = _as_rows(emb)
    n, E = emb.shape
    smp = torch.as_tensor(np.asarray(samples, dtype=np.int32)).reshape(len(samples), -1)
    steps, S = (int(smp.shape[0]), int(smp.shape[1])) if smp.numel() else (0, 1)
    if steps and (int(smp.min()) < 0 or int(smp.max()) >= n):
        raise ValueError("sample index out of range")
    smp_d = smp.to(emb.device).contiguous() if steps else torch.zeros(1, dtype=torch.int32, device=emb.device)
    maxsim = torch.empty(n, dtype=torch.float32, device=emb.device)
    order = torch.empty(steps + 1, dtype=torch.int32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.load().b2c_diversity_order(
            C.c_void_p(emb.data_ptr()), _dtype_code(emb), n, E, stride, int(first_row), C.c_void_p(smp_d.data_ptr()),
            steps, S, C.c_void_p(maxsim.data_ptr()), C.c_void_p(order.data_ptr()),
            C.c_void_p(_lib.current_stream_ptr())), "b2c_diversity_order")
    return order.cpu().numpy().astype(np.int64)


# ----------------------------------------------------------------------------------------- reference-shaped entry points
def get_filepaths(root_dir, extension=(".pt",)):
    """tools/find_similar_imgs.py:11-17."""
    out = []
    for root, _dirs, files in os.walk(root_dir):
        for file in files:
            if file.endswith(tuple(extension)):
                out.append(os.path.join(root, file))
    return out


def _load_dir_embeddings(args, directory, want_jpg: bool, exclude_names=()):
    """[(pt_path, f32[n_models*E])] for a directory tree, with the reference's skip rules."""
    rows, paths, skips = [], [], 0
    for embedding_path in get_filepaths(directory):
        if want_jpg:
            img_path = embedding_path.replace(".pt", ".jpg")
            if not os.path.exists(img_path) or (Path(img_path).name in exclude_names):
                continue
        try:
            full = torch.load(embedding_path, map_location="cpu")
            if args.clip_models_to_use[0] == "all":
                args.clip_models_to_use = list(full.keys())
                print(f"\n----> Using all found clip models: {args.clip_models_to_use}")
            rows.append(torch.cat([full[m][args.crop_name_to_use].flatten() for m in args.clip_models_to_use], dim=0).float())
            paths.append(embedding_path)
        except Exception as e:  # noqa: BLE001  (tools/find_similar_imgs.py:52-55,127-130: skip the sample)
            print(e)
            skips += 1
    return paths, rows, skips


def create_context_embedding(args, context_dir):
    """tools/find_similar_imgs.py:19-62: mean embedding of the context directory + the context file names."""
    paths, rows, skips = _load_dir_embeddings(args, context_dir, want_jpg=False)
    print(f"Loaded {len(rows)} samples from {context_dir}")
    if skips > 0:
        print(f"(skipped {skips} samples due to loading errors)..")
    feats = torch.stack(rows, dim=0).float()
    return torch.mean(feats, dim=0), [Path(p).name for p in paths]


class topN:
    """Result holder with the reference's attribute names (tools/find_similar_imgs.py:67-85)."""

    def __init__(self, top_n):
        self.top_n = top_n
        self.best_img_paths = []
        self.best_distances = []


def find_similar_imgs(args, context_clip_embedding, context_pathnames, device="cuda"):
    """tools/find_similar_imgs.py:96-137 with the distance + top-N on the device.  NOTE the reference compares
    ``Path(img_path).name`` (a .jpg name) with the context's .pt names (:109), which never match; that behaviour is
    kept.  Returns a ``topN`` whose lists are sorted by ascending distance."""
    print(f"\nSearching {args.search_dir} for similar imgs. Saving results to {args.output_dir}..")
    paths, rows, skips = _load_dir_embeddings(args, args.search_dir, want_jpg=True, exclude_names=context_pathnames)
    top = topN(args.top_n)
    if rows:
        emb = torch.stack(rows).to(device)
        idx, val = nearest(emb, context_clip_embedding, args.top_n, args.similarity_measure)
        top.best_img_paths = [paths[i].replace(".pt", ".jpg") for i in idx.tolist()]
        top.best_distances = [float(v) for v in val]
    print(f"Searched through {len(rows)} samples from {args.search_dir}")
    if skips > 0:
        print(f"(skipped {skips} samples due to loading errors)..")
    return top


def find_similar_in_store(store, context_indices, top_n=30, similarity_measure="l2", crop_name_to_use="square_padded_crop",
                          device="cuda"):
    """Packed-store form: the context is the mean of rows ``context_indices``; those rows are excluded from the
    search.  Returns (paths, distances)."""
    emb = store.crop(crop_name_to_use, torch.float32, device)
    ctx = emb[torch.as_tensor(list(context_indices), device=emb.device)].mean(dim=0)
    skip = np.zeros(len(store), np.uint8)
    skip[list(context_indices)] = 1
    skip |= (~store.has_all([crop_name_to_use])).astype(np.uint8)
    idx, val = nearest(emb, ctx, top_n, similarity_measure, skip)
    return [store.paths[i] for i in idx.tolist()], val


def diversity_ordered_image_files(image_files, root_directory, total_n_ordered_imgs=500, sample_size=100, embeddings=None,
                                  device="cuda"):
    """_3_label_images.py:135-177.  ``embeddings``: optional [N,E] tensor aligned with ``image_files`` (e.g. from a
    packed store); otherwise each image's ``square_padded_crop`` is read from its ``.pt`` — from the first model's
    dict, or from the top level as the reference's older layout had it (:141)."""
    if embeddings is None:
        rows = []
        for f in image_files:
            d = torch.load(os.path.join(root_directory, os.path.basename(f).replace(".jpg", ".pt")), map_location="cpu")
            if "square_padded_crop" not in d:
                d = d[list(d.keys())[0]]
            rows.append(d["square_padded_crop"].squeeze().float())
        embeddings = torch.stack(rows)
    steps = min(total_n_ordered_imgs, len(image_files) - 1)
    print("Creating the most CLIP-diverse ordering of the first ", total_n_ordered_imgs, " images...")
    index_of = {f: i for i, f in enumerate(image_files)}
    samples = [[index_of[f] for f in random.sample(image_files, sample_size)] for _ in range(steps)]
    order = diversity_order(embeddings.to(device), samples) if steps > 0 else np.zeros(1, np.int64)
    img_files = [image_files[i] for i in order.tolist()]
    chosen = set(img_files)
    return img_files + [f for f in image_files if f not in chosen]
