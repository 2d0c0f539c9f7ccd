"""Drop-in mirror of the reference's duplicate search (_2_remove_duplicates.py) on the B200 path.

  * ``get_paths_and_embeddings(args, crop_to_use, shuffle=False)``  — _2_remove_duplicates.py:8-49 (host I/O, unchanged semantics)
  * ``find_near_duplicates(args, sim_type='cosine', crop_to_use='square_padded_crop')`` — :52-99
  * ``fix_duplicate(duplicate_index, img_paths, outdir, sim_value, mode)`` — :102-125
  * ``duplicate_pairs(...)`` / ``duplicate_pairs_distributed(...)`` — the device core (:63-80) as one call.

The N x N similarity matrix is never materialised (b2c_dedup_pairs thresholds and emits pairs in the GEMM
epilogue), so ``chunk_size`` is no longer a memory cap; with ``chunk_size >= N`` results equal the
reference's: strict ``>``, ``i < j``, pairs in row-major order, comparison on fp16-rounded similarities.
"""
from __future__ import annotations

import ctypes as C
import os
import random
import shutil

import numpy as np
import torch

from . import _lib

BAND_ROWS = 2048  # rows per scheduling band of b2c_dedup_pairs (16 row blocks of 128)


# ----------------------------------------------------------------------------------------- host I/O
def get_paths_and_embeddings(args, crop_to_use, shuffle=False):
    """Per-subdirectory generator of (paths, embeddings) chunks — _2_remove_duplicates.py:8-49."""
    for subdir, dirs, files in os.walk(args.root_dir):
        print(f"\nParsing {subdir}, subdirs: {dirs}, n_files: {len(files)}..")
        paths, embeddings = [], []
        if shuffle:
            random.shuffle(files)
        unique_filenames = {}
        for file in files:
            filename, ext = os.path.splitext(file)
            unique_filenames.setdefault(filename, []).append(ext)
        print(f"Loading embeddings for {len(unique_filenames)} unique filenames..")
        for filename, exts in unique_filenames.items():
            if ".jpg" in exts and ".pt" in exts:
                try:
                    path = os.path.join(subdir, filename + ".jpg")
                    embedding_dict = torch.load(os.path.join(subdir, filename + ".pt"))
                    if args.clip_model_to_use is None:
                        args.clip_model_to_use = list(embedding_dict.keys())[0]
                        print(f"\n ----> args.clip_model_to_use was not specified, defaulting to first found one: "
                              f"{args.clip_model_to_use} \n")
                    embedding_dict = embedding_dict[args.clip_model_to_use]
                    embedding = embedding_dict[crop_to_use].squeeze().to(torch.float16)
                    paths.append(path)
                    embeddings.append(embedding)
                    if len(paths) == args.chunk_size:
                        yield paths, embeddings
                        paths, embeddings = [], []
                except Exception:  # noqa: BLE001  (the reference silently skips unreadable samples, :45-46)
                    continue
        if len(paths) > 0:
            yield paths, embeddings


# ----------------------------------------------------------------------------------------- device core
def owned_bands(n_total: int, rank: int = 0, world_size: int = 1, band_rows: int = BAND_ROWS):
    """Row ranges [(begin, end), ...] of the upper-triangle bands this rank computes.  Band b has
    (n_total - b*band_rows) columns of work — linearly decreasing — so bands are dealt in a snake
    (0..P-1, P-1..0, ...): every pair of rounds hands each rank the same amount of work."""
    n_bands = (n_total + band_rows - 1) // band_rows
    out = []
    for b in range(n_bands):
        rnd, pos = divmod(b, world_size)
        owner = pos if rnd % 2 == 0 else world_size - 1 - pos
        if owner == rank:
            out.append((b * band_rows, min(n_total, (b + 1) * band_rows)))
    return out


def sort_pairs(pairs: np.ndarray, sims: np.ndarray):
    """Row-major (i asc, then j asc): the order torch.where returns at _2_remove_duplicates.py:74."""
    if len(pairs) == 0:
        return pairs.reshape(0, 2).astype(np.int64), sims.astype(np.float32)
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    return pairs[order].astype(np.int64), sims[order].astype(np.float32)


def normalize_rows_f16(embeddings: torch.Tensor) -> torch.Tensor:
    """[n,E] f32/f16 (device) -> unit-norm f16 [n, E_pad], E_pad = round_up(E, 64) — _2_remove_duplicates.py:67."""
    lib = _lib.load()
    if embeddings.dtype not in (torch.float32, torch.float16):
        embeddings = embeddings.float()
    embeddings = embeddings.contiguous()
    n, E = embeddings.shape
    E_pad = (E + 63) // 64 * 64
    out = torch.empty(n, E_pad, dtype=torch.float16, device=embeddings.device)
    if n:
        with torch.cuda.device(embeddings.device):
            _lib.check(lib.b2c_normalize_rows_f16(C.c_void_p(embeddings.data_ptr()),
                                                  _lib.B2C_F32 if embeddings.dtype == torch.float32 else _lib.B2C_F16,
                                                  n, E, C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr())),
                       "b2c_normalize_rows_f16")
    return out


def _pairs_for_ranges(emb_n: torch.Tensor, ranges, threshold: float, compare: str, capacity: int):
    """Run b2c_dedup_pairs over row ranges of normalised f16 [n_total, E_pad]; returns (pairs[K,2], sims[K]) numpy, unsorted."""
    lib = _lib.load()
    n_total, E_pad = emb_n.shape
    mode = _lib.CMP_REF_FP16 if compare == "ref_fp16" else _lib.CMP_FP32
    dev = emb_n.device
    merged = []  # coalesce adjacent ranges: one call covers them (the library iterates the bands itself)
    for (r0, r1) in ranges:
        if merged and merged[-1][1] == r0:
            merged[-1] = (merged[-1][0], r1)
        else:
            merged.append((r0, r1))
    ranges = merged
    while True:
        buf = torch.empty(max(capacity, 1), 3, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(_lib.current_stream_ptr())
            for (r0, r1) in ranges:
                _lib.check(lib.b2c_dedup_pairs(C.c_void_p(emb_n.data_ptr()), n_total, E_pad, r0, r1, C.c_float(threshold), mode,
                                               C.c_void_p(buf.data_ptr()), capacity, C.c_void_p(cnt.data_ptr()), st),
                           "b2c_dedup_pairs")
        k = int(cnt.item())
        if k <= capacity:
            break
        capacity = int(k * 1.25) + 1024  # overflow: the count is exact, re-run with room for every pair
    raw = buf[:k].cpu().numpy()
    pairs = raw[:, :2].astype(np.int64)
    sims = raw[:, 2].copy().view(np.float32)
    return pairs, sims


def duplicate_pairs(embeddings: torch.Tensor, threshold: float, compare: str = "ref_fp16", device=None,
                    capacity: int | None = None):
    """All pairs (i < j) whose cosine similarity exceeds ``threshold`` — the core of find_near_duplicates
    (_2_remove_duplicates.py:63-80).  Returns (pairs int64 [K,2] in row-major order, sims float32 [K])."""
    dev = torch.device(device) if device is not None else (embeddings.device if embeddings.is_cuda else torch.device("cuda"))
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.B2CError("duplicate_pairs needs a CUDA device (sm_100a); there is no CPU fallback")
    emb = embeddings.to(dev)
    n = emb.shape[0]
    if n < 2:
        return np.zeros((0, 2), np.int64), np.zeros((0,), np.float32)
    emb_n = normalize_rows_f16(emb)
    cap = capacity if capacity is not None else max(1 << 16, 4 * n)
    pairs, sims = _pairs_for_ranges(emb_n, owned_bands(n), float(threshold), compare, cap)
    return sort_pairs(pairs, sims)


def duplicate_pairs_distributed(local_embeddings: torch.Tensor, threshold: float, compare: str = "ref_fp16",
                                group=None, capacity: int | None = None, _pair_fn=None):
    """Multi-GPU form: every rank passes its shard [n_local, E] (equal n_local on all ranks; pad the last shard
    with zero rows, which never match).  The shards are normalised locally, all-gathered ONCE over NCCL, and
    each rank searches the bands ``owned_bands`` deals it.  Every rank returns the full sorted result."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pair_fn = _pair_fn or _pairs_for_ranges
    if _pair_fn is None:
        local_n = normalize_rows_f16(local_embeddings)
    else:  # host-logic tests on CPU inject the pair finder and skip the CUDA normalise
        local_n = torch.nn.functional.normalize(local_embeddings.float(), dim=1).to(torch.float16)
    n_local, E_pad = local_n.shape
    gathered = torch.empty(world * n_local, E_pad, dtype=local_n.dtype, device=local_n.device)
    dist.all_gather_into_tensor(gathered, local_n.contiguous(), group=group)
    n_total = gathered.shape[0]
    cap = capacity if capacity is not None else max(1 << 16, 4 * n_total // world)
    pairs, sims = pair_fn(gathered, owned_bands(n_total, rank, world), float(threshold), compare, cap)
    # variable-length exchange of the (small) pair lists
    payload = [None] * world
    dist.all_gather_object(payload, (pairs, sims), group=group)
    all_pairs = np.concatenate([p for p, _ in payload]) if payload else pairs
    all_sims = np.concatenate([s for _, s in payload]) if payload else sims
    return sort_pairs(all_pairs, all_sims)


# ----------------------------------------------------------------------------------------- reference entry
def find_near_duplicates(args, sim_type="cosine", crop_to_use="square_padded_crop"):
    """_2_remove_duplicates.py:52-99 with the similarity/threshold/where core on the GPU kernel."""
    if sim_type != "cosine":
        raise NotImplementedError("only sim_type='cosine' is implemented (the reference CLI never selects 'euclidean')")
    results = []
    for paths, embeddings in get_paths_and_embeddings(args, crop_to_use):
        if len(paths) == 0 or len(embeddings) == 0:
            continue
        embeddings = torch.stack(embeddings)
        print(f"Got first batch of embeddings of shape: {embeddings.shape}, computing similarity matrix..")
        pairs, sims = duplicate_pairs(embeddings, args.threshold)
        near_duplicates = [(paths[i], paths[j]) for i, j in pairs.tolist()]
        # the reference reads values back from its fp16 matrix (:80)
        near_duplicate_values = [float(np.float16(s)) for s in sims]
        output_dir = os.path.join(os.path.dirname(args.root_dir), f"near_duplicates_{sim_type}_{args.threshold}")
        os.makedirs(output_dir, exist_ok=True)
        i = 0
        print(f"Found {len(near_duplicates)} duplicates!")
        if len(near_duplicates) > 0 and not args.test:
            verb = "copying" if args.mode == "copy" else "moving"
            print(f"{verb} {len(near_duplicates)} near duplicates to {output_dir}...")
            for i, (img_paths, sim_value) in enumerate(zip(near_duplicates, near_duplicate_values)):
                fix_duplicate(i, img_paths, output_dir, sim_value, args.mode)
            if args.mode == "move":
                print(f"Moved {i} duplicates to {output_dir}")
            elif args.mode == "copy":
                print(f"Copied {i} duplicates (not removed from data yet!) to {output_dir}")
        results.append((near_duplicates, near_duplicate_values))
    return results


def find_near_duplicates_in_store(store, threshold=0.96, crop_to_use="square_padded_crop", per_directory=True,
                                  compare="ref_fp16"):
    """The same search fed from a packed store (store.PackedStore) instead of one torch.load per
    image (_2_remove_duplicates.py:25-46).  ``per_directory=True`` keeps the reference's scope — duplicates are only
    looked for among the images of one directory (:10) — and compares them in sorted-path order; ``False`` searches the
    whole store at once.  Embeddings go through fp16 like the reference's loader (:38).  Returns
    [(near_duplicates [(path_i, path_j)], near_duplicate_values [float])] per group."""
    emb = store.crop(crop_to_use, torch.float16)
    usable = store.has_all([crop_to_use])
    groups = {}
    for i, p in enumerate(store.paths):
        if usable[i]:
            groups.setdefault(os.path.dirname(p) if per_directory else "", []).append(i)
    results = []
    for key in sorted(groups):
        idx = sorted(groups[key], key=lambda i: store.paths[i])
        if len(idx) < 2:
            results.append(([], []))
            continue
        pairs, sims = duplicate_pairs(emb[idx], threshold, compare)
        results.append(([(store.paths[idx[i]], store.paths[idx[j]]) for i, j in pairs.tolist()],
                        [float(np.float16(s)) for s in sims]))
    return results


def fix_duplicate(duplicate_index, img_paths, outdir, sim_value, mode):
    """_2_remove_duplicates.py:102-125: copy both images' companion files, or move only the target's."""
    dirname = os.path.dirname(img_paths[0])
    basename1 = os.path.splitext(os.path.basename(img_paths[0]))[0]
    basename2 = os.path.splitext(os.path.basename(img_paths[1]))[0]
    files1 = [os.path.join(dirname, f) for f in os.listdir(os.path.dirname(img_paths[0])) if basename1 in f]
    files2 = [os.path.join(dirname, f) for f in os.listdir(os.path.dirname(img_paths[1])) if basename2 in f]
    for f in files1:
        if mode == "copy":
            shutil.copy(f, os.path.join(outdir, f"{sim_value:.3f}_{duplicate_index:08d}_source_{os.path.basename(f)}"))
    for f in files2:
        if mode == "copy":
            shutil.copy(f, os.path.join(outdir, f"{sim_value:.3f}_{duplicate_index:08d}_target_{os.path.basename(f)}"))
        if mode == "move":
            os.rename(f, os.path.join(outdir, f"{sim_value:.3f}_{duplicate_index:08d}_target_{os.path.basename(f)}"))
    return


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--root_dir", type=str, help="Root directory of the dataset")
    parser.add_argument("--threshold", type=float, default=0.96, help="Cosine-similarity threshold for near-duplicate detection")
    parser.add_argument("--mode", type=str, default="copy", help="copy / move, Use copy to test the script, move after")
    parser.add_argument("--clip_model_to_use", type=str, default=None, help="Which CLIP model to use, if None, use the first one found")
    parser.add_argument("--chunk_size", type=int, default=10_000_000,
                        help="Max embeddings compared at once per directory (the reference's 10000 memory cap is lifted)")
    parser.add_argument("--test", action="store_true", help="Test the script without doing anything")
    find_near_duplicates(parser.parse_args(argv))


if __name__ == "__main__":
    main()
