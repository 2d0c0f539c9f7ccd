"""Drop-in mirror of the reference's duplicate search (_2_remove_duplicates.py) on the B200 path.

  * ``get_paths_and_embeddings(args, crop_to_use, shuffle=False)``  — _2_remove_duplicates.py:8-49 (host I/O)
  * ``find_near_duplicates(args, sim_type='cosine', crop_to_use='square_padded_crop')`` — :52-99
  * ``fix_duplicate(duplicate_index, img_paths, outdir, sim_value, mode)`` — :102-125
  * ``duplicate_pairs(...)`` / ``duplicate_pairs_distributed(...)`` — the device core (:63-80) as one call.

The N x N similarity matrix is never materialised (b2c_dedup_pairs thresholds and emits pairs in the GEMM
epilogue), so ``chunk_size`` is no longer a memory cap; with ``chunk_size >= N`` results equal the
reference's: strict ``>``, ``i < j``, pairs in row-major order, comparison on fp16-rounded similarities.
The host functions keep the reference's behaviour (which files pair up, what is skipped, how the copies are
named) but are written for this package; INTEGRATION.md shows that the reference's own script can equally
keep its host half and call ``duplicate_pairs`` for lines 63-80.
"""
from __future__ import annotations

import ctypes as C
import os
import random
import shutil

import numpy as np
import torch

from . import _lib

STREAM_MIN_ROWS = 200_000  # find_near_duplicates_in_store: groups from this size on are streamed to the device
BAND_ROWS = 2048  # rows per scheduling band of b2c_dedup_pairs (16 row blocks of 128)
PAIR_DTYPE = np.dtype([("i", np.int32), ("j", np.int32), ("sim", np.float32)])  # b2c_pair


# ----------------------------------------------------------------------------------------- host I/O
def _stems_with_image_and_embedding(files, shuffle):
    """File stems of one directory that have both ``X.jpg`` and ``X.pt`` (:20-27), in first-seen order."""
    names = list(files)
    if shuffle:
        random.shuffle(names)
    exts_of = {}
    for name in names:
        stem, ext = os.path.splitext(name)
        exts_of.setdefault(stem, set()).add(ext)
    return [stem for stem, exts in exts_of.items() if ".jpg" in exts and ".pt" in exts]


def _read_embedding(pt_path, args, crop_to_use):
    """One crop embedding as fp16 [E] (:30-38); the model key defaults to the file's first one and then sticks (:32-35)."""
    per_model = torch.load(pt_path)
    if args.clip_model_to_use is None:
        args.clip_model_to_use = next(iter(per_model))
        print(f"\n ----> args.clip_model_to_use was not specified, using the first one found: {args.clip_model_to_use}\n")
    return per_model[args.clip_model_to_use][crop_to_use].squeeze().to(torch.float16)


def get_paths_and_embeddings(args, crop_to_use, shuffle=False):
    """Per directory of ``args.root_dir``: chunks ``(paths, embeddings)`` of at most ``args.chunk_size`` images that
    have an ``.pt`` file next to the ``.jpg`` — _2_remove_duplicates.py:8-49.  Unreadable samples are skipped (:45-46)."""
    for subdir, dirs, files in os.walk(args.root_dir):
        stems = _stems_with_image_and_embedding(files, shuffle)
        print(f"\n{subdir}: {len(files)} files in {len(dirs)} sub-directories, {len(stems)} images with embeddings")
        paths, embeddings = [], []
        for stem in stems:
            try:
                vec = _read_embedding(os.path.join(subdir, stem + ".pt"), args, crop_to_use)
            except Exception:  # noqa: BLE001
                continue
            paths.append(os.path.join(subdir, stem + ".jpg"))
            embeddings.append(vec)
            if len(paths) == args.chunk_size:
                yield paths, embeddings
                paths, embeddings = [], []
        if paths:
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


def normalize_rows_f16(embeddings: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[n,E] f32/f16 (device) -> unit-norm f16 [n, E_pad], E_pad = round_up(E, 64) — _2_remove_duplicates.py:67.
    ``out``: optional destination (e.g. this rank's slice of the all-gather buffer)."""
    lib = _lib.load()
    if not embeddings.is_cuda:
        raise _lib.B2CError("normalize_rows_f16 needs a CUDA tensor (sm_100a); there is no CPU fallback")
    if embeddings.dtype not in (torch.float32, torch.float16):
        embeddings = embeddings.float()
    embeddings = embeddings.contiguous()
    n, E = embeddings.shape
    E_pad = (E + 63) // 64 * 64
    if out is None:
        out = torch.empty(n, E_pad, dtype=torch.float16, device=embeddings.device)
    assert out.shape == (n, E_pad) and out.dtype == torch.float16 and out.is_contiguous()
    if n:
        with torch.cuda.device(embeddings.device):
            _lib.check(lib.b2c_normalize_rows_f16(C.c_void_p(embeddings.data_ptr()),
                                                  _lib.B2C_F32 if embeddings.dtype == torch.float32 else _lib.B2C_F16,
                                                  n, E, C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr())),
                       "b2c_normalize_rows_f16")
    return out


_MODES = {"fp32": _lib.CMP_FP32, "ref_fp16": _lib.CMP_REF_FP16, "euclidean": _lib.CMP_EUCLID}


def _merge_ranges(ranges):
    merged = []  # adjacent ranges become one call (the library iterates the bands itself)
    for (r0, r1) in ranges:
        if merged and merged[-1][1] == r0:
            merged[-1] = (merged[-1][0], r1)
        else:
            merged.append((r0, r1))
    return merged


def launch_pair_search(emb_n: torch.Tensor, ranges, threshold: float, compare: str, buf: torch.Tensor, cnt: torch.Tensor):
    """Enqueue the pair search over blocks of normalised f16 [n_total, E_pad] on the current stream (no sync): pairs are
    appended to ``buf`` (int32 [capacity, 3] = b2c_pair), ``cnt`` (int64 [1]) counts every hit.  ``ranges``: row ranges
    ``(r0, r1)`` (all columns j > i) or blocks ``(r0, r1, c0, c1)`` (columns c0 <= j < c1 only)."""
    lib = _lib.load()
    n_total, E_pad = emb_n.shape
    plain = _merge_ranges([r for r in ranges if len(r) == 2])
    blocks = [r for r in ranges if len(r) == 4]
    with torch.cuda.device(emb_n.device):
        st = C.c_void_p(_lib.current_stream_ptr())
        for (r0, r1, c0, c1) in [(a, b, 0, n_total) for a, b in plain] + blocks:
            _lib.check(lib.b2c_dedup_pairs_block(C.c_void_p(emb_n.data_ptr()), n_total, E_pad, r0, r1, c0, c1, C.c_float(threshold),
                                                 _MODES[compare], C.c_void_p(buf.data_ptr()), buf.shape[0],
                                                 C.c_void_p(cnt.data_ptr()), st), "b2c_dedup_pairs_block")


def _unpack(raw: np.ndarray):
    """int32 [K,3] rows of b2c_pair -> (pairs int64 [K,2], sims float32 [K])"""
    return raw[:, :2].astype(np.int64), raw[:, 2].copy().view(np.float32)


def _pairs_for_ranges(emb_n: torch.Tensor, ranges, threshold: float, compare: str, capacity: int):
    """Run b2c_dedup_pairs over row ranges of normalised f16 [n_total, E_pad]; returns (pairs[K,2], sims[K]) numpy, unsorted."""
    dev = emb_n.device
    while True:
        buf = torch.empty(max(capacity, 1), 3, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        launch_pair_search(emb_n, ranges, threshold, compare, buf, cnt)
        k = int(cnt.item())
        if k <= capacity:
            break
        capacity = int(k * 1.25) + 1024  # overflow: the count is exact, re-run with room for every pair
    return _unpack(buf[:k].cpu().numpy())


def duplicate_pairs(embeddings: torch.Tensor, threshold: float, compare: str = "ref_fp16", device=None,
                    capacity: int | None = None):
    """All pairs (i < j) whose cosine similarity exceeds ``threshold`` — the core of find_near_duplicates
    (_2_remove_duplicates.py:63-80).  Returns (pairs int64 [K,2] in row-major order, sims float32 [K]).
    compare: 'ref_fp16' (the reference's comparison on fp16-rounded similarities), 'fp32', or 'euclidean' (pairs whose
    distance exceeds the threshold, the reference's other ``sim_type``; ``sims`` then holds distances)."""
    if compare not in _MODES:
        raise ValueError(f"compare must be one of {sorted(_MODES)}")
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


_stream_slots = {}   # pinned staging buffers of _stage_rows, kept between calls (release_stream_buffers())
_stream_free = {}    # per slot: the event after which it may be refilled (kept with the slot: the next call reuses it)
_stream_pool = None


def release_stream_buffers():
    """Give back the pinned host buffers duplicate_pairs_streamed keeps between calls (2 x chunk_rows x E elements)."""
    for evs in _stream_free.values():
        for ev in evs:
            if ev is not None:
                ev.synchronize()
    _stream_free.clear()
    _stream_slots.clear()


def _stage_rows(rows_into, n: int, E: int, src_dtype, dev, chunk_rows: int, consume):
    """Host rows -> device, chunk by chunk through two pinned slots: chunk k = rows [a, b) is gathered by four host threads
    (``rows_into(a, b, out)`` fills the numpy array ``out`` [b-a, E]), copied to the device on a side stream, and handed to
    ``consume(a, b, d)`` — called with the side stream current, ``d`` = the chunk on the device ([b-a, E] of ``src_dtype``).
    The gather of chunk k+1 overlaps the copy (and whatever ``consume`` queues) of chunk k.  Returns the side stream; work
    the caller queues elsewhere afterwards has to wait for it (``wait_stream``)."""
    global _stream_pool
    import concurrent.futures as cf
    key = (chunk_rows, E, src_dtype)
    if key not in _stream_slots:
        _stream_slots[key] = [torch.empty(chunk_rows, E, dtype=src_dtype, pin_memory=True) for _ in range(2)]
        _stream_free[key] = [None, None]
    slots, free = _stream_slots[key], _stream_free[key]
    if _stream_pool is None:
        _stream_pool = cf.ThreadPoolExecutor(4)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(main)  # what the caller allocated / queued so far exists before the side stream touches it
    for k, a in enumerate(range(0, n, chunk_rows)):
        b = min(n, a + chunk_rows)
        m = b - a
        slot = slots[k & 1]
        if free[k & 1] is not None:
            free[k & 1].synchronize()
        view = slot[:m].numpy()
        cuts = [m * t // 4 for t in range(5)]
        list(_stream_pool.map(lambda t: rows_into(a + cuts[t], a + cuts[t + 1], view[cuts[t]:cuts[t + 1]]), range(4)))
        with torch.cuda.stream(side):
            d = slot[:m].to(dev, non_blocking=True)
            free[k & 1] = torch.cuda.Event()
            free[k & 1].record(side)
            consume(a, b, d)
    return side


def duplicate_pairs_streamed(rows_into, n: int, E: int, src_dtype, threshold: float, compare: str = "ref_fp16", device=None,
                             chunk_rows: int = 1 << 16, capacity: int | None = None):
    """``duplicate_pairs`` for embeddings that still sit on the host (a packed store): the rows are brought over in chunks
    of ``chunk_rows`` through two pinned buffers on a side stream, and as soon as chunk k = rows [a, b) is on the device
    (normalised straight into its place) the search of the block rows [0, b) x columns [a, b) is queued — every pair
    (i < j) belongs to the chunk that holds j — so the host gather, the H2D copies and the search overlap and the call
    takes about as long as the search alone.  ``rows_into(a, b, out)`` fills the numpy array ``out`` [b-a, E] with rows
    a..b of the set, in the order the pairs are to be reported in."""
    if compare not in _MODES:
        raise ValueError(f"compare must be one of {sorted(_MODES)}")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.B2CError("duplicate_pairs_streamed needs a CUDA device (sm_100a); there is no CPU fallback")
    if n < 2:
        return np.zeros((0, 2), np.int64), np.zeros((0,), np.float32)
    E_pad = (E + 63) // 64 * 64
    cap = capacity if capacity is not None else max(1 << 16, 4 * n)
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        gathered = torch.empty(n, E_pad, dtype=torch.float16, device=dev)
        buf = torch.empty(max(cap, 1), 3, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)

        def consume(a, b, d):
            if d.dtype not in (torch.float16, torch.float32):
                d = d.float()
            normalize_rows_f16(d.to(torch.float16), out=gathered[a:b])  # fp16 first, like the reference's loader (_2:38)
            ready = torch.cuda.Event()
            ready.record()  # (on the side stream, which is current here)
            main.wait_event(ready)
            with torch.cuda.stream(main):
                launch_pair_search(gathered, [(0, b, a, b)], float(threshold), compare, buf, cnt)

        _stage_rows(rows_into, n, E, src_dtype, dev, chunk_rows, consume)
        kfound = int(cnt.item())
        if kfound > cap:  # the count is exact: search the resident set again with room for every pair
            pairs, sims = _pairs_for_ranges(gathered, owned_bands(n), float(threshold), compare, int(kfound * 1.25) + 1024)
        else:
            pairs, sims = _unpack(buf[:kfound].cpu().numpy())
    return sort_pairs(pairs, sims)


def bands_right_of_diagonal(n_local: int, world_size: int, band_rows: int = BAND_ROWS):
    """Everything to the right of the shards' diagonal blocks as (r0, r1, c0, c1) bands of ``band_rows`` rows (cut at shard
    boundaries), columns from the end of the band's shard to n_total — largest first (ties by row), the order the ranks
    draw them in."""
    n_total = n_local * world_size
    bands = []
    for s in range(world_size - 1):  # the last shard has nothing to its right
        for r0 in range(s * n_local, (s + 1) * n_local, band_rows):
            bands.append((r0, min(r0 + band_rows, (s + 1) * n_local), (s + 1) * n_local, n_total))
    bands.sort(key=lambda t: (-(t[1] - t[0]) * (t[3] - t[2]), t[0]))
    return bands


def owned_blocks(n_local: int, rank: int, world_size: int, band_rows: int = BAND_ROWS):
    """Work of one rank in the multi-GPU search, as (r0, r1, c0, c1) blocks of the global upper triangle:
      first   the block of its OWN shard (rows and columns in [rank*n_local, (rank+1)*n_local)) — needs no peer data,
              so it runs while the all-gather is in flight;
      then    its share of everything to the right of the shards' diagonal blocks (``bands_right_of_diagonal``).  Work per
              band falls from shard to shard, so the bands are dealt largest-first to the least-loaded rank (every rank
              computes the same deal).  This is the STATIC split (single calls, tests); duplicate_pairs_distributed lets
              the ranks draw the same bands from a shared counter instead, because the GPUs of a box run at different
              power-capped clocks.
    Every pair (i < j) belongs to exactly one block of exactly one rank."""
    lo = rank * n_local
    local = [(lo, lo + n_local, lo, lo + n_local)] if n_local > 1 else []
    load = [n_local * (n_local - 1) / 2.0] * world_size  # everybody starts with its own shard's triangle
    rest = []
    for blk in bands_right_of_diagonal(n_local, world_size, band_rows):
        w = (blk[1] - blk[0]) * (blk[3] - blk[2])
        owner = min(range(world_size), key=lambda r: (load[r], r))
        load[owner] += w
        if owner == rank:
            rest.append(blk)
    rest.sort()
    return local, rest


def launch_marker():
    """An event recorded on the current stream after the launches so far; ``.synchronize()`` waits for them."""
    ev = torch.cuda.Event()
    ev.record()
    return ev


_ticket_calls = 0


class BandTickets:
    """A job-wide counter the ranks draw band numbers from (`Store.add` of the process group's rendezvous store is an
    atomic fetch-and-add served by rank 0's store thread; ~50 us per draw on one box).  Generation g of a call hands out
    tickets [g*stride, (g+1)*stride): an overflow re-run needs no reset.  Without a store (a backend that has none)
    ``next`` deals the static round-robin share instead."""

    def __init__(self, n_items: int, rank: int, world: int, tag: str = ""):
        """``tag``: names the process group (its global ranks), so that two groups searching at the same time do not
        draw from each other's counter."""
        global _ticket_calls
        _ticket_calls += 1
        self.n, self.rank, self.world = n_items, rank, world
        self.stride = n_items + world  # every rank draws one ticket past the end before it stops
        self.gen = -1
        self.key = f"b2c/dedup_tickets/{tag}/{_ticket_calls}"
        try:
            from torch.distributed.distributed_c10d import _get_default_store
            self.store = _get_default_store()
            if rank == 0 and _ticket_calls > 1:  # every rank has left the previous call: its counter can go
                try:
                    self.store.delete_key(f"b2c/dedup_tickets/{tag}/{_ticket_calls - 1}")
                except Exception:  # noqa: BLE001 — not every store type can delete
                    pass
        except Exception:  # noqa: BLE001
            self.store = None
        self._static = 0

    def new_generation(self):
        self.gen += 1
        self._static = self.rank

    def next(self) -> int:
        """Next band index of this generation, or -1 when they are all taken."""
        if self.store is None:
            i, self._static = self._static, self._static + self.world
        else:
            i = int(self.store.add(self.key, 1)) - 1 - self.gen * self.stride
        return i if 0 <= i < self.n else -1


def duplicate_pairs_distributed(local_embeddings: torch.Tensor, threshold: float, compare: str = "ref_fp16",
                                group=None, capacity: int | None = None):
    """Multi-GPU form: every rank passes its shard [n_local, E] (equal n_local on all ranks; pad the last shard
    with zero rows, which never match).  Each rank normalises its shard straight into its slice of the gather buffer;
    ONE all-gather (NCCL) of the shards then runs on NCCL's stream WHILE the rank searches the block of its own shard
    on a side stream; after the gather the ranks draw the remaining bands (largest first) from a shared counter, a few
    at a time, so the work follows the GPUs' actual speed.  Counts and pair buffers come back through two fixed-shape
    all-gathers (no pickling).  Every rank returns the full sorted result."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = local_embeddings.device  # normalize_rows_f16 refuses anything but a CUDA tensor: there is no CPU fallback
    n_local, E = local_embeddings.shape
    E_pad = (E + 63) // 64 * 64
    n_total = world * n_local
    gathered = torch.empty(n_total, E_pad, dtype=torch.float16, device=dev)
    mine = gathered[rank * n_local:(rank + 1) * n_local]
    normalize_rows_f16(local_embeddings, out=mine)
    lo = rank * n_local
    local_blocks = [(lo, lo + n_local, lo, lo + n_local)] if n_local > 1 else []
    # everything right of the shards' diagonal blocks, as bands, largest first: the ranks DRAW them from a shared counter,
    # so a GPU running at a lower power-capped clock simply takes fewer and all ranks finish together.  (The list is built
    # directly: going through the static deal of owned_blocks for every rank cost 10-17 ms of Python per call at N = 8.)
    bands = bands_right_of_diagonal(n_local, world)
    try:
        tag = "-".join(str(r) for r in dist.get_process_group_ranks(group if group is not None else dist.group.WORLD))
    except Exception:  # noqa: BLE001
        tag = "world"
    tickets = BandTickets(len(bands), rank, world, tag)
    cap = capacity if capacity is not None else max(1 << 16, 4 * n_total // world)
    first = True
    while True:
        buf = torch.empty(cap, 3, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        tickets.new_generation()
        if first:
            # the gather only WRITES the peers' slices; this rank's slice is read by both the gather and the local search
            work = dist.all_gather_into_tensor(gathered, mine, group=group, async_op=True)
            launch_pair_search(gathered, local_blocks, float(threshold), compare, buf, cnt)
            work.wait()  # stream-level: what follows on this stream is ordered after the gather, the host does not block
            first = False
        else:
            launch_pair_search(gathered, local_blocks, float(threshold), compare, buf, cnt)
        inflight = []
        while True:
            i = tickets.next()
            if i < 0:
                break
            launch_pair_search(gathered, [bands[i]], float(threshold), compare, buf, cnt)
            inflight.append(launch_marker())
            if len(inflight) > 2:  # keep two bands queued behind the running one, draw the next when one retires
                inflight.pop(0).synchronize()
        # counts of all ranks (one small collective), then every rank's pairs in fixed-capacity slots
        counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, cnt, group=group)
        counts_h = counts.cpu().numpy()
        k_max = int(counts_h.max())
        if k_max <= cap:
            break
        cap = int(k_max * 1.25) + 1024  # some rank overflowed: everybody re-runs with room for the largest list
    slot = max(int(k_max), 1)
    allp = torch.empty(world * slot, 3, dtype=torch.int32, device=dev)  # concatenation along dim 0 (what gloo accepts too)
    dist.all_gather_into_tensor(allp, buf[:slot].contiguous(), group=group)
    allp_h = allp.cpu().numpy().reshape(world, slot, 3)
    raw = np.concatenate([allp_h[r, :int(counts_h[r])] for r in range(world)]) if world else allp_h.reshape(0, 3)
    pairs, sims = _unpack(raw)
    return sort_pairs(pairs, sims)


# ----------------------------------------------------------------------------------------- reference entry
def find_near_duplicates(args, sim_type="cosine", crop_to_use="square_padded_crop"):
    """_2_remove_duplicates.py:52-99 with the similarity/threshold/where core on the GPU kernel.  Returns, per chunk,
    (near_duplicates [(path_i, path_j)], near_duplicate_values [float]) — the two lists the reference builds (:77,80)."""
    if sim_type not in ("cosine", "euclidean"):
        raise ValueError(f"sim_type must be 'cosine' or 'euclidean', got {sim_type!r}")
    compare = "ref_fp16" if sim_type == "cosine" else "euclidean"
    # the folder sits next to root_dir and exists even for a dry run (:83-84)
    output_dir = os.path.join(os.path.dirname(args.root_dir), f"near_duplicates_{sim_type}_{args.threshold}")
    results = []
    for paths, embeddings in get_paths_and_embeddings(args, crop_to_use):
        if not paths or not embeddings:
            continue
        stacked = torch.stack(embeddings)
        print(f"Searching {tuple(stacked.shape)} embeddings for pairs above {args.threshold} ({sim_type})..")
        pairs, sims = duplicate_pairs(stacked, args.threshold, compare)
        near_duplicates = [(paths[i], paths[j]) for i, j in pairs.tolist()]
        near_duplicate_values = [float(np.float16(v)) for v in sims]  # the reference reads them from its fp16 matrix (:80)
        os.makedirs(output_dir, exist_ok=True)
        print(f"Found {len(near_duplicates)} duplicates!")
        if near_duplicates and not args.test:
            print(f"{'copying' if args.mode == 'copy' else 'moving'} {len(near_duplicates)} near duplicates to {output_dir}...")
            for index, (pair, value) in enumerate(zip(near_duplicates, near_duplicate_values)):
                fix_duplicate(index, pair, output_dir, value, args.mode)
            print(f"{'Copied' if args.mode == 'copy' else 'Moved'} them to {output_dir}"
                  + (" (nothing was removed from the data yet)" if args.mode == "copy" else ""))
        results.append((near_duplicates, near_duplicate_values))
    return results


def find_near_duplicates_in_store(store, threshold=0.96, crop_to_use="square_padded_crop", per_directory=True,
                                  compare="ref_fp16"):
    """The same search fed from a packed store (store.PackedStore) instead of one torch.load per
    image (_2_remove_duplicates.py:25-46).  ``per_directory=True`` keeps the reference's scope — duplicates are only
    looked for among the images of one directory (:10) — and compares them in sorted-path order; ``False`` searches the
    whole store at once.  Embeddings go through fp16 like the reference's loader (:38).  Returns
    [(near_duplicates [(path_i, path_j)], near_duplicate_values [float])] per group."""
    usable = store.has_all([crop_to_use])
    in_order = store.paths_sorted()  # the embedding run writes a shard in sorted-path order and says so in its index
    paths = store.paths if (per_directory or not in_order) else None  # whole-store search: only the found pairs' paths
    if per_directory:
        groups = {}
        for i, p in enumerate(paths):
            if usable[i]:
                groups.setdefault(os.path.dirname(p), []).append(i)
    else:
        groups = {"": np.nonzero(usable)[0]}
    arr = store.array()  # [N, C, E] memory map
    ci = store.crop_names.index(crop_to_use)
    src_dtype = torch.float16 if arr.dtype == np.float16 else torch.float32
    results = []
    for key in sorted(groups):
        idx = np.asarray(groups[key] if in_order else sorted(groups[key], key=paths.__getitem__), np.int64)
        if len(idx) < 2:
            results.append(([], []))
            continue
        contiguous = bool((np.diff(idx) == 1).all())

        def rows_into(a, b, out, idx=idx, contiguous=contiguous):
            if contiguous:
                out[...] = arr[int(idx[a]):int(idx[a]) + (b - a), ci, :]
            else:
                np.take(arr[:, ci, :], idx[a:b], axis=0, out=out)

        if len(idx) >= STREAM_MIN_ROWS:  # large group: host gather, H2D and search overlap
            pairs, sims = duplicate_pairs_streamed(rows_into, len(idx), arr.shape[2], src_dtype, threshold, compare)
        else:
            host = np.empty((len(idx), arr.shape[2]), arr.dtype)
            rows_into(0, len(idx), host)
            pairs, sims = duplicate_pairs(torch.from_numpy(host).to(torch.float16), threshold, compare)
        found = store.paths_at(idx[pairs.reshape(-1)]) if len(pairs) else []
        results.append((list(zip(found[0::2], found[1::2])), [float(np.float16(s)) for s in sims]))
    return results


def find_near_duplicates_in_store_distributed(store_dir, threshold=0.96, crop_to_use="square_padded_crop", model_name=None,
                                              compare="ref_fp16", group=None):
    """Whole-store search with one process per GPU: rank r opens ONLY shard r of the packed store (the shard rank r of
    the embedding run wrote), the shards are all-gathered once on the device and searched like
    ``duplicate_pairs_distributed``.  Shards are padded with zero rows to the longest (zero rows never match).
    Every rank returns (near_duplicates [(path_i, path_j)], near_duplicate_values [float]) over all shards."""
    import torch.distributed as dist
    from .store import PackedStore
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    store = PackedStore(store_dir, model_name, shards=[rank])
    arr = store.array()  # [n, C, E] memory map of this rank's shard
    ci = store.crop_names.index(crop_to_use)
    n_mine, E = int(arr.shape[0]), int(arr.shape[2])
    dev = torch.device("cuda", torch.cuda.current_device())
    n_all = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(n_all, torch.tensor([n_mine], dtype=torch.int64, device=dev), group=group)
    n_local = max(n_all.cpu().tolist())
    local = torch.zeros(n_local, E, dtype=torch.float16, device=dev)
    if n_mine:
        # the crop's column goes to the device through the pinned slots of the streamed search (four gather threads, H2D
        # of chunk k under the gather of chunk k+1) and is rounded to fp16 there, like the reference's loader (_2:38) —
        # a single-threaded copy out of the memory map plus a pageable H2D copy took longer than the search itself at N = 8
        def rows_into(a, b, out):
            out[...] = arr[a:b, ci, :]

        def consume(a, b, d):
            local[a:b].copy_(d)

        src_dtype = torch.float16 if arr.dtype == np.float16 else torch.float32
        # at least four chunks per shard (gather / copy overlap), and pinned slots no larger than a small shard needs:
        # every rank of the box allocates its two slots at the same time on the first call
        chunk_rows = min(1 << 16, max(8192, -(-n_mine // 4)))
        with torch.cuda.device(dev):
            side = _stage_rows(rows_into, n_mine, E, src_dtype, dev, chunk_rows, consume)
            torch.cuda.current_stream(dev).wait_stream(side)
        usable = store.has_all([crop_to_use])
        if not usable.all():  # images without this crop take no part
            local[torch.from_numpy(np.nonzero(~usable)[0]).to(dev)] = 0
    pairs, sims = duplicate_pairs_distributed(local, threshold, compare, group=group)
    # global row -> path: row = shard * n_local + index inside the shard; only the paths of rows that occur are exchanged
    need = sorted({int(r) for r in pairs.reshape(-1)})
    mine_rows = [r for r in need if r // n_local == rank]
    mine = dict(zip(mine_rows, store.paths_at([r - rank * n_local for r in mine_rows])))
    merged = [None] * world
    dist.all_gather_object(merged, mine, group=group)  # a few path strings per found pair, not the embeddings
    path_of = {k: v for part in merged for k, v in part.items()}
    return ([(path_of[int(i)], path_of[int(j)]) for i, j in pairs.tolist()], [float(np.float16(s)) for s in sims])


def _files_sharing_stem(img_path):
    """Every file of the image's directory whose name CONTAINS the image's stem (the reference's substring match, :111-112:
    the image, its .pt, its .json, ... and, as there, anything else that happens to contain it)."""
    folder = os.path.dirname(img_path)
    stem = os.path.splitext(os.path.basename(img_path))[0]
    return [os.path.join(folder, name) for name in os.listdir(folder) if stem in name]


def fix_duplicate(duplicate_index, img_paths, outdir, sim_value, mode):
    """_2_remove_duplicates.py:102-125.  'copy': both images' files are copied for inspection; 'move': only the second
    image (the 'target') leaves the dataset.  Names: ``{sim:.3f}_{index:08d}_{source|target}_{original name}``."""
    tag = f"{sim_value:.3f}_{duplicate_index:08d}"
    for role, img_path in (("source", img_paths[0]), ("target", img_paths[1])):
        for src in _files_sharing_stem(img_path):
            dst = os.path.join(outdir, f"{tag}_{role}_{os.path.basename(src)}")
            if mode == "copy":
                shutil.copy(src, dst)
            elif mode == "move" and role == "target":
                os.rename(src, dst)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Near-duplicate search over CLIP embeddings (flags of _2_remove_duplicates.py:135-142)")
    ap.add_argument("--root_dir", type=str, help="Root directory of the dataset")
    ap.add_argument("--threshold", type=float, default=0.96, help="Cosine-similarity threshold above which two images are duplicates")
    ap.add_argument("--mode", type=str, default="copy", help="copy (inspect first) / move (take the duplicates out)")
    ap.add_argument("--clip_model_to_use", type=str, default=None, help="CLIP model key inside the .pt files; default: the first one found")
    ap.add_argument("--chunk_size", type=int, default=10_000_000,
                    help="Most embeddings compared at once per directory (the reference's 10000 was a memory cap; lifted here)")
    ap.add_argument("--test", action="store_true", help="Dry run: report, do not copy or move anything")
    find_near_duplicates(ap.parse_args(argv))


if __name__ == "__main__":
    main()
