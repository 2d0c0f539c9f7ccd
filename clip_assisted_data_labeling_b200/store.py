"""Packed embedding store — SURVEY.md §8f row 1.

The reference keeps one ``<img>.pt`` pickle per image (written at _1_embed_with_CLIP.py:146-170, re-read one file
at a time by _2_remove_duplicates.py:25-46, _4_train_model.py:42-75, _5_predict_labels.py:69-88 and
tools/find_similar_imgs.py:36-46).  Once the embedding pass runs at >1000 images/s those 14 KB pickles *are* the
wall time, so the B200 path writes, per rank, one flat shard instead:

    <store_dir>/shard-00000.emb          raw little-endian f32 (or f16) [n, C, E], C = len(crop_names), row-major
    <store_dir>/shard-00000.paths        the image paths, UTF-8, one per line      } per-image lists as sidecar files: a
    <store_dir>/shard-00000.kept         uint8 bitmask per image                   } million rows parse in 0.05 s
    <store_dir>/shard-00000.stats        raw f32 [n, S]: the S = 22 ``img_stat_*`` scalars of every image, in the order of
                                         "stat_names" (only when the embedding run computed them; _1:149-152)
    <store_dir>/shard-00000.json         {"format", "model", "crop_names", "dtype", "embed", "count", "sidecars": true,
                                          "stat_names": [...] | absent}
                                         — written LAST: the commit point.  (A path with a line break in it, or an
                                         older shard, keeps "paths": [...] and "kept": [...] inside the JSON.)

``kept`` bit c is 0 when the reference would have dropped crop c as empty (utils/embedder.py:243-247); the row is
all zeros then and the compat exporter omits the key.  Shards are memory-mapped by the readers, so a whole
directory tree loads as one ``[N, C, E]`` array with no per-image ``torch.load``; ``export_pt`` writes the
per-image ``.pt`` files lazily, in exactly the layout the reference's consumers read, and
``import_pt`` is the bulk loader in the other direction (existing ``.pt`` trees -> one shard, thread pool).

Host-side file I/O only: nothing here touches CUDA.
"""
from __future__ import annotations

import concurrent.futures
import glob
import json
import os

import numpy as np
import torch

from .vit_arch import CROP_NAMES

FORMAT = "b2c-packed-embeddings/1"
_DTYPES = {"float32": np.float32, "float16": np.float16}


def _shard_paths(store_dir: str, shard: int):
    base = os.path.join(store_dir, f"shard-{shard:05d}")
    return base + ".emb", base + ".json"


class PackedWriter:
    """Append-only writer of one shard (one per rank).  ``append`` takes the ``[B, 4, E]`` block
    ``CLIP_Encoder.encode_images_u8`` returns (already on the host), in CROP_NAMES order, and keeps the columns of
    ``crop_names``.  The index is written on ``close()``; a shard without its ``.json`` is ignored by the reader,
    so a crashed run never leaves a half-valid shard behind."""

    def __init__(self, store_dir: str, model_name: str, embed: int, crop_names=CROP_NAMES, shard: int = 0,
                 dtype: str = "float32", weights_source: str | None = None, stat_names=None):
        """``stat_names``: names of the per-image statistics (imgstats.STAT_NAMES) every ``append`` will bring along; they
        go to the shard's ``.stats`` file so that ``export_pt`` writes the complete .pt layout."""
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        os.makedirs(store_dir, exist_ok=True)
        self.store_dir, self.model_name, self.embed = store_dir, model_name, int(embed)
        self.crop_names = list(crop_names)
        self._cols = [CROP_NAMES.index(c) for c in self.crop_names]
        self.dtype = dtype
        self.weights_source = weights_source  # where the encoder's weights came from (CLIP_Encoder.weights_source)
        self.emb_path, self.idx_path = _shard_paths(store_dir, shard)
        if os.path.exists(self.idx_path):
            os.remove(self.idx_path)  # invalidate first, then rewrite the data
        self._fh = open(self.emb_path, "wb")
        self.stat_names = list(stat_names) if stat_names else []
        self.stats_path = self.idx_path[:-5] + ".stats"
        self._sfh = open(self.stats_path, "wb") if self.stat_names else None
        if self._sfh is None and os.path.exists(self.stats_path):
            os.remove(self.stats_path)  # left by an earlier run of this shard that had statistics
        self.paths: list[str] = []
        self.kept: list[int] = []
        self._sorted = True  # paths appended so far are in ascending order (recorded in the index: readers skip the check)

    def append(self, features, paths, kept=None, stats=None) -> None:
        """features: [B, 4, E] (torch CPU tensor or numpy) in CROP_NAMES order; paths: B image paths;
        kept: optional B x 4 booleans (False = crop dropped as empty); stats: [B, len(stat_names)] when the writer was
        opened with ``stat_names`` (and only then)."""
        f = features.detach().cpu().numpy() if isinstance(features, torch.Tensor) else np.asarray(features)
        if f.ndim != 3 or f.shape[1] != len(CROP_NAMES) or f.shape[2] != self.embed or f.shape[0] != len(paths):
            raise ValueError(f"expected [{len(paths)},{len(CROP_NAMES)},{self.embed}], got {f.shape}")
        if (stats is not None) != bool(self.stat_names):
            raise ValueError("statistics must accompany every append of a writer opened with stat_names, and no other")
        if stats is not None:
            st = stats.detach().cpu().numpy() if isinstance(stats, torch.Tensor) else np.asarray(stats)
            if st.shape != (len(paths), len(self.stat_names)):
                raise ValueError(f"expected statistics [{len(paths)},{len(self.stat_names)}], got {st.shape}")
            self._sfh.write(np.ascontiguousarray(st, dtype=np.float32))
        # (the driver's case — every crop, f32 in, f32 stored, nothing dropped — writes the caller's block as it is)
        if self._cols != list(range(len(CROP_NAMES))):
            f = f[:, self._cols, :]
        if kept is None:
            self.kept.extend([(1 << len(self._cols)) - 1] * len(paths))
        else:
            k = np.asarray(kept, dtype=bool).reshape(len(paths), len(CROP_NAMES))[:, self._cols]
            self.kept.extend((k.astype(np.int64) << np.arange(len(self._cols))).sum(axis=1).tolist())
            if not k.all():
                f = np.array(f, dtype=_DTYPES[self.dtype])  # a private copy: the dropped crops are stored as zeros
                f[~k] = 0
        f = np.ascontiguousarray(f, dtype=_DTYPES[self.dtype])
        self._fh.write(f)  # (buffer protocol: no intermediate bytes object)
        new = [os.fspath(p) for p in paths]
        if self._sorted and new:
            prev = self.paths[-1] if self.paths else None
            for q in new:
                if prev is not None and q < prev:
                    self._sorted = False
                    break
                prev = q
        self.paths.extend(new)

    def close(self) -> str:
        self._fh.flush()
        os.fsync(self._fh.fileno())
        self._fh.close()
        meta = {"format": FORMAT, "model": self.model_name, "crop_names": self.crop_names, "dtype": self.dtype,
                "embed": self.embed, "count": len(self.paths), "weights_source": self.weights_source,
                "sorted": self._sorted}
        if self._sfh is not None:
            self._sfh.flush()
            os.fsync(self._sfh.fileno())
            self._sfh.close()
            meta["stat_names"] = self.stat_names
        # The per-image lists go to sidecar files (a million paths parse in 0.05 s as text against 0.3 s as JSON); the
        # .json stays the commit point and is written last.  Paths with a line break in them keep the JSON form.
        base = self.idx_path[:-5]
        if self.paths and not any("\n" in p for p in self.paths):
            with open(base + ".paths", "w", encoding="utf-8", newline="\n") as fh:
                fh.write("\n".join(self.paths))
            np.asarray(self.kept, dtype=np.uint8).tofile(base + ".kept")
            meta["sidecars"] = True
        else:
            meta["paths"] = self.paths
            meta["kept"] = self.kept
        tmp = self.idx_path + ".tmp"
        with open(tmp, "w") as fh:
            json.dump(meta, fh)
        os.replace(tmp, self.idx_path)
        return self.idx_path

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            self.close()
        else:
            self._fh.close()
            if self._sfh is not None:
                self._sfh.close()


class PackedStore:
    """Read side: every complete shard of ``store_dir`` for ``model_name`` (default: the first model found, like
    _2_remove_duplicates.py:32-34 defaults to the first key), memory-mapped."""

    def __init__(self, store_dir: str, model_name: str | None = None, shards=None):
        """``shards``: optional shard numbers to open (default: all) — one process per GPU reads only its own."""
        self.store_dir = store_dir
        self.shards = []
        self.model_name = model_name
        index_files = sorted(glob.glob(os.path.join(store_dir, "shard-*.json")))
        if shards is not None:
            wanted = {_shard_paths(store_dir, int(k))[1] for k in shards}
            index_files = [f for f in index_files if f in wanted]
        for idx_path in index_files:
            with open(idx_path) as fh:
                meta = json.load(fh)
            if meta.get("format") != FORMAT:
                raise ValueError(f"{idx_path}: unknown store format {meta.get('format')!r}")
            if self.model_name is None:
                self.model_name = meta["model"]
            if meta["model"] != self.model_name:
                continue
            C, E, n = len(meta["crop_names"]), meta["embed"], meta["count"]
            emb_path = idx_path[:-5] + ".emb"
            want = n * C * E * np.dtype(_DTYPES[meta["dtype"]]).itemsize
            have = os.path.getsize(emb_path) if os.path.exists(emb_path) else -1
            if have != want:
                raise ValueError(f"{emb_path}: {have} bytes on disk, index says {want}")
            arr = (np.memmap(emb_path, dtype=_DTYPES[meta["dtype"]], mode="r", shape=(n, C, E)) if n
                   else np.zeros((0, C, E), _DTYPES[meta["dtype"]]))
            if meta.get("sidecars"):
                # the path list stays raw bytes until somebody asks for ``.paths``: a million str objects take ~0.1 s to
                # create, and the duplicate search only needs the paths of the pairs it finds (``paths_at``)
                with open(idx_path[:-5] + ".paths", "rb") as fh:
                    meta["_raw_paths"] = fh.read()
                meta["kept"] = np.fromfile(idx_path[:-5] + ".kept", dtype=np.uint8)
                n_paths = meta["_raw_paths"].count(b"\n") + 1 if n else 0
                if n_paths != n or len(meta["kept"]) != n:
                    raise ValueError(f"{idx_path}: sidecar files hold {n_paths} paths / {len(meta['kept'])} masks, index says {n}")
            if meta.get("stat_names"):
                S, stats_path = len(meta["stat_names"]), idx_path[:-5] + ".stats"
                have = os.path.getsize(stats_path) if os.path.exists(stats_path) else -1
                if have != n * S * 4:
                    raise ValueError(f"{stats_path}: {have} bytes on disk, index says {n * S * 4}")
                meta["_stats"] = np.memmap(stats_path, dtype=np.float32, mode="r", shape=(n, S)) if n else np.zeros((0, S), np.float32)
            self.shards.append((meta, arr))
        if not self.shards:
            raise FileNotFoundError(f"no packed shards for model {model_name!r} under {store_dir}")
        first = self.shards[0][0]
        self.crop_names = list(first["crop_names"])
        self.embed = first["embed"]
        for meta, _ in self.shards:
            if meta["crop_names"] != self.crop_names or meta["embed"] != self.embed:
                raise ValueError("shards of one model disagree on crop_names / embed")
        self._paths = None
        self._first_row = np.cumsum([0] + [meta["count"] for meta, _ in self.shards])  # global row of each shard's first
        self.kept = np.concatenate([np.asarray(meta["kept"], dtype=np.int64) for meta, _ in self.shards])
        # image statistics: present only when EVERY shard of the model carries the same list
        names = [tuple(meta.get("stat_names") or ()) for meta, _ in self.shards]
        self.stat_names = list(names[0]) if names[0] and all(nm == names[0] for nm in names) else []

    def __len__(self):
        return int(self._first_row[-1])

    @staticmethod
    def _shard_paths_list(meta) -> list:
        if "paths" not in meta:
            meta["paths"] = meta["_raw_paths"].decode("utf-8").split("\n") if meta["count"] else []
        return meta["paths"]

    @property
    def paths(self) -> list:
        """All image paths, in row order (built on first use)."""
        if self._paths is None:
            self._paths = [p for meta, _ in self.shards for p in self._shard_paths_list(meta)]
        return self._paths

    def paths_at(self, rows) -> list:
        """Paths of the given global rows without materialising the whole list: the raw sidecar bytes are cut at the line
        breaks around each requested row (one vectorised newline scan per shard, on first use)."""
        rows = np.asarray(rows, np.int64).reshape(-1)
        if rows.size and (rows.min() < 0 or rows.max() >= len(self)):
            raise IndexError(f"row out of range for a store of {len(self)} images")
        if self._paths is not None:
            return [self._paths[int(r)] for r in rows]
        out = [None] * len(rows)
        which = np.searchsorted(self._first_row, rows, side="right") - 1
        for s in np.unique(which):
            meta = self.shards[int(s)][0]
            sel = np.nonzero(which == s)[0]
            local = rows[sel] - self._first_row[s]
            if "paths" in meta or "_raw_paths" not in meta:
                lst = self._shard_paths_list(meta)
                for k, r in zip(sel, local):
                    out[k] = lst[int(r)]
                continue
            if "_line_starts" not in meta:
                raw = np.frombuffer(meta["_raw_paths"], np.uint8)
                meta["_line_starts"] = np.concatenate([[0], np.flatnonzero(raw == 10) + 1, [raw.size + 1]])
            st, rawb = meta["_line_starts"], meta["_raw_paths"]
            for k, r in zip(sel, local):
                out[k] = rawb[int(st[r]):int(st[r + 1]) - 1].decode("utf-8")
        return out

    def paths_sorted(self) -> bool:
        """True when the rows are in ascending path order.  Shards written by this version record it in their index
        (``"sorted"``); only older shards, or several shards, need a look at the paths themselves."""
        if self._paths is None and all(meta.get("sorted") for meta, _ in self.shards):
            if len(self.shards) == 1:
                return True
            edges = [self.paths_at([self._first_row[k], self._first_row[k + 1] - 1]) for k, (meta, _) in enumerate(self.shards)
                     if meta["count"]]
            return all(a[1] <= b[0] for a, b in zip(edges, edges[1:]))
        return sorted(self.paths) == self.paths

    def array(self) -> np.ndarray:
        """[N, C, E] over all shards (a view for one shard, one concatenation otherwise)."""
        if len(self.shards) == 1:
            return self.shards[0][1]
        return np.concatenate([a for _, a in self.shards], axis=0)

    def stats(self):
        """f32 [N, S] image statistics over all shards in ``stat_names`` order, or None when the store has none."""
        if not self.stat_names:
            return None
        if len(self.shards) == 1:
            return self.shards[0][0]["_stats"]
        return np.concatenate([meta["_stats"] for meta, _ in self.shards], axis=0)

    def crop(self, crop_name: str, dtype=torch.float32, device=None) -> torch.Tensor:
        """[N, E] embeddings of one crop, the bulk form of ``d[model][crop].squeeze()`` (_2_remove_duplicates.py:38)."""
        ci = self.crop_names.index(crop_name)
        t = torch.from_numpy(np.array(self.array()[:, ci, :])).to(dtype)  # (a copy: the memory map is read-only)
        return t.to(device) if device is not None else t

    def features(self, crop_names=None, device=None, with_stats: bool = False) -> torch.Tensor:
        """[N, len(crop_names) * E] f32: the regressor's input row per image, crops concatenated in the order given
        (_4_train_model.py:55, _5_predict_labels.py:78-79); ``with_stats`` appends the image statistics after the crops
        (``use_img_stat_features``, _4_train_model.py:60-63)."""
        names = list(crop_names) if crop_names is not None else self.crop_names
        idx = [self.crop_names.index(c) for c in names]
        t = torch.from_numpy(np.array(self.array()[:, idx, :])).float().reshape(len(self), -1)
        if with_stats:
            if not self.stat_names:
                raise ValueError("this packed store holds no image statistics (the embedding run had img_stats off)")
            t = torch.cat([t, torch.from_numpy(np.array(self.stats()))], dim=1)
        return t.to(device) if device is not None else t

    def has_all(self, crop_names) -> np.ndarray:
        """bool [N]: image has every crop of ``crop_names`` (the reference raises/skips on a missing crop, _4:56-58)."""
        need = 0
        for c in crop_names:
            need |= 1 << self.crop_names.index(c)
        return (self.kept & need) == need

    def feature_dict(self, i: int) -> dict:
        """{img_stat_*: f32 0-d, crop_name: f32[1,E]} of image i, the per-model dict of the reference's .pt layout
        (statistics ahead of the crops like _1:149-161 builds it)."""
        row, srow, off = None, None, i
        for meta, arr in self.shards:
            if off < meta["count"]:
                row = arr[off]
                if self.stat_names:
                    srow = torch.from_numpy(np.array(meta["_stats"][off], dtype=np.float32))
                break
            off -= meta["count"]
        d = {n: srow[k] for k, n in enumerate(self.stat_names)} if srow is not None else {}
        for ci, name in enumerate(self.crop_names):
            if (int(self.kept[i]) >> ci) & 1:
                d[name] = torch.from_numpy(np.array(row[ci], dtype=np.float32)).unsqueeze(0)
        return d


def pt_path_for(img_path: str) -> str:
    return os.path.splitext(img_path)[0] + ".pt"


def _merge_save(pt_path: str, model_name: str, feature_dict: dict, force: bool) -> None:
    final = {}
    if os.path.exists(pt_path) and not force:
        try:
            final = torch.load(pt_path, map_location="cpu")
        except Exception as e:  # noqa: BLE001
            print(f"Warning: Failed to load existing {pt_path} for update: {e}")
    final[model_name] = feature_dict
    torch.save(final, pt_path)


def export_pt(store: PackedStore, force_reencode: bool = False, threads: int = 8, only=None) -> int:
    """Compat exporter: write ``<img>.pt`` next to each image in the reference's layout
    ``{model: {img_stat_*: f32 0-d (when the store has them), crop: f32[1,E]}}``, merged into an existing file unless ``force_reencode``
    (_1_embed_with_CLIP.py:138-170).  ``only``: optional iterable of indices.  Returns the number of files written."""
    idx = list(range(len(store))) if only is None else list(only)
    with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as ex:
        futs = [ex.submit(_merge_save, pt_path_for(store.paths[i]), store.model_name, store.feature_dict(i), force_reencode)
                for i in idx]
        for f in futs:
            f.result()
    return len(idx)


def _load_one(img_path: str, model_name, crop_names):
    try:
        d = torch.load(pt_path_for(img_path), map_location="cpu")
        if model_name is None:
            model_name = list(d.keys())[0]
        fd = d[model_name]
        rows, mask, E = [], 0, None
        for ci, c in enumerate(crop_names):
            if c in fd:
                v = fd[c].reshape(-1).float().numpy()
                E = v.shape[0]
                rows.append(v)
                mask |= 1 << ci
            else:
                rows.append(None)
        if E is None:
            return None
        stat = {k: float(v) for k, v in fd.items() if k.startswith("img_stat_")}  # in file order, like _4:61
        return model_name, np.stack([r if r is not None else np.zeros(E, np.float32) for r in rows]), mask, stat
    except Exception:  # noqa: BLE001  (the reference skips unreadable samples: _2:45-46, _4:72-74)
        return None


def import_pt(root_dir: str, store_dir: str, model_name: str | None = None, crop_names=CROP_NAMES, threads: int = 16,
              img_extensions=(".jpg",), shard: int = 0) -> PackedStore:
    """Bulk loader: walk ``root_dir`` for images that have a ``.pt`` companion (the pairing rule of
    _2_remove_duplicates.py:27), read the pickles on a thread pool and pack them into one shard.  Images are taken
    in sorted path order so the result is deterministic."""
    imgs = []
    for sub, _dirs, files in os.walk(root_dir):
        names = set(files)
        for f in sorted(files):
            stem, ext = os.path.splitext(f)
            if ext in img_extensions and stem + ".pt" in names:
                imgs.append(os.path.join(sub, f))
    imgs.sort()
    crop_names = list(crop_names)
    with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as ex:
        loaded = list(ex.map(lambda p: _load_one(p, model_name, crop_names), imgs))
    good = [(p, r) for p, r in zip(imgs, loaded) if r is not None]
    if not good:
        raise FileNotFoundError(f"no readable .pt embeddings under {root_dir}")
    model = model_name or good[0][1][0]
    good = [(p, r) for p, r in good if r[0] == model]
    E = good[0][1][1].shape[1]
    full = np.zeros((len(good), len(CROP_NAMES), E), np.float32)
    kept = []
    cols = [CROP_NAMES.index(c) for c in crop_names]
    for b, (_, (_, rows, mask, _stat)) in enumerate(good):
        full[b, cols, :] = rows
        k = [False] * len(CROP_NAMES)
        for ci, c in enumerate(cols):
            k[c] = bool((mask >> ci) & 1)
        kept.append(k)
    # the statistics travel along when every file has the same list of them (files of the reference's _1 and of this
    # package's driver do; a tree embedded with img_stats off has none)
    stat_names = list(good[0][1][3])
    if not stat_names or any(list(r[3]) != stat_names for _, r in good):
        stat_names = []
    stats = np.asarray([[r[3][n] for n in stat_names] for _, r in good], np.float32) if stat_names else None
    with PackedWriter(store_dir, model, E, crop_names, shard=shard, stat_names=stat_names) as w:
        w.append(full, [p for p, _ in good], kept, stats=stats)
    return PackedStore(store_dir, model)
