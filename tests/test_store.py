"""Packed embedding store: round trip, the compat exporter writes exactly the layout the
reference's consumers read, the bulk loader inverts it, and — when /root/reference is present — the reference's
own `_2_remove_duplicates.get_paths_and_embeddings` reads the exported files."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import reference_present
from test_abi_and_host import _FakeEncoder, _write_images


def _fill(tmp_path, n=9, E=8, model="ViT-L-14/openai"):
    from clip_assisted_data_labeling_b200.store import PackedWriter
    rng = np.random.default_rng(3)
    root = tmp_path / "imgs"
    root.mkdir()
    feats = rng.standard_normal((n, 4, E)).astype(np.float32)
    paths = [str(root / f"{i:03d}.jpg") for i in range(n)]
    for p in paths:
        open(p, "wb").close()
    kept = [[True, True, True, i != 4] for i in range(n)]
    sd = str(tmp_path / "store")
    with PackedWriter(sd, model, E, shard=0) as w:
        w.append(feats[:5], paths[:5], kept[:5])
        w.append(torch.from_numpy(feats[5:]), paths[5:], kept[5:])
    return sd, feats, paths, kept


def test_round_trip_and_views(tmp_path):
    from clip_assisted_data_labeling_b200.store import PackedStore
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    sd, feats, paths, kept = _fill(tmp_path)
    st = PackedStore(sd)
    assert len(st) == 9 and st.paths == paths and st.model_name == "ViT-L-14/openai" and st.crop_names == CROP_NAMES
    want = feats.copy()
    want[4, 3] = 0  # dropped crop is stored as zeros
    assert np.array_equal(st.array(), want)
    assert torch.equal(st.crop("square_padded_crop"), torch.from_numpy(want[:, 1]))
    assert st.crop("subcrop1", torch.float16).dtype == torch.float16
    f = st.features(["centre_crop", "subcrop2"])
    assert f.shape == (9, 16) and torch.equal(f[2], torch.from_numpy(np.concatenate([want[2, 0], want[2, 3]])))
    assert st.has_all(CROP_NAMES).tolist() == [i != 4 for i in range(9)]
    assert st.has_all(["centre_crop"]).all()


def test_incomplete_shard_is_invisible_and_size_is_checked(tmp_path):
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
    sd = str(tmp_path / "s")
    w = PackedWriter(sd, "M/x", 4)
    w.append(np.ones((2, 4, 4), np.float32), ["a.jpg", "b.jpg"])
    with pytest.raises(FileNotFoundError):
        PackedStore(sd)  # no index yet: the shard does not exist as far as readers are concerned
    w.close()
    assert len(PackedStore(sd)) == 2
    with open(os.path.join(sd, "shard-00000.emb"), "ab") as fh:
        fh.write(b"x")
    with pytest.raises(ValueError):
        PackedStore(sd)


def test_sidecar_index_files_and_json_fallback(tmp_path):
    """Per-image lists live in sidecar files (fast to parse at a million rows); the .json is still the commit point, and
    a path that contains a line break keeps the lists inside the JSON."""
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
    sd = str(tmp_path / "s")
    names = ["a b.jpg", "ü/é.jpg", "c\rd.jpg"]
    with PackedWriter(sd, "M/x", 4, shard=0) as w:
        w.append(np.arange(48, dtype=np.float32).reshape(3, 4, 4), names, [[True] * 4, [True, False, True, True], [True] * 4])
    assert os.path.exists(os.path.join(sd, "shard-00000.paths")) and os.path.exists(os.path.join(sd, "shard-00000.kept"))
    st = PackedStore(sd)
    assert st.paths == names and st.kept.tolist() == [15, 13, 15]
    os.remove(os.path.join(sd, "shard-00000.paths"))
    with pytest.raises(FileNotFoundError):
        PackedStore(sd)
    with PackedWriter(sd, "M/x", 4, shard=0) as w:
        w.append(np.zeros((2, 4, 4), np.float32), ["line\nbreak.jpg", "x.jpg"])
    st = PackedStore(sd)
    assert st.paths == ["line\nbreak.jpg", "x.jpg"] and st.kept.tolist() == [15, 15]


def test_multi_shard_and_model_filter(tmp_path):
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
    sd = str(tmp_path / "s")
    for r in range(3):
        with PackedWriter(sd, "M/x", 4, shard=r) as w:
            w.append(np.full((2, 4, 4), r, np.float32), [f"{r}_{k}.jpg" for k in range(2)])
    st = PackedStore(sd, "M/x")
    assert len(st) == 6 and st.array()[:, 0, 0].tolist() == [0, 0, 1, 1, 2, 2]
    assert st.paths[2] == "1_0.jpg"
    assert torch.equal(st.feature_dict(3)["centre_crop"], torch.full((1, 4), 1.0))
    with pytest.raises(FileNotFoundError):
        PackedStore(sd, "other/model")


def test_export_matches_reference_layout_and_import_inverts_it(tmp_path):
    from clip_assisted_data_labeling_b200.store import PackedStore, export_pt, import_pt
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    sd, feats, paths, kept = _fill(tmp_path)
    st = PackedStore(sd)
    # an older file with another model's entry must be merged, not overwritten (_1_embed_with_CLIP.py:139-143)
    torch.save({"ViT-B-32/openai": {"centre_crop": torch.zeros(1, 3)}}, paths[0][:-4] + ".pt")
    assert export_pt(st) == 9
    d0 = torch.load(paths[0][:-4] + ".pt")
    assert list(d0.keys()) == ["ViT-B-32/openai", "ViT-L-14/openai"]
    d4 = torch.load(paths[4][:-4] + ".pt")["ViT-L-14/openai"]
    assert list(d4.keys()) == CROP_NAMES[:3]  # empty crop omitted (utils/embedder.py:243-247)
    d2 = torch.load(paths[2][:-4] + ".pt")["ViT-L-14/openai"]
    for ci, c in enumerate(CROP_NAMES):
        t = d2[c]
        assert t.dtype == torch.float32 and tuple(t.shape) == (1, 8) and torch.equal(t[0], torch.from_numpy(feats[2, ci]))
    back = import_pt(str(tmp_path / "imgs"), str(tmp_path / "store2"), "ViT-L-14/openai")
    assert back.paths == sorted(paths) and np.array_equal(back.array(), st.array()) and np.array_equal(back.kept, st.kept)


def test_feature_dataset_writes_packed_shard(tmp_path, lib):
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.store import PackedStore, export_pt
    root = str(tmp_path / "data")
    _write_images(root, 7)
    enc = _FakeEncoder()
    enc.embed_dim = 8
    sd = str(tmp_path / "packed")
    ds = Feature_Dataset(root, "ViT-L-14/openai", 4, shuffle_filenames=False, encoder=enc, packed_dir=sd, write_pt=False)
    assert ds.process() == (7, 0)
    assert not any(f.endswith(".pt") for f in os.listdir(root))
    st = PackedStore(sd)
    assert st.paths == sorted(ds.img_filepaths) and st.array().shape == (7, 4, 8)
    export_pt(st)
    # per-image files now equal what the direct path writes
    root2 = str(tmp_path / "data2")
    _write_images(root2, 7)
    Feature_Dataset(root2, "ViT-L-14/openai", 4, shuffle_filenames=False, encoder=_FakeEncoder()).process()
    for i in range(7):
        a = torch.load(os.path.join(root, f"im{i:03d}.pt"))
        b = torch.load(os.path.join(root2, f"im{i:03d}.pt"))
        assert a.keys() == b.keys()
        for c in a["ViT-L-14/openai"]:
            assert torch.equal(a["ViT-L-14/openai"][c], b["ViT-L-14/openai"][c])


@pytest.mark.skipif(not reference_present(), reason="/root/reference only exists in the build container")
def test_reference_dedup_loader_reads_exported_files(tmp_path):
    """The unmodified _2_remove_duplicates.get_paths_and_embeddings (reference :8-49) consumes export_pt's output."""
    from clip_assisted_data_labeling_b200.store import PackedStore, export_pt
    from oracle.reference_shim import import_reference
    sd, feats, paths, kept = _fill(tmp_path)
    st = PackedStore(sd)
    export_pt(st)
    ref = import_reference("_2_remove_duplicates")
    args = types.SimpleNamespace(root_dir=str(tmp_path / "imgs"), clip_model_to_use=None, chunk_size=100)
    got_paths, got_emb = next(iter(ref.get_paths_and_embeddings(args, "square_padded_crop")))
    order = [paths.index(p) for p in got_paths]
    assert sorted(order) == list(range(9))
    want = torch.from_numpy(feats[order, 1]).to(torch.float16)
    assert torch.equal(torch.stack(got_emb), want)


def _fake_stats(images, device=None):
    """Stand-in for imgstats.image_stats in host-logic tests (no CUDA): [B, 22] with the image's width, height and mean."""
    out = torch.zeros(len(images), 22, dtype=torch.float64)
    for i, im in enumerate(images):
        out[i, 0], out[i, 1], out[i, 2] = im.shape[1], im.shape[0], im.shape[1] / im.shape[0]
        out[i, 3:] = im.double().mean() + torch.arange(19)
    return out


def test_statistics_sidecar_round_trip_export_and_import(tmp_path):
    """A shard written with stat_names carries the 22 img_stat_* scalars: export_pt then writes the complete .pt layout
    (statistics ahead of the crops as 0-d f32 tensors, _1_embed_with_CLIP.py:149-161), import_pt brings them back, and a
    store whose shards disagree reports none."""
    from clip_assisted_data_labeling_b200.imgstats import STAT_NAMES
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter, export_pt, import_pt
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    rng = np.random.default_rng(5)
    root = tmp_path / "imgs"
    root.mkdir()
    n, E = 5, 8
    feats = rng.standard_normal((n, 4, E)).astype(np.float32)
    stats = rng.standard_normal((n, 22)).astype(np.float32)
    paths = [str(root / f"{i}.jpg") for i in range(n)]
    for p in paths:
        open(p, "wb").close()
    sd = str(tmp_path / "s")
    with PackedWriter(sd, "M/x", E, stat_names=STAT_NAMES) as w:
        w.append(feats[:2], paths[:2], stats=stats[:2])
        w.append(feats[2:], paths[2:], stats=torch.from_numpy(stats[2:]))
        with pytest.raises(ValueError):
            w.append(feats[:1], paths[:1])  # statistics must accompany every append
    st = PackedStore(sd)
    assert st.stat_names == STAT_NAMES and np.array_equal(st.stats(), stats)
    f = st.features(["centre_crop"], with_stats=True)
    assert f.shape == (n, E + 22) and torch.equal(f[3, E:], torch.from_numpy(stats[3]))
    export_pt(st)
    d = torch.load(paths[1][:-4] + ".pt")["M/x"]
    assert list(d.keys()) == STAT_NAMES + CROP_NAMES
    for k, name in enumerate(STAT_NAMES):
        assert d[name].dim() == 0 and d[name].dtype == torch.float32 and float(d[name]) == float(stats[1, k])
    back = import_pt(str(root), str(tmp_path / "s2"), "M/x")
    assert back.stat_names == STAT_NAMES and np.array_equal(back.stats(), stats) and np.array_equal(back.array(), feats)
    # a second shard without statistics: the store as a whole has none, and says so when they are asked for
    with PackedWriter(sd, "M/x", E, shard=1) as w:
        w.append(feats[:1], ["other.jpg"])
    st2 = PackedStore(sd)
    assert st2.stat_names == [] and st2.stats() is None and list(st2.feature_dict(0).keys()) == CROP_NAMES
    with pytest.raises(ValueError, match="no image statistics"):
        st2.features(["centre_crop"], with_stats=True)
    # rewriting a shard without statistics removes the stale .stats file; a truncated one is refused
    with PackedWriter(sd, "M/x", E, shard=0) as w:
        w.append(feats[:1], paths[:1])
    assert not os.path.exists(os.path.join(sd, "shard-00000.stats"))
    with PackedWriter(sd, "M/x", E, shard=0, stat_names=STAT_NAMES) as w:
        w.append(feats, paths, stats=stats)
    with open(os.path.join(sd, "shard-00000.stats"), "ab") as fh:
        fh.write(b"1234")
    with pytest.raises(ValueError, match="stats"):
        PackedStore(sd)


def test_driver_stores_statistics_and_resume_refills_the_shard(tmp_path, lib, monkeypatch):
    """Feature_Dataset with packed_dir: the statistics land in the shard and in the .pt files; a resumed run (every image
    already has this model's key) re-embeds nothing and still leaves a COMPLETE shard, rows read back from the .pt files;
    an image whose file lacks the statistics this run stores is embedded again."""
    from clip_assisted_data_labeling_b200 import imgstats
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.store import PackedStore
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    monkeypatch.setattr(imgstats, "image_stats", _fake_stats)
    root = str(tmp_path / "data")
    _write_images(root, 6)

    def run(**kw):
        enc = _FakeEncoder()
        enc.embed_dim = 8
        ds = Feature_Dataset(root, "ViT-L-14/openai", 4, shuffle_filenames=False, encoder=enc, packed_dir=str(tmp_path / "p"),
                             img_stats=True, **kw)
        return ds.process(), enc, PackedStore(str(tmp_path / "p"))

    (n, skipped), enc, st = run()
    assert (n, skipped) == (6, 0) and st.stat_names == imgstats.STAT_NAMES and len(st) == 6
    assert st.stats()[:, 0].tolist() == [40.0] * 6 and st.stats()[:, 1].tolist() == [30.0] * 6
    d = torch.load(os.path.join(root, "im002.pt"))["ViT-L-14/openai"]
    assert list(d.keys()) == imgstats.STAT_NAMES + CROP_NAMES and float(d["img_stat_width"]) == 40.0
    first = {p: (st.array()[i].copy(), st.stats()[i].copy()) for i, p in enumerate(st.paths)}
    # resume: nothing is embedded, the shard is complete again and holds the same rows
    (n, skipped), enc, st = run()
    assert (n, skipped, enc.calls) == (0, 6, 0) and sorted(st.paths) == sorted(first)
    for i, p in enumerate(st.paths):
        assert np.array_equal(st.array()[i], first[p][0]) and np.array_equal(st.stats()[i], first[p][1])
    # one file loses its statistics (as if written by a run with img_stats off): that image is embedded again
    pt = os.path.join(root, "im004.pt")
    d = torch.load(pt)
    d["ViT-L-14/openai"] = {k: v for k, v in d["ViT-L-14/openai"].items() if not k.startswith("img_stat_")}
    torch.save(d, pt)
    (n, skipped), enc, st = run()
    assert (n, skipped, enc.calls) == (1, 5, 1) and len(st) == 6
    assert list(torch.load(pt)["ViT-L-14/openai"].keys()) == imgstats.STAT_NAMES + CROP_NAMES


def test_lazy_paths_paths_at_and_sorted_flag(tmp_path):
    """The path list is not materialised on open: paths_at cuts the requested rows out of the raw sidecar bytes (several
    shards, multi-byte characters, carriage returns), .paths builds the same list on demand, and paths_sorted trusts the
    writer's flag for one shard, checks the shard edges for several, and falls back to the list for older shards."""
    import json
    from clip_assisted_data_labeling_b200.store import PackedStore, PackedWriter
    sd = str(tmp_path / "s")
    names = [["a/é\rx.jpg", "a/ü.jpg", "b/0.jpg"], ["b/1.jpg", "c.jpg"], ["d.jpg"]]
    for r, ns in enumerate(names):
        with PackedWriter(sd, "M/x", 4, shard=r) as w:
            w.append(np.full((len(ns), 4, 4), r, np.float32), ns)
    flat = [n for ns in names for n in ns]
    st = PackedStore(sd)
    assert st._paths is None and len(st) == 6
    assert st.paths_at([5, 0, 3, 1, 1]) == [flat[5], flat[0], flat[3], flat[1], flat[1]] and st._paths is None
    assert st.paths_at([]) == []
    for bad in ([6], [-1], [0, 99]):
        with pytest.raises(IndexError):
            st.paths_at(bad)
    assert st.paths_sorted() and st._paths is None      # from the flags and the shard edges alone
    assert st.paths == flat and st.paths_at([2, 4]) == [flat[2], flat[4]]
    # an unsorted shard is recorded as such
    with PackedWriter(sd, "M/x", 4, shard=1) as w:
        w.append(np.zeros((2, 4, 4), np.float32), ["z.jpg", "b/1.jpg"])
    assert json.load(open(os.path.join(sd, "shard-00001.json")))["sorted"] is False
    assert not PackedStore(sd).paths_sorted()
    # sorted inside every shard but not across them
    with PackedWriter(sd, "M/x", 4, shard=1) as w:
        w.append(np.zeros((2, 4, 4), np.float32), ["a/0.jpg", "a/1.jpg"])
    assert not PackedStore(sd).paths_sorted()
    # a shard written before the flag existed: decided from the list itself
    with PackedWriter(sd, "M/x", 4, shard=1) as w:
        w.append(np.zeros((2, 4, 4), np.float32), names[1])
    meta = json.load(open(os.path.join(sd, "shard-00001.json")))
    del meta["sorted"]
    json.dump(meta, open(os.path.join(sd, "shard-00001.json"), "w"))
    st = PackedStore(sd)
    assert st.paths_sorted() and st._paths is not None
