"""CPU-side checks: the C-ABI library loads and exports every symbol include/b2c.h declares, the host-only
entry points agree with the oracle, argument errors are reported without touching a GPU, and the host logic
(sharding, band partition, .pt layout, resume) behaves like the reference's drivers."""
import ctypes as C
import os
import re
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle.preprocess_oracle import crop_geometry


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b2c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from clip_assisted_data_labeling_b200 import _lib
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2c.h but not exported by libb2c.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes prototypes out of sync with include/b2c.h"
    assert lib.b2c_version() >= 100


def test_struct_layouts():
    from clip_assisted_data_labeling_b200 import _lib
    assert C.sizeof(_lib.Crop) == 32 and C.sizeof(_lib.Pair) == 12 and C.sizeof(_lib.VitCfg) == 32


def test_crop_geometry_matches_oracle(lib, golden):
    from clip_assisted_data_labeling_b200 import _lib
    rng = np.random.default_rng(0)
    sizes = [(512, 512), (768, 512), (100, 1000), (513, 512), (333, 517), (1, 50), (7, 3), (2, 2), (4000, 3000)]
    sizes += [tuple(int(v) for v in rng.integers(1, 3000, 2)) for _ in range(400)]
    sizes += [(r[0], r[1]) for r in golden("crop_geometry_ref.npz")["rows"].tolist()]
    for R in (224, 336):
        for (W, H) in sizes:
            g = (_lib.Crop * 4)()
            assert lib.b2c_crop_geometry(W, H, R, g) == 0
            want = crop_geometry(W, H, R)
            got = [dict(cw=c.cw, ch=c.ch, dx=c.dx, dy=c.dy, out_w=c.out_w, out_h=c.out_h, off_x=c.off_x, off_y=c.off_y) for c in g]
            assert got == want, (W, H, R)


def test_argument_errors_do_not_need_a_gpu(lib):
    from clip_assisted_data_labeling_b200 import _lib
    assert lib.b2c_crop_geometry(0, 5, 224, (_lib.Crop * 4)()) == -1
    assert b"positive" in lib.b2c_last_error()
    assert lib.b2c_gemm_bf16(None, None, None, None, 128, 256, 64, 0, None) == -1
    assert lib.b2c_dedup_pairs(None, 10, 768, 0, 10, C.c_float(0.96), 1, None, 0, None, None) == -1
    cfg = _lib.VitCfg(224, 14, 1000, 24, 16, 4096, 768, 0)  # width not a multiple of 256
    h = C.c_void_p()
    assert lib.b2c_vit_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _lib.VitCfg(224, 32, 768, 12, 12, 3072, 512, 0)
    assert lib.b2c_vit_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.b2c_vit_ready(h) == -3 and b"has not been set" in lib.b2c_last_error()
    need = C.c_size_t()
    assert lib.b2c_vit_workspace_bytes(h, 32, C.byref(need)) == 0 and need.value > 32 * 50 * 768 * 4
    assert lib.b2c_vit_destroy(h) == 0
    ws = C.c_size_t()
    assert lib.b2c_preprocess_workspace_bytes(8, 512, 224, C.byref(ws)) == 0 and ws.value > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from clip_assisted_data_labeling_b200 import _lib
    from clip_assisted_data_labeling_b200.dedup import duplicate_pairs
    from clip_assisted_data_labeling_b200.embedder import CLIP_Encoder
    with pytest.raises(_lib.B2CError):
        duplicate_pairs(torch.randn(8, 64), 0.9)
    with pytest.raises((RuntimeError, _lib.B2CError)):
        CLIP_Encoder("ViT-B-32/openai", device="cpu")
    with pytest.raises((RuntimeError, _lib.B2CError)):
        CLIP_Encoder("ViT-B-32/openai")


def test_owned_blocks_cover_every_pair_once_and_balance():
    """Multi-GPU work split (dedup.owned_blocks): own-shard block first (runs under the all-gather), then bands right of
    the shards' diagonal blocks; every pair (i < j) exactly once, work balanced at BASELINE's 1 M x 8."""
    import numpy as np
    from clip_assisted_data_labeling_b200.dedup import owned_blocks
    for n_local, world, br in [(10, 1, 4), (10, 2, 4), (7, 3, 2), (64, 4, 16), (300, 8, 128)]:
        n = n_local * world
        cov = np.zeros((n, n), np.int32)
        for r in range(world):
            local, rest = owned_blocks(n_local, r, world, br)
            assert all(b[0] >= r * n_local and b[1] <= (r + 1) * n_local and b[2:] == b[:2] for b in local)  # no peer data
            for (r0, r1, c0, c1) in local + rest:
                for i in range(r0, r1):
                    lo = max(i + 1, c0)
                    if lo < c1:
                        cov[i, lo:c1] += 1
        iu = np.triu_indices(n, 1)
        assert (cov[iu] == 1).all() and cov.sum() == len(iu[0]), (n_local, world)
    for n_local, world in [(125_000, 8), (250_000, 4), (500_000, 2)]:
        work = []
        for r in range(world):
            local, rest = owned_blocks(n_local, r, world)
            work.append(sum((b[1] - b[0]) * (b[3] - b[2]) for b in rest) + sum((b[1] - b[0]) * (b[1] - b[0] - 1) / 2 for b in local))
        assert max(work) / min(work) < 1.01, (world, work)


def test_band_tickets_generations_and_static_fallback():
    """dedup.BandTickets: every band index of a generation is handed out exactly once across the ranks, each rank stops at
    its first ticket past the end, and the next generation (an overflow re-run) needs no reset; without a store the ranks
    take the static round-robin share."""
    from clip_assisted_data_labeling_b200 import dedup

    class FakeStore:
        def __init__(self):
            self.v = {}

        def add(self, key, n):
            self.v[key] = self.v.get(key, 0) + n
            return self.v[key]

    n, world = 11, 3
    shared = FakeStore()
    ranks = []
    for r in range(world):
        t = dedup.BandTickets(n, r, world)
        t.store, t.key = shared, "k"      # one job-wide counter
        ranks.append(t)
    for gen in range(3):
        for t in ranks:
            t.new_generation()
        got, live = [], list(range(world))
        while live:                        # ranks draw in an arbitrary interleaving; a rank stops at its first -1
            for r in list(live):
                i = ranks[r].next()
                if i < 0:
                    live.remove(r)
                else:
                    got.append(i)
        assert sorted(got) == list(range(n)), (gen, got)
    solo = [dedup.BandTickets(n, r, world) for r in range(world)]
    got = []
    for t in solo:
        t.store = None
        t.new_generation()
        while (i := t.next()) >= 0:
            got.append(i)
    assert sorted(got) == list(range(n))


def test_owned_bands_partition():
    from clip_assisted_data_labeling_b200.dedup import BAND_ROWS, owned_bands
    for n in (1, 2047, 2048, 2049, 100_000, 1_000_003):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += owned_bands(n, r, world)
            seen.sort()
            assert seen[0][0] == 0 and seen[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
            assert all(0 < e - b <= BAND_ROWS for b, e in seen)
    # cyclic dealing balances the triangular work within a few percent at 1M x 8
    n, world = 1_000_000, 8
    work = [sum((e - b) * (n - (b + e) / 2) for b, e in owned_bands(n, r, world)) for r in range(world)]
    assert max(work) / min(work) < 1.02


def test_sort_pairs_row_major():
    from clip_assisted_data_labeling_b200.dedup import sort_pairs
    p = np.array([[5, 9], [0, 3], [5, 6], [0, 2]])
    s = np.array([0.1, 0.2, 0.3, 0.4], np.float32)
    sp, ss = sort_pairs(p, s)
    assert sp.tolist() == [[0, 2], [0, 3], [5, 6], [5, 9]] and ss.tolist() == pytest.approx([0.4, 0.2, 0.3, 0.1])
    sp, ss = sort_pairs(np.zeros((0, 2)), np.zeros(0))
    assert sp.shape == (0, 2)


def test_shard_for_rank_covers_everything():
    from clip_assisted_data_labeling_b200.embed_driver import shard_for_rank
    paths = [f"{i:05d}.png" for i in range(1003)]
    for world in (1, 2, 4, 8):
        parts = [shard_for_rank(paths, r, world) for r in range(world)]
        assert sum(parts, []) == paths
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


class _FakeEncoder:
    """Host-logic stand-in (no CUDA): embeds an image as its mean colour so files can be checked."""
    device = "cpu"
    img_resolution = 224

    def __init__(self):
        self.calls = 0

    def encode_images_u8(self, images):
        self.calls += 1
        out = torch.zeros(len(images), 4, 8)
        for i, im in enumerate(images):
            out[i] = im.float().mean() / 255.0 + torch.arange(4)[:, None]
        return out


def _write_images(root, n, size=(40, 30)):
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    rng = np.random.default_rng(0)
    for i in range(n):
        Image.fromarray(rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)).save(os.path.join(root, f"im{i:03d}.png"))


def test_writer_processes_write_the_same_files_as_writer_threads(tmp_path, lib):
    """Large jobs hand whole batches to spawned writer processes (pickling is GIL-bound); same bytes in the tensors,
    same keys, and every value of a file is a view of one storage."""
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    out = {}
    for procs in (0, 2):
        root = str(tmp_path / f"data{procs}")
        _write_images(root, 9)
        n, _ = Feature_Dataset(root, "ViT-L-14/openai", batch_size=4, shuffle_filenames=False, encoder=_FakeEncoder(),
                               writer_procs=procs).process()
        assert n == 9
        out[procs] = [torch.load(os.path.join(root, f"im{i:03d}.pt"))["ViT-L-14/openai"] for i in range(9)]
    for a, b in zip(out[0], out[2]):
        assert list(a.keys()) == list(b.keys()) == CROP_NAMES
        assert all(torch.equal(a[k], b[k]) and tuple(a[k].shape) == (1, 8) and a[k].dtype == torch.float32 for k in a)
        assert len({t.untyped_storage().data_ptr() for t in b.values()}) == 1


def test_feature_dataset_layout_resume_and_consumers(tmp_path, lib):
    """`.pt` layout, merge across models, per-image resume, and the reference consumers'
    read paths (_2:30-38 first-key default + squeeze; _5:77-82 cat over crop_names)."""
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.vit_arch import CROP_NAMES
    root = str(tmp_path / "data")
    _write_images(root, 11)
    open(os.path.join(root, "broken.jpg"), "wb").write(b"not an image")
    enc = _FakeEncoder()
    ds = Feature_Dataset(root, "ViT-L-14/openai", batch_size=4, shuffle_filenames=False, encoder=enc)
    assert len(ds) == 12
    n_emb, n_skip = ds.process()
    assert (n_emb, n_skip) == (11, 0) and len(ds.failed) == 1 and enc.calls == 3
    d = torch.load(os.path.join(root, "im003.pt"))
    assert list(d.keys()) == ["ViT-L-14/openai"]
    assert list(d["ViT-L-14/openai"].keys()) == CROP_NAMES
    for i, c in enumerate(CROP_NAMES):
        t = d["ViT-L-14/openai"][c]
        assert t.dtype == torch.float32 and tuple(t.shape) == (1, 8) and not t.is_cuda
        assert abs(float(t[0, 0]) - float(t[0, 0] // 1) - float(d["ViT-L-14/openai"]["centre_crop"][0, 0])) < 1e-6 or i == 0
    # resume: nothing is re-encoded; a second model merges into the same files
    enc2 = _FakeEncoder()
    assert Feature_Dataset(root, "ViT-L-14/openai", 4, shuffle_filenames=False, encoder=enc2).process() == (0, 11)
    assert enc2.calls == 0
    Feature_Dataset(root, "ViT-B-32/openai", 4, shuffle_filenames=False, encoder=_FakeEncoder()).process()
    d = torch.load(os.path.join(root, "im003.pt"))
    assert list(d.keys()) == ["ViT-L-14/openai", "ViT-B-32/openai"]
    # force_reencode rewrites only this model's entry
    Feature_Dataset(root, "ViT-B-32/openai", 4, force_reencode=True, shuffle_filenames=False, encoder=_FakeEncoder()).process()
    assert list(torch.load(os.path.join(root, "im003.pt")).keys()) == ["ViT-B-32/openai"]
    # _5_predict_labels.py:77-82 feature assembly
    feats = torch.cat([d["ViT-L-14/openai"][c] for c in ["centre_crop", "subcrop2"]], dim=0).flatten()
    assert feats.shape == (16,)
    # rank sharding: two ranks together cover the directory
    a = Feature_Dataset(root, "ViT-L-14/openai", 4, encoder=_FakeEncoder(), rank=0, world_size=2)
    b = Feature_Dataset(root, "ViT-L-14/openai", 4, encoder=_FakeEncoder(), rank=1, world_size=2)
    assert sorted(a.img_filepaths + b.img_filepaths) == sorted(ds.img_filepaths)


def test_get_paths_and_embeddings_matches_reference_semantics(tmp_path):
    """Needs .jpg + .pt, first-key default, fp16 cast, chunking, unreadable samples skipped (_2:8-49)."""
    from clip_assisted_data_labeling_b200.dedup import get_paths_and_embeddings
    root = tmp_path / "d"
    root.mkdir()
    for i in range(7):
        (root / f"{i}.jpg").write_bytes(b"")
        torch.save({"M/x": {"square_padded_crop": torch.full((1, 4), float(i))}}, root / f"{i}.pt")
    (root / "only_pt.pt").write_bytes(b"junk")
    (root / "png_only.png").write_bytes(b"")
    torch.save({"M/x": {}}, root / "png_only.pt")
    (root / "bad.jpg").write_bytes(b"")
    (root / "bad.pt").write_bytes(b"junk")
    args = types.SimpleNamespace(root_dir=str(root), clip_model_to_use=None, chunk_size=3)
    chunks = list(get_paths_and_embeddings(args, "square_padded_crop"))
    assert args.clip_model_to_use == "M/x"
    assert [len(p) for p, _ in chunks] == [3, 3, 1]
    for paths, embs in chunks:
        for p, e in zip(paths, embs):
            assert p.endswith(".jpg") and e.dtype == torch.float16 and e.shape == (4,)
            assert float(e[0]) == float(os.path.basename(p)[:-4])


def test_scorer_module_twin_loads_reference_pickles(tmp_path):
    from clip_assisted_data_labeling_b200.scorer import SimpleFC, load_regressor
    m = SimpleFC(12, [8, 4], 1, clip_models=["A/b"], crop_names=["centre_crop"], dropout_prob=0.5)
    import sys
    mod = types.ModuleType("utils.nn_model")
    pkg = types.ModuleType("utils")
    pkg.__path__ = []
    mod.SimpleFC = SimpleFC
    SimpleFC.__module__ = "utils.nn_model"
    sys.modules["utils"], sys.modules["utils.nn_model"] = pkg, mod
    try:
        torch.save(m, tmp_path / "m.pth")  # pickled under the reference's class path, like _4_train_model.py:237
    finally:
        SimpleFC.__module__ = "clip_assisted_data_labeling_b200.scorer"
        del sys.modules["utils"], sys.modules["utils.nn_model"]
    back = load_regressor(str(tmp_path / "m.pth"))
    x = torch.randn(5, 12)
    assert torch.equal(back(x), m.eval()(x)) and back.crop_names == ["centre_crop"]


def test_bench_stdout_carries_exactly_one_json_line():
    """bench.py claims fd 1 for its JSON line: Python prints and C-level writes to stdout during the run (NCCL's version
    banner at N > 1) end up on stderr."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, bench; fd = bench._claim_stdout(); print('python noise'); os.system('echo c-level noise'); "
            "bench._emit(fd, {'value': 1.5})")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("\n") == 1 and json.loads(r.stdout) == {"value": 1.5}
    assert "python noise" in r.stderr and "c-level noise" in r.stderr



def test_bands_right_of_diagonal_is_what_the_ranks_draw_from():
    """dedup.bands_right_of_diagonal (the list duplicate_pairs_distributed hands out through the shared counter): together
    with every rank's own-shard block it covers each pair (i < j) exactly once, it is the union of the static deal, and it
    is ordered largest band first."""
    import numpy as np
    from clip_assisted_data_labeling_b200.dedup import bands_right_of_diagonal, owned_blocks
    for n_local, world, br in [(10, 1, 4), (10, 2, 4), (7, 3, 2), (64, 4, 16), (300, 8, 128)]:
        n = n_local * world
        bands = bands_right_of_diagonal(n_local, world, br)
        cov = np.zeros((n, n), np.int32)
        for r in range(world):
            lo = r * n_local
            for i in range(lo, lo + n_local):
                cov[i, i + 1:lo + n_local] += 1
        for (r0, r1, c0, c1) in bands:
            assert c0 >= r1  # entirely right of the diagonal block of its shard
            cov[r0:r1, c0:c1] += 1
        iu = np.triu_indices(n, 1)
        assert (cov[iu] == 1).all() and cov.sum() == len(iu[0])
        sizes = [(b[1] - b[0]) * (b[3] - b[2]) for b in bands]
        assert sizes == sorted(sizes, reverse=True)
        assert sorted(bands) == sorted(b for r in range(world) for b in owned_blocks(n_local, r, world, br)[1])
    assert bands_right_of_diagonal(1000, 1) == []
