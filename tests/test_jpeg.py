"""K14 (JPEG decode ahead of K0).  The checker is Pillow itself — the library the reference decodes
with (`Image.open(path).convert('RGB')`, utils/embedder.py:167) and which is installed on both boxes — so every case is
compared bit for bit with `PIL.Image.open(...).convert('RGB')`.
  not gpu: the host stage (marker parse + Huffman decode of sequential and progressive streams through the C-ABI) feeding oracle/jpeg_oracle.py's numpy
           restatement of the device stage; refusal of streams outside the covered set; corrupt input.
  gpu:     the device stage (b2c_jpeg_reconstruct) on ragged batches, and the embedding driver end to end on .jpg files."""
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import jpeg_oracle
from oracle.preprocess_oracle import synthetic_image

CASES = [  # (W, H, subsampling, quality, extra save kwargs)
    (64, 64, 0, 90, {}), (512, 512, 2, 90, {}), (513, 511, 2, 75, {}), (200, 300, 1, 95, {}), (37, 911, 0, 50, {}),
    (640, 427, 2, 85, {"optimize": True}), (33, 17, 2, 100, {}), (16, 16, 1, 30, {}), (1000, 3, 2, 80, {}),
    (301, 200, "gray", 80, {}), (256, 256, 2, 90, {"restart_marker_blocks": 4}), (255, 257, 1, 60, {"restart_marker_rows": 1}),
    (8, 8, 0, 70, {}), (2, 2, 0, 70, {}), (9, 1, "gray", 70, {}), (1280, 960, 2, 5, {}),
    # progressive (SOF2): DC / AC first and refinement scans, end-of-band runs, non-interleaved AC scans on the
    # component's own block grid (odd sizes), restart intervals inside scans
    (512, 512, 2, 90, {"progressive": True}), (333, 201, 1, 75, {"progressive": True}), (97, 131, 0, 95, {"progressive": True}),
    (640, 427, 2, 40, {"progressive": True, "optimize": True}), (300, 300, "gray", 85, {"progressive": True}),
    (250, 250, 2, 85, {"progressive": True, "restart_marker_blocks": 3}), (17, 9, 2, 100, {"progressive": True}),
]


def make_jpeg(w, h, sub, q, kw, seed=0):
    rng = np.random.default_rng(seed + w * 7 + h)
    im = synthetic_image(w + h, h, w) if min(w, h) >= 100 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    pil = Image.fromarray(im)
    buf = io.BytesIO()
    if sub == "gray":
        pil.convert("L").save(buf, "JPEG", quality=q, **kw)
    else:
        pil.save(buf, "JPEG", quality=q, subsampling=sub, **kw)
    data = buf.getvalue()
    return data, np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}-s{c[2]}-q{c[3]}")
def test_host_huffman_stage_and_oracle_vs_pillow(lib, case):
    from clip_assisted_data_labeling_b200 import jpeg
    data, ref = make_jpeg(*case)
    info, coefs = jpeg.entropy_decode(data)
    assert (info.width, info.height) == (case[0], case[1]) and info.ncomp == (1 if case[2] == "gray" else 3)
    assert coefs.dtype == torch.int16 and coefs.numel() == info.coef_count
    got = jpeg_oracle.reconstruct(info.as_dict(), coefs.numpy())
    assert np.array_equal(got, ref)
    # the packed form the workers ship (per block: coefficients up to the last non-zero one in scan order) expands to
    # exactly the dense coefficients
    pinfo, packed = jpeg.entropy_decode_packed(data)
    assert packed.dtype == torch.uint8 and packed.numel() == pinfo.packed_bytes and packed.numel() % 16 == 0
    assert packed.numel() <= pinfo.packed_capacity and torch.equal(jpeg.expand_packed(pinfo, packed), coefs)


def test_streams_outside_the_covered_set_are_refused_not_misdecoded(lib):
    from clip_assisted_data_labeling_b200 import _lib, jpeg
    im = Image.fromarray(synthetic_image(1, 120, 160))
    buf = io.BytesIO()
    im.convert("CMYK").save(buf, "JPEG", quality=85)
    with pytest.raises(jpeg.UnsupportedJPEG):
        jpeg.entropy_decode(buf.getvalue())
    buf = io.BytesIO()
    im.save(buf, "JPEG", quality=85, subsampling=0)
    good = buf.getvalue()
    with pytest.raises(_lib.B2CError):
        jpeg.entropy_decode(b"\x89PNG\r\n" + good[6:])           # not a JPEG
    with pytest.raises(_lib.B2CError):
        jpeg.entropy_decode(good[:200])                           # truncated inside the headers
    # a header claiming a gigantic image is left to Pillow's decompression-bomb guard (no 13 GB coefficient buffer)
    bomb = bytearray(good)
    sof = bomb.index(b"\xff\xc0")
    bomb[sof + 5:sof + 9] = b"\xff\xff\xff\xff"
    with pytest.raises(jpeg.UnsupportedJPEG):
        jpeg.entropy_decode(bytes(bomb))
    # truncated entropy data: libjpeg pads with zeros and Pillow raises OSError ("image file is truncated"); this stage
    # refuses the stream instead of returning a partly grey image, so the file goes to Pillow, which has the last word
    for cut in (good[:len(good) // 2] + b"\xff\xd9", good[:len(good) // 2], good[:-40]):
        with pytest.raises(jpeg.UnsupportedJPEG):
            jpeg.entropy_decode(cut)
        with pytest.raises(jpeg.UnsupportedJPEG):
            jpeg.entropy_decode_packed(cut)
    # a complete stream followed by garbage is still a complete stream
    info, coefs = jpeg.entropy_decode(good + b"\x00" * 64)
    assert coefs.numel() == info.coef_count


def test_truncated_file_is_a_reported_failure_like_the_reference(lib, tmp_path):
    """Pillow raises OSError for a truncated JPEG (the reference then skips/substitutes the image, utils/embedder.py:176-181);
    the device-JPEG item path must not turn such a file into a partly grey image that gets embedded and saved."""
    from clip_assisted_data_labeling_b200.embedder import RawImageDataset
    for k, kw in enumerate(({}, {"progressive": True}, {"restart_marker_blocks": 4})):
        buf = io.BytesIO()
        Image.fromarray(synthetic_image(5 + k, 200, 240)).save(buf, "JPEG", quality=85, **kw)
        data = buf.getvalue()
        (tmp_path / f"t{k}.jpg").write_bytes(data[:len(data) * 2 // 3])
        with pytest.raises(OSError):
            Image.open(tmp_path / f"t{k}.jpg").convert("RGB")
        item, path = RawImageDataset([str(tmp_path / f"t{k}.jpg")], device_jpeg=True, device_huffman=False)[0]
        assert item is None and path.endswith(f"t{k}.jpg")
        # with the Huffman stage on the device the worker only parses the markers: a cut sequential file travels on as
        # file bytes and is reported by the device (test_device_path_reports_truncated_files_to_the_driver)
        item, _ = RawImageDataset([str(tmp_path / f"t{k}.jpg")], device_jpeg=True, device_huffman=True)[0]
        assert item is None if kw.get("progressive") else item[0] == "jpegf"


def test_host_stage_survives_corrupted_streams(lib):
    """600 random corruptions (byte flips, truncations, insertions, header edits) of baseline / progressive / restart
    streams: every call returns coefficients of the announced size or raises a clean error — the stage reads untrusted files."""
    from clip_assisted_data_labeling_b200 import _lib, jpeg
    rng = np.random.default_rng(7)
    seeds = [make_jpeg(*c)[0] for c in [(64, 48, 2, 80, {}), (97, 131, 0, 95, {"progressive": True}), (120, 80, 1, 60, {"restart_marker_blocks": 3}),
                                        (33, 17, "gray", 70, {}), (250, 250, 2, 85, {"progressive": True, "restart_marker_blocks": 3})]]
    outcomes = {"ok": 0, "refused": 0, "error": 0}
    for it in range(600):
        d = bytearray(seeds[it % len(seeds)])
        mode = it % 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 8))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        elif mode == 1:
            d = d[:int(rng.integers(2, len(d)))]
        elif mode == 2:
            p = int(rng.integers(0, len(d)))
            d[p:p] = bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
        else:
            d[int(rng.integers(2, min(len(d), 700)))] = int(rng.integers(0, 256))
        try:
            info, c = jpeg.entropy_decode(bytes(d))
            assert c.numel() == info.coef_count and info.width > 0 and info.height > 0
            outcomes["ok"] += 1
            if mode == 1:  # a stream cut inside its entropy data never "decodes"
                assert len(d) > len(seeds[it % len(seeds)]) - 3, (it, len(d))
        except jpeg.UnsupportedJPEG:
            outcomes["refused"] += 1
        except _lib.B2CError:
            outcomes["error"] += 1
    assert outcomes["ok"] > 60 and outcomes["error"] + outcomes["refused"] > 200, outcomes


def test_dataset_items_fall_back_to_pillow_per_file(lib, tmp_path):
    """RawImageDataset(device_jpeg=True): baseline / progressive .jpg -> coefficient item; CMYK .jpg and .png -> Pillow tensors."""
    from clip_assisted_data_labeling_b200.embedder import RawImageDataset
    im = Image.fromarray(synthetic_image(3, 90, 70))
    im.save(tmp_path / "a.jpg", quality=90)
    im.convert("CMYK").save(tmp_path / "b.jpg", quality=90)
    im.save(tmp_path / "c.png")
    (tmp_path / "d.jpg").write_bytes(b"")
    im.save(tmp_path / "e.jpg", format="PNG")  # mis-named file: Pillow identifies it by content, so must this path
    ds = RawImageDataset([str(tmp_path / n) for n in ("a.jpg", "b.jpg", "c.png", "d.jpg", "e.jpg")], device_jpeg=True)
    a, b, c, d, e = (ds[i][0] for i in range(5))
    assert isinstance(e, torch.Tensor) and tuple(e.shape) == (90, 70, 3)
    assert isinstance(a, tuple) and a[0] == "jpegf" and a[3].dtype == torch.uint8 and a[3].numel() % 16 == 0
    a2 = RawImageDataset([str(tmp_path / "a.jpg")], device_jpeg=True, device_huffman=False)[0][0]
    assert isinstance(a2, tuple) and a2[0] == "jpegp" and a2[2].dtype == torch.uint8
    assert isinstance(b, torch.Tensor) and np.array_equal(b.numpy(), np.asarray(Image.open(tmp_path / "b.jpg").convert("RGB")))
    assert isinstance(c, torch.Tensor) and tuple(c.shape) == (90, 70, 3)
    assert d is None


@pytest.mark.gpu
def test_device_reconstruct_ragged_batch_vs_pillow(lib):
    from clip_assisted_data_labeling_b200 import jpeg
    made = [make_jpeg(*c) for c in CASES]
    items = [jpeg.entropy_decode(d) for d, _ in made]
    outs = jpeg.reconstruct(items)                       # one batch: every sampling / size / restart variant together
    for (data, ref), got, case in zip(made, outs, CASES):
        assert np.array_equal(got.cpu().numpy(), ref), case
    pouts = jpeg.reconstruct_packed([jpeg.entropy_decode_packed(d) for d, _ in made])   # same from the packed form
    for (data, ref), got, case in zip(made, pouts, CASES):
        assert np.array_equal(got.cpu().numpy(), ref), case
    again = jpeg.reconstruct(items[3:5])                 # workspace / descriptor reuse with a different batch
    assert np.array_equal(again[0].cpu().numpy(), made[3][1]) and np.array_equal(again[1].cpu().numpy(), made[4][1])


@pytest.mark.gpu
def test_device_reconstruct_many_random_streams(lib):
    """Property run: 60 random sizes / qualities / samplings, noise and smooth content, all in two batches."""
    from clip_assisted_data_labeling_b200 import jpeg
    rng = np.random.default_rng(5)
    made = []
    for k in range(60):
        w, h = (int(v) for v in rng.integers(2, 400, 2))
        sub = [0, 1, 2, "gray"][k % 4]
        kw = {"restart_marker_blocks": int(rng.integers(1, 9))} if k % 5 == 0 else ({"optimize": True} if k % 7 == 0 else {})
        if k % 3 == 1:
            kw["progressive"] = True
        made.append(make_jpeg(w, h, sub, int(rng.integers(1, 101)), kw, seed=k))
    items, refs = [], []
    for data, ref in made:
        try:
            items.append(jpeg.entropy_decode_packed(data) if len(items) % 2 else jpeg.entropy_decode(data))
            refs.append(ref)
        except jpeg.UnsupportedJPEG:  # subsampled images narrower than two chroma samples stay on Pillow
            assert ref.shape[1] <= 2
    assert len(items) >= 55
    for kind, fn in ((torch.uint8, jpeg.reconstruct_packed), (torch.int16, jpeg.reconstruct)):
        sel = [i for i, it in enumerate(items) if it[1].dtype == kind]
        for lo in range(0, len(sel), 17):
            part = sel[lo:lo + 17]
            for i, got in zip(part, fn([items[i] for i in part])):
                assert np.array_equal(got.cpu().numpy(), refs[i])


def test_worker_collate_packs_a_batch_into_one_shared_segment(lib, tmp_path):
    """DataLoader workers (2 processes) + collate_raw: the file bytes / packed coefficients of a batch arrive in the main
    process as views of ONE shared-memory tensor per kind, with the right contents."""
    from torch.utils.data import DataLoader
    from clip_assisted_data_labeling_b200 import jpeg
    from clip_assisted_data_labeling_b200.embedder import RawImageDataset, collate_raw
    paths = []
    for k in range(6):
        p = tmp_path / f"{k}.jpg"
        Image.fromarray(synthetic_image(k, 80 + 8 * k, 120)).save(p, quality=90, progressive=(k % 3 == 2))
        paths.append(str(p))
    dl = DataLoader(RawImageDataset(paths, device_jpeg=True, device_huffman=True), batch_size=6, shuffle=False, num_workers=2,
                    collate_fn=collate_raw)
    items, got_paths = next(iter(dl))
    assert got_paths == paths and [it[0] for it in items] == ["jpegf", "jpegf", "jpegp"] * 2
    for kind in ("jpegf", "jpegp"):
        ts = [it[-1] for it in items if it[0] == kind]
        assert all(t.is_shared() for t in ts) and len({t.untyped_storage().data_ptr() for t in ts}) == 1
    for it, p in zip(items, paths):
        data = open(p, "rb").read()
        if it[0] == "jpegf":
            assert bytes(it[3][:len(data)].numpy()) == data and it[3].numel() % 16 == 0
        else:
            _, want = jpeg.entropy_decode_packed(data)
            assert torch.equal(it[2], want)


SEQ_CASES = [c for c in CASES if not c[4].get("progressive")]


def test_huff_prepare_accepts_sequential_and_refuses_progressive(lib):
    """Host half of K14b: the marker parse that feeds the device Huffman stage (no GPU needed)."""
    from clip_assisted_data_labeling_b200 import jpeg
    for case in CASES:
        data, _ = make_jpeg(*case)
        if case[4].get("progressive"):
            with pytest.raises(jpeg.UnsupportedJPEG):
                jpeg.huff_prepare(data)
            continue
        info, huff = jpeg.huff_prepare(data)
        ref_info, _ = jpeg.entropy_decode(data)
        assert bytes(info)[:64] == bytes(ref_info)[:64] and info.coef_count == ref_info.coef_count
        assert [list(q) for q in info.qt] == [list(q) for q in ref_info.qt]
        assert 0 < huff.scan_begin < len(data) and huff.scan_begin + huff.scan_bytes == len(data)
        assert data[huff.scan_begin - 3:huff.scan_begin - 1] == b"\x00\x3f"  # SOS tail: Ss = 0, Se = 63
        assert huff.restart_interval == ref_info.restart_interval
        assert any(huff.tab[0].look) and any(huff.tab[1].look)


@pytest.mark.gpu
def test_device_huffman_stage_equals_host_stage(lib):
    """K14b: coefficients decoded on the device (destuff + self-synchronising parallel Huffman decode + DC prefix sums) are
    bit-identical to the host stage's on every sequential case of the corpus (all samplings, grey, restart intervals,
    optimised tables, 2x2 .. 1280x960), in ONE batch; the images are Pillow's."""
    from clip_assisted_data_labeling_b200 import jpeg
    made = [make_jpeg(*c) for c in SEQ_CASES]
    items = [jpeg.prepare_file(d) for d, _ in made]
    coefs, status = jpeg.huffman_device(items)
    assert status.cpu().tolist() == [0] * len(items)
    off = 0
    for (data, _), it, case in zip(made, items, SEQ_CASES):
        _, want = jpeg.entropy_decode(data)
        got = coefs[off:off + want.numel()].cpu()
        off += want.numel()
        assert torch.equal(got, want), case
    outs, st = jpeg.decode_device(items)
    for (data, ref), got, case in zip(made, outs, SEQ_CASES):
        assert np.array_equal(got.cpu().numpy(), ref), case


@pytest.mark.gpu
def test_device_huffman_stage_many_random_streams_and_large_files(lib):
    """Property run: random sizes / qualities / samplings / restart intervals (including one MCU), noise content (long
    codes, many sub-sequences per image: several rounds per CTA) and smooth content (short streams)."""
    from clip_assisted_data_labeling_b200 import jpeg
    rng = np.random.default_rng(11)
    made = []
    for k in range(48):
        w, h = (int(v) for v in rng.integers(2, 500, 2))
        if k % 12 == 0:
            w, h = 1536, 1024 + k  # > 512 sub-sequences: more than one round
        sub = [0, 1, 2, "gray"][k % 4]
        kw = {}
        if k % 3 == 0:
            kw["restart_marker_blocks"] = int(rng.integers(1, 12))
        elif k % 3 == 1 and k % 2:
            kw["restart_marker_rows"] = 1
        if k % 5 == 0:
            kw["optimize"] = True
        rngk = np.random.default_rng(100 + k)
        if k % 2:
            im = rngk.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            im = synthetic_image(k, h, w)
        buf = io.BytesIO()
        pil = Image.fromarray(im)
        q = int(rng.integers(1, 101))
        (pil.convert("L") if sub == "gray" else pil).save(buf, "JPEG", quality=q, **({} if sub == "gray" else {"subsampling": sub}), **kw)
        made.append(buf.getvalue())
    items, wants = [], []
    for d in made:
        try:
            it = jpeg.prepare_file(d)
        except jpeg.UnsupportedJPEG:
            continue
        items.append(it)
        wants.append(jpeg.entropy_decode(d)[1])
    assert len(items) >= 44
    for lo in range(0, len(items), 19):
        part = items[lo:lo + 19]
        coefs, status = jpeg.huffman_device(part)
        assert status.cpu().tolist() == [0] * len(part)
        off = 0
        for j, it in enumerate(part):
            n = wants[lo + j].numel()
            assert torch.equal(coefs[off:off + n].cpu(), wants[lo + j]), (lo + j, it[0].width, it[0].height, it[1].restart_interval)
            off += n


@pytest.mark.gpu
def test_device_huffman_stage_reports_damaged_and_truncated_streams(lib):
    """A stream the host stage refuses must not come back from the device as status 0 with different coefficients: cut
    files, flipped bytes.  Either the device reports it (status != 0) or its coefficients equal the host stage's."""
    from clip_assisted_data_labeling_b200 import _lib, jpeg
    rng = np.random.default_rng(3)
    seeds = [make_jpeg(*c)[0] for c in [(200, 300, 1, 95, {}), (256, 256, 2, 90, {"restart_marker_blocks": 4}), (301, 200, "gray", 80, {}),
                                        (512, 512, 2, 90, {})]]
    datas = []
    for it in range(120):
        d = bytearray(seeds[it % len(seeds)])
        info, huff = jpeg.huff_prepare(bytes(d))
        lo = int(huff.scan_begin)
        if it % 3 == 0:
            d = d[:int(rng.integers(lo + 1, len(d)))]
        elif it % 3 == 1:
            for _ in range(int(rng.integers(1, 4))):
                d[int(rng.integers(lo, len(d)))] = int(rng.integers(0, 256))
        else:
            d = d[:len(d) // 2] + b"\xff\xd9"
        datas.append(bytes(d))
    items, host = [], []
    for d in datas:
        try:
            items.append(jpeg.prepare_file(d))
        except (jpeg.UnsupportedJPEG, _lib.B2CError):
            continue
        try:
            host.append(jpeg.entropy_decode(d)[1])
        except (jpeg.UnsupportedJPEG, _lib.B2CError):
            host.append(None)
    coefs, status = jpeg.huffman_device(items)
    st = status.cpu().tolist()
    off, reported, agreed = 0, 0, 0
    for code, it, want in zip(st, items, host):
        n = int(it[0].coef_count)
        if code == 0:
            assert want is not None and torch.equal(coefs[off:off + n].cpu(), want)
            agreed += 1
        else:
            reported += 1
        off += n
    assert reported > 60 and agreed >= 0


@pytest.mark.gpu
def test_device_path_reports_truncated_files_to_the_driver(lib, tmp_path):
    """Files cut inside their entropy data: Pillow raises OSError (the reference skips them, utils/embedder.py:176-181).
    With the Huffman stage on the device they are found there, retried on the host, and end up in Feature_Dataset.failed —
    never embedded from a partly grey image.  Intact files of the same batch are unaffected."""
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from clip_assisted_data_labeling_b200.embedder import RawImageDataset, collate_raw, to_device_images
    from oracle import vit_oracle
    paths = []
    for k in range(5):
        buf = io.BytesIO()
        Image.fromarray(synthetic_image(20 + k, 160, 200)).save(buf, "JPEG", quality=85, **({"restart_marker_blocks": 4} if k == 3 else {}))
        data = buf.getvalue()
        (tmp_path / f"f{k}.jpg").write_bytes(data if k % 2 == 0 else data[:len(data) * 2 // 3])
        paths.append(str(tmp_path / f"f{k}.jpg"))
    ds = RawImageDataset(paths, device_jpeg=True, device_huffman=True)
    items, _ = collate_raw([ds[i] for i in range(5)])
    assert all(it[0] == "jpegf" for it in items)
    outs = to_device_images(items, "cuda")
    assert [o is None for o in outs] == [False, True, False, True, False]
    for k in (0, 2, 4):
        assert np.array_equal(outs[k].cpu().numpy(), np.asarray(Image.open(paths[k]).convert("RGB")))
    m = vit_oracle.build_visual("ViT-B-32", "openai", seed=0)
    fd = Feature_Dataset(str(tmp_path), "ViT-B-32/openai", batch_size=8, shuffle_filenames=False,
                         state_dict=vit_oracle.visual_state_dict(m), device_jpeg=True)
    n, _ = fd.process()
    assert n == 3 and sorted(os.path.basename(p) for p in fd.failed) == ["f1.jpg", "f3.jpg"]
    assert not (tmp_path / "f1.pt").exists() and (tmp_path / "f0.pt").exists()


@pytest.mark.gpu
def test_driver_embeds_jpg_files_identically_to_the_pillow_path(lib, tmp_path):
    """Feature_Dataset on .jpg files: device_jpeg on vs off write the same .pt contents (same decoded pixels -> same bits)."""
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from oracle import vit_oracle
    m = vit_oracle.build_visual("ViT-B-32", "openai", seed=0)
    sd = vit_oracle.visual_state_dict(m)
    res = {}
    for mode in (True, False):
        root = tmp_path / f"ds{int(mode)}"
        root.mkdir()
        for k in range(6):
            Image.fromarray(synthetic_image(k, 150 + 20 * k, 210)).save(root / f"{k:02d}.jpg", quality=88, subsampling=[2, 1, 0][k % 3],
                                                                       progressive=(k == 5))
        ds = Feature_Dataset(str(root), "ViT-B-32/openai", batch_size=4, shuffle_filenames=False, state_dict=sd, device_jpeg=mode)
        n, _ = ds.process()
        assert n == 6 and not ds.failed
        res[mode] = [torch.load(root / f"{k:02d}.pt")["ViT-B-32/openai"] for k in range(6)]
    for a, b in zip(res[True], res[False]):
        assert list(a.keys()) == list(b.keys())
        for key in a:
            assert torch.equal(a[key], b[key]), key


@pytest.mark.gpu
def test_driver_handles_camera_sized_files(lib, tmp_path):
    """24 MP photographs next to small images in one batch: the device Huffman stage runs many rounds per CTA, the
    reconstruction and the statistics stream the planes, K0 switches to one colour channel per CTA — and the .pt files
    equal the Pillow path's bit for bit."""
    from clip_assisted_data_labeling_b200.embed_driver import Feature_Dataset
    from oracle import vit_oracle
    sd = vit_oracle.visual_state_dict(vit_oracle.build_visual("ViT-B-32", "openai", seed=0))
    rng = np.random.default_rng(9)
    specs = [(6000, 4000, {}), (4000, 6000, {"progressive": True}), (512, 512, {}), (3000, 2000, {"restart_marker_rows": 1}), (97, 61, {})]
    res = {}
    for mode in (True, False):
        root = tmp_path / f"big{int(mode)}"
        root.mkdir()
        for k, (w, h, kw) in enumerate(specs):
            small = np.repeat(np.repeat(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8), 8, 0), 8, 1)[:h, :w]
            im = np.clip(small.astype(np.int16) + np.random.default_rng(k).integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
            Image.fromarray(im).save(root / f"{k}.jpg", quality=85, **kw)
        rng = np.random.default_rng(9)  # the same files for both modes
        ds = Feature_Dataset(str(root), "ViT-B-32/openai", batch_size=8, shuffle_filenames=False, state_dict=sd, device_jpeg=mode)
        n, _ = ds.process()
        assert n == len(specs) and not ds.failed
        res[mode] = [torch.load(root / f"{k}.pt")["ViT-B-32/openai"] for k in range(len(specs))]
    for a, b in zip(res[True], res[False]):
        assert list(a.keys()) == list(b.keys())
        for key in a:
            assert torch.equal(a[key], b[key]), key
