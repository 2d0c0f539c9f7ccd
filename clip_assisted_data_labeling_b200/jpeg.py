"""K14 host side: JPEG files -> uint8 [H,W,3] device tensors, bit-exact with `Image.open(path).convert('RGB')`
(utils/embedder.py:167) for the streams include/b2c.h's b2c_jpeg_* covers.

`entropy_decode` (marker parse + Huffman decode, pure host work, no CUDA call, releases the GIL inside ctypes) is what
DataLoader workers run instead of Pillow's full decode; `reconstruct` (dequantise + inverse DCT + chroma upsampling +
colour conversion) is one batched device call on the main process.  A stream the device path does not cover raises
`UnsupportedJPEG`; the driver then keeps that file on the Pillow path."""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

ERR_UNSUPPORTED = -5
MAX_PIXELS = 89_478_485  # Pillow's Image.MAX_IMAGE_PIXELS: above it Pillow warns, above twice it refuses


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("hs", C.c_int32 * 3),
                ("vs", C.c_int32 * 3), ("mcus_x", C.c_int32), ("mcus_y", C.c_int32), ("blocks_w", C.c_int32 * 3),
                ("blocks_h", C.c_int32 * 3), ("comp_w", C.c_int32 * 3), ("comp_h", C.c_int32 * 3),
                ("restart_interval", C.c_int32), ("adobe_transform0", C.c_int32), ("progressive", C.c_int32),
                ("reserved_", C.c_int32), ("coef_offset", C.c_int64 * 3),
                ("coef_count", C.c_int64), ("nblocks", C.c_int32), ("reserved2_", C.c_int32), ("offs_off", C.c_int64),
                ("counts_off", C.c_int64), ("vals_off", C.c_int64), ("packed_capacity", C.c_int64), ("packed_bytes", C.c_int64),
                ("qt", (C.c_uint16 * 64) * 3)]

    def as_dict(self) -> dict:
        return {"width": self.width, "height": self.height, "ncomp": self.ncomp, "hs": list(self.hs), "vs": list(self.vs),
                "blocks_w": list(self.blocks_w), "blocks_h": list(self.blocks_h), "comp_w": list(self.comp_w),
                "comp_h": list(self.comp_h), "coef_offset": list(self.coef_offset), "coef_count": self.coef_count,
                "qt": [list(q) for q in self.qt], "restart_interval": self.restart_interval}


class HuffTab(C.Structure):
    _fields_ = [("look", C.c_uint16 * 512), ("maxcode", C.c_int32 * 18), ("valoffset", C.c_int32 * 18), ("vals", C.c_uint8 * 256)]


class JpegHuff(C.Structure):
    """include/b2c.h b2c_jpeg_huff: where the entropy-coded segment sits in the file + its four Huffman tables."""
    _fields_ = [("scan_begin", C.c_int64), ("scan_bytes", C.c_int64), ("restart_interval", C.c_int32), ("reserved_", C.c_int32),
                ("tab", HuffTab * 4)]


HUFF_CORRUPT, HUFF_TRUNCATED, HUFF_NOSYNC = 1, 2, 4


class UnsupportedJPEG(ValueError):
    """The stream is valid but outside what the device path decodes (progressive, CMYK, ...)."""


def _raise(rc: int, what: str):
    msg = _lib.load().b2c_last_error().decode("utf-8", "replace")
    if rc == ERR_UNSUPPORTED:
        raise UnsupportedJPEG(msg)
    raise _lib.B2CError(f"{what} failed (code {rc}): {msg}")


_scratch = threading.local()


def huff_prepare(data: bytes) -> Tuple[JpegInfo, JpegHuff]:
    """Marker parse only (host, microseconds): what the device Huffman stage (`decode_device`) needs besides the file's
    bytes.  Raises UnsupportedJPEG for progressive / multi-scan files — those keep `entropy_decode_packed`."""
    lib = _lib.load()
    info, huff = JpegInfo(), JpegHuff()
    rc = lib.b2c_jpeg_huff_prepare(data, len(data), C.byref(info), C.byref(huff))  # bytes are passed by pointer, no copy
    if rc != 0:
        _raise(rc, "b2c_jpeg_huff_prepare")
    if info.width * info.height > MAX_PIXELS:
        raise UnsupportedJPEG(f"{info.width}x{info.height} pixels exceeds the device path's limit of {MAX_PIXELS}")
    return info, huff


def file_tensor(data: bytes) -> torch.Tensor:
    """The file's bytes as a uint8 host tensor padded to a multiple of 16 bytes (files of a batch are laid back to back in
    one pinned buffer and each must start on a 16-byte boundary)."""
    n = len(data)
    t = torch.empty((n + 15) & ~15, dtype=torch.uint8)
    C.memmove(t.data_ptr(), data, n)
    if t.numel() > n:
        t[n:] = 0
    return t


def prepare_file(data: bytes) -> Tuple[JpegInfo, JpegHuff, torch.Tensor]:
    """What a DataLoader worker hands to the main process for a file the device decodes entirely: (info, huff, bytes)."""
    info, huff = huff_prepare(data)
    return info, huff, file_tensor(data)


def entropy_decode_packed(data: bytes) -> Tuple[JpegInfo, torch.Tensor]:
    """JPEG bytes -> (info, uint8 tensor holding the PACKED coefficients: per block only the coefficients up to the last
    non-zero one in scan order, see include/b2c.h).  This is what the DataLoader workers hand to the main process: 4-10x
    smaller than the dense form.  Host only; safe in worker processes and threads."""
    lib = _lib.load()
    info = JpegInfo()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    rc = lib.b2c_jpeg_parse(buf, len(data), C.byref(info))
    if rc != 0:
        _raise(rc, "b2c_jpeg_parse")
    if info.width * info.height > MAX_PIXELS:
        raise UnsupportedJPEG(f"{info.width}x{info.height} pixels exceeds the device path's limit of {MAX_PIXELS}")
    # per-thread scratch: the dense decode target and a worst-case packed buffer, grown on demand
    sc = getattr(_scratch, "dense", None)
    if sc is None or sc.numel() < info.coef_count:
        sc = _scratch.dense = torch.empty(int(info.coef_count), dtype=torch.int16)
    pk = getattr(_scratch, "packed", None)
    if pk is None or pk.numel() < info.packed_capacity + 16:
        pk = _scratch.packed = torch.empty(int(info.packed_capacity) + 16, dtype=torch.uint8)
    base = (pk.data_ptr() + 15) & ~15
    rc = lib.b2c_jpeg_decode_packed(buf, len(data), C.byref(info), C.c_void_p(base), int(info.packed_capacity), C.c_void_p(sc.data_ptr()))
    if rc != 0:
        _raise(rc, "b2c_jpeg_decode_packed")
    out = torch.empty(int(info.packed_bytes), dtype=torch.uint8)
    C.memmove(out.data_ptr(), base, int(info.packed_bytes))  # (Tensor.clone would wake the intra-op thread pool for 100 KB)
    return info, out


def expand_packed(info: JpegInfo, packed: torch.Tensor) -> torch.Tensor:
    """Packed form -> dense int16 coefficients in natural order (what entropy_decode returns): tests / the oracle."""
    zz = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35,
          42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    p = packed.numpy()
    offs = p[info.offs_off:info.offs_off + 4 * info.nblocks].view(np.uint32).astype(np.int64)
    counts = p[info.counts_off:info.counts_off + info.nblocks].astype(np.int64)
    vals = p[info.vals_off:].view(np.int16)
    assert counts.max(initial=0) <= 64 and (offs + counts).max(initial=0) <= vals.size
    dense = np.zeros((info.nblocks, 64), np.int16)
    zz = np.asarray(zz)
    for b in np.nonzero(counts)[0]:
        dense[b, zz[:counts[b]]] = vals[offs[b]:offs[b] + counts[b]]
    return torch.from_numpy(dense.reshape(-1))


def entropy_decode(data: bytes, pin: bool = False) -> Tuple[JpegInfo, torch.Tensor]:
    """JPEG bytes -> (info, DENSE int16 coefficient tensor on the host, natural order).  Host only; safe in worker
    processes.  The dense form is the oracle's input and the reference point of the packed form."""
    lib = _lib.load()
    info = JpegInfo()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    rc = lib.b2c_jpeg_parse(buf, len(data), C.byref(info))
    if rc != 0:
        _raise(rc, "b2c_jpeg_parse")
    if info.width * info.height > MAX_PIXELS:
        # a header can claim 65535 x 65535: leave such files to Pillow, whose decompression-bomb guard decides
        raise UnsupportedJPEG(f"{info.width}x{info.height} pixels exceeds the device path's limit of {MAX_PIXELS}")
    coefs = torch.empty(int(info.coef_count), dtype=torch.int16, pin_memory=pin)
    rc = lib.b2c_jpeg_decode_coefs(buf, len(data), C.byref(info), C.c_void_p(coefs.data_ptr()), coefs.numel())
    if rc != 0:
        _raise(rc, "b2c_jpeg_decode_coefs")
    return info, coefs


class Staging:
    """Grow-only pinned host buffer a batch of host tensors (one dtype) is gathered into before ONE async H2D copy.  The
    gather is spread over a few threads: the sources are fresh shared-memory mappings from the DataLoader workers, and
    faulting their pages in is what costs time when the workers are busy writing the next batches.  The event guards the
    buffer against being refilled while the previous copy still reads it."""
    _pool = None

    def __init__(self, dtype):
        self.dtype = dtype
        self.buf = None
        self.event = None
        self.timing = {}  # seconds per step when B2C_DRIVER_TIMING is set (diagnostics of the embedding driver)

    def gather(self, tensors, dev, threads: int = 4):
        """tensors: contiguous host tensors of self.dtype -> (device buffer holding them back to back, element offsets)."""
        import concurrent.futures as cf
        import os
        import time
        t0 = time.perf_counter()
        sizes = [t.numel() for t in tensors]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(offs[-1])
        if self.buf is None or self.buf.numel() < total:
            self.buf = torch.empty(max(total, 1 << 20), dtype=self.dtype, pin_memory=True)
            self.event = None
        if self.event is not None:
            self.event.synchronize()
        t1 = time.perf_counter()
        flat = self.buf[:total]

        def part(lo, hi):
            if hi > lo:
                torch.cat([t.reshape(-1) for t in tensors[lo:hi]], out=flat[int(offs[lo]):int(offs[hi])])

        n = len(tensors)
        if n >= 2 * threads and total * flat.element_size() >= (8 << 20):
            if Staging._pool is None:
                Staging._pool = cf.ThreadPoolExecutor(threads)
            cuts = [n * i // threads for i in range(threads + 1)]
            list(Staging._pool.map(lambda i: part(cuts[i], cuts[i + 1]), range(threads)))
        else:
            part(0, n)
        t2 = time.perf_counter()
        dflat = flat.to(dev, non_blocking=True)
        self.event = torch.cuda.Event(blocking=True)
        self.event.record()
        if os.environ.get("B2C_DRIVER_TIMING") is not None:
            t3 = time.perf_counter()
            for k, v in (("staging: wait for the previous H2D", t1 - t0), ("staging: gather into pinned memory", t2 - t1),
                         ("staging: enqueue H2D", t3 - t2)):
                self.timing[k] = self.timing.get(k, 0.0) + v
        return dflat, offs


_coef_staging = Staging(torch.int16)


def reconstruct_device(infos: Sequence[JpegInfo], dflat: torch.Tensor) -> List[torch.Tensor]:
    """Device stage alone: `dflat` = the batch's coefficient buffers back to back on the device, in `infos` order."""
    lib = _lib.load()
    n = len(infos)
    dev = dflat.device
    with torch.cuda.device(dev):
        arr = (JpegInfo * n)(*infos)
        outs = [torch.empty(i.height, i.width, 3, dtype=torch.uint8, device=dev) for i in infos]
        need = C.c_size_t()
        _lib.check(lib.b2c_jpeg_workspace_bytes(arr, n, C.byref(need)), "b2c_jpeg_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        offs = np.concatenate([[0], np.cumsum([int(i.coef_count) for i in infos])])
        assert int(offs[-1]) <= dflat.numel() and dflat.dtype == torch.int16 and dflat.is_contiguous()
        cptr = (C.c_void_p * n)(*[dflat.data_ptr() + 2 * int(o) for o in offs[:-1]])
        optr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        pitch = (C.c_int * n)(*[3 * i.width for i in infos])
        _lib.check(lib.b2c_jpeg_reconstruct(arr, cptr, optr, pitch, n, C.c_void_p(ws.data_ptr()), ws.numel(),
                                            C.c_void_p(_lib.current_stream_ptr())), "b2c_jpeg_reconstruct")
        # ws / dflat go back to torch's caching allocator in stream order, after the launches above
    return outs


_packed_staging = Staging(torch.uint8)


def reconstruct_packed(items: Sequence[Tuple[JpegInfo, torch.Tensor]], device="cuda") -> List[torch.Tensor]:
    """[(info, host PACKED coefficients)] -> [uint8 [H,W,3] device tensors]: one gather into pinned memory, one H2D copy
    and two kernel launches per batch."""
    if not items:
        return []
    lib = _lib.load()
    dev = torch.device(device)
    n = len(items)
    infos = [it[0] for it in items]
    with torch.cuda.device(dev):
        dflat, offs = _packed_staging.gather([it[1] for it in items], dev)  # every buffer is a multiple of 16 bytes
        arr = (JpegInfo * n)(*infos)
        outs = [torch.empty(i.height, i.width, 3, dtype=torch.uint8, device=dev) for i in infos]
        need = C.c_size_t()
        _lib.check(lib.b2c_jpeg_workspace_bytes(arr, n, C.byref(need)), "b2c_jpeg_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        pptr = (C.c_void_p * n)(*[dflat.data_ptr() + int(o) for o in offs[:-1]])
        optr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        pitch = (C.c_int * n)(*[3 * i.width for i in infos])
        _lib.check(lib.b2c_jpeg_reconstruct_packed(arr, pptr, optr, pitch, n, C.c_void_p(ws.data_ptr()), ws.numel(),
                                                   C.c_void_p(_lib.current_stream_ptr())), "b2c_jpeg_reconstruct_packed")
    return outs


def reconstruct(items: Sequence[Tuple[JpegInfo, torch.Tensor]], device="cuda") -> List[torch.Tensor]:
    """[(info, host coefficients)] -> [uint8 [H,W,3] device tensors]: one gather into pinned memory, one H2D copy and two
    kernel launches per batch."""
    if not items:
        return []
    dev = torch.device(device)
    with torch.cuda.device(dev):
        dflat, _ = _coef_staging.gather([it[1] for it in items], dev)
    return reconstruct_device([it[0] for it in items], dflat)


def decode_files(paths: Sequence[str], device="cuda") -> List[torch.Tensor]:
    """Convenience: read + entropy-decode on this thread, reconstruct on the device."""
    items = []
    for p in paths:
        with open(p, "rb") as fh:
            items.append(entropy_decode(fh.read()))
    return reconstruct(items, device)


_file_staging = Staging(torch.uint8)


def huffman_device(items: Sequence[Tuple[JpegInfo, JpegHuff, torch.Tensor]], device="cuda"):
    """[(info, huff, uint8 host tensor with the file's bytes)] -> (dense int16 coefficients of the batch back to back on
    the device, int32 status per image on the device).  One gather into pinned memory, one H2D copy of the FILES (not of
    coefficients: 5-25x fewer bytes than the packed form) and one kernel launch.  status != 0 means the stream has to go
    back to the host stage (damaged, truncated, or it did not synchronise)."""
    lib = _lib.load()
    dev = torch.device(device)
    n = len(items)
    infos = [it[0] for it in items]
    with torch.cuda.device(dev):
        # every file starts on a 16-byte boundary of the staging buffer
        files = [it[2] if it[2].numel() % 16 == 0 else torch.cat([it[2], it[2].new_zeros(16 - it[2].numel() % 16)])
                 for it in items]  # (file_tensor() pads already; a foreign tensor is padded here)
        dflat, offs = _file_staging.gather(files, dev)
        iarr = (JpegInfo * n)(*infos)
        harr = (JpegHuff * n)(*[it[1] for it in items])
        need = C.c_size_t()
        _lib.check(lib.b2c_jpeg_huff_workspace_bytes(harr, n, C.byref(need)), "b2c_jpeg_huff_workspace_bytes")
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) // 256 * 256
        coff = np.concatenate([[0], np.cumsum([int(i.coef_count) for i in infos])])
        coefs = torch.empty(int(coff[-1]), dtype=torch.int16, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        fptr = (C.c_void_p * n)(*[dflat.data_ptr() + int(o) for o in offs[:-1]])
        cptr = (C.c_void_p * n)(*[coefs.data_ptr() + 2 * int(o) for o in coff[:-1]])
        _lib.check(lib.b2c_jpeg_huff_decode(iarr, harr, fptr, cptr, C.c_void_p(status.data_ptr()), n, C.c_void_p(wp),
                                            need.value, C.c_void_p(_lib.current_stream_ptr())), "b2c_jpeg_huff_decode")
    return coefs, status


_side_streams = {}
_status_pinned = None


def _status_slot(n: int) -> torch.Tensor:
    global _status_pinned
    if _status_pinned is None or _status_pinned.numel() < n:
        _status_pinned = torch.empty(max(n, 1024), dtype=torch.int32, pin_memory=True)
    return _status_pinned[:n]


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def file_bytes(item) -> bytes:
    """The file behind a prepare_file() item (the tensor is padded to 16 bytes; the scan ends where the file does)."""
    n = int(item[1].scan_begin + item[1].scan_bytes)
    return bytes(item[2][:n].numpy())


def decode_device(items: Sequence[Tuple[JpegInfo, JpegHuff, torch.Tensor]], device="cuda", fallback: bool = True):
    """File bytes -> uint8 [H,W,3] device tensors with BOTH stages on the device (Huffman + reconstruction).  Returns
    (images, status list).  The H2D copy of the files and the Huffman kernel run on a side stream, and the host waits for
    THAT stream only before it reads the per-image status — work queued on the caller's stream (the previous batch's
    tower) keeps the GPU busy meanwhile.  An image whose status is non-zero (damaged, truncated, not synchronised) is
    decoded again through the host stage and then through Pillow, which has the last word like in the reference
    (utils/embedder.py:167,176-181); it is None if that fails too (or when `fallback` is off)."""
    if not items:
        return [], []
    dev = torch.device(device)
    infos = [it[0] for it in items]
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        with torch.cuda.stream(side):
            coefs, status = huffman_device(items, dev)
            # waits for the side stream only, asleep (blocking event) rather than spinning on a core the workers need
            st_host = _status_slot(len(items))
            st_host.copy_(status, non_blocking=True)
            ev = torch.cuda.Event(blocking=True)
            ev.record(side)
            ev.synchronize()
            st = st_host.tolist()
        main.wait_stream(side)
        coefs.record_stream(main)
        outs = reconstruct_device(infos, coefs)
    for i, code in enumerate(st):
        if code:
            outs[i] = _decode_on_host(file_bytes(items[i]), dev) if fallback else None
    return outs, st


def _decode_on_host(data: bytes, dev):
    import io
    from PIL import Image
    try:
        return reconstruct([entropy_decode(data)], dev)[0]
    except (UnsupportedJPEG, _lib.B2CError):
        pass
    try:
        with Image.open(io.BytesIO(data)) as im:
            arr = np.array(im.convert("RGB"))
        return torch.from_numpy(arr).to(dev)
    except Exception as e:  # noqa: BLE001 — Pillow raises OSError / SyntaxError / ValueError for damaged files
        print(f"Error decoding JPEG stream: {e}")
        return None
