"""CPU: the native scan-file decoder (rf_ingest_png, csrc/ingest.cu) against files written here
with every PNG scanline filter, against cv2.imread, and — in the authoring container — against the data/tiny scans
the reference reads with cv2.imread(..., IMREAD_GRAYSCALE) (parseData.py:160-179).  Host code only: no GPU needed."""
import glob
import os
import struct
import zlib

import numpy as np
import pytest

from oracle import ref_import as ri
from radarslampy_b200 import _ffi


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, img, filters, idat_chunk=4096, level=6):
    """Minimal 8-bit grayscale PNG writer: row r uses filter filters[r % len(filters)], IDAT split into small chunks."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    raw = bytearray()
    prev = np.zeros(w, np.int32)
    for r in range(h):
        row = img[r].astype(np.int32)
        ft = filters[r % len(filters)]
        left = np.concatenate([[0], row[:-1]])
        upleft = np.concatenate([[0], prev[:-1]])
        if ft == 0:
            f = row
        elif ft == 1:
            f = row - left
        elif ft == 2:
            f = row - prev
        elif ft == 3:
            f = row - ((left + prev) >> 1)
        else:
            f = row - np.array([_paeth(int(a), int(b), int(c)) for a, b, c in zip(left, prev, upleft)])
        raw.append(ft)
        raw += (f & 255).astype(np.uint8).tobytes()
        prev = row

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    z = zlib.compress(bytes(raw), level)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)))
        f.write(chunk(b"tEXt", b"Comment\x00radar"))
        for i in range(0, len(z), idat_chunk):
            f.write(chunk(b"IDAT", z[i:i + idat_chunk]))
        f.write(chunk(b"IEND", b""))


def test_every_filter_and_parallel_decode(tmp_path):
    rng = np.random.default_rng(0)
    imgs, paths = [], []
    for i, filt in enumerate([(0,), (1,), (2,), (3,), (4,), (0, 1, 2, 3, 4), (4, 3, 1)]):
        img = (rng.random((40, 97)) * 255).astype(np.uint8)
        img[5:20] = np.clip(np.cumsum(rng.integers(-3, 4, (15, 97)), axis=1) + 100, 0, 255)   # smooth rows: filters matter
        p = os.path.join(tmp_path, f"{1547131046353776 + i}.png")
        write_png(p, img, filt, idat_chunk=300 + 50 * i)
        imgs.append(img); paths.append(p)
    assert _ffi.png_info(paths[0]) == (40, 97)
    for threads in (1, 3, 0):
        out = _ffi.ingest_png(paths, threads=threads)
        assert out.shape == (7, 40, 97) and out.dtype == np.uint8
        assert all(np.array_equal(out[i], imgs[i]) for i in range(7))
    cv2 = pytest.importorskip("cv2")
    assert all(np.array_equal(cv2.imread(p, cv2.IMREAD_GRAYSCALE), im) for p, im in zip(paths, imgs))
    # a caller-provided (e.g. pinned) buffer is filled in place
    buf = np.zeros((8, 40, 97), np.uint8)
    got = _ffi.ingest_png(paths, out=buf)
    assert got.base is buf or got is buf[:7].base or np.shares_memory(got, buf)
    assert np.array_equal(buf[3], imgs[3])


def test_scan_sized_file_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    scan = (rng.exponential(20, (400, 3779))).clip(0, 255).astype(np.uint8)
    p = os.path.join(tmp_path, "scan.png")
    cv2 = pytest.importorskip("cv2")
    assert cv2.imwrite(p, scan)
    out = _ffi.ingest_png([p, p, p], threads=2)
    assert np.array_equal(out[0], scan) and np.array_equal(out[2], scan)
    with pytest.raises(ValueError, match="cannot open"):
        _ffi.ingest_png([os.path.join(tmp_path, "missing.png")])
    with pytest.raises(ValueError, match="expected"):
        _ffi.ingest_png([p], out=np.zeros((1, 400, 3000), np.uint8))
    bad = os.path.join(tmp_path, "bad.png")
    data = bytearray(open(p, "rb").read())
    data[len(data) // 2] ^= 0xFF
    open(bad, "wb").write(bytes(data))
    with pytest.raises(ValueError, match="CRC|corrupt|incomplete"):
        _ffi.ingest_png([bad])
    open(bad, "wb").write(bytes(data[:len(data) // 3]))
    with pytest.raises(ValueError):
        _ffi.ingest_png([bad])
    rgb = os.path.join(tmp_path, "rgb.png")
    cv2.imwrite(rgb, np.zeros((4, 4, 3), np.uint8))
    with pytest.raises(ValueError, match="grayscale"):
        _ffi.ingest_png([rgb])


@pytest.mark.skipif(not ri.available(), reason="reference checkout not present (GPU box)")
def test_reference_scans_decode_like_cv2_imread(golden):
    cv2 = pytest.importorskip("cv2")
    paths = sorted(glob.glob(os.path.join(ri.REFERENCE_ROOT, "data", "tiny", "radar", "*.png")))
    assert len(paths) == 11
    out = _ffi.ingest_png(paths)
    assert out.shape == (11, 400, 3779)
    for p, o in zip(paths, out):
        assert np.array_equal(o, cv2.imread(p, cv2.IMREAD_GRAYSCALE))
    fr = golden["tiny_frames"]
    assert all(np.array_equal(out[i], fr[f"raw_{i}"]) for i in range(3))
