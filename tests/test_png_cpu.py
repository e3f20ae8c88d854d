"""cvsteer-run's PNG codec (cvsteer_b200/cli/png_io.h) against OpenCV: the gray image it decodes must be what
`cv::imread(name)` + `cv::cvtColor(BGR2GRAY)` give (reference example/steer.cpp:73-82), and what it encodes must read back
unchanged (example/steer.cpp:106-121).  CPU only; exotic files (palette, sub-byte depths, Adam7, every filter type) come from
the small PNG writer below, because cv2.imencode cannot produce them."""
import os
import struct
import subprocess
import zlib

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tests", "cpp", "png_tool")


@pytest.fixture(scope="module")
def tool():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "png_tool"], check=True, capture_output=True)
    return TOOL


def _read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    magic, dims, maxv, rest = data.split(b"\n", 3)
    assert magic == b"P5" and maxv == b"255"
    cols, rows = map(int, dims.split())
    return np.frombuffer(rest, np.uint8, rows * cols).reshape(rows, cols)


def _decode(tool, png_bytes, tmp_path, name="x"):
    src, dst = tmp_path / (name + ".png"), tmp_path / (name + ".pgm")
    src.write_bytes(png_bytes)
    r = subprocess.run([tool, "decode", str(src), str(dst)])
    return _read_pgm(dst) if r.returncode == 0 else None


def _cv_gray(png_bytes):
    img = cv2.imdecode(np.frombuffer(png_bytes, np.uint8), cv2.IMREAD_COLOR)   # what cv::imread(name) returns
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)


def _chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows_bytes, bpp, ftype_of_row):
    """PNG filtering of a list of byte rows; ftype_of_row(y) picks the filter type 0..4."""
    out, prev = bytearray(), None
    for y, row in enumerate(rows_bytes):
        ft = ftype_of_row(y)
        out.append(ft)
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i] if prev is not None else 0
            c = prev[i - bpp] if prev is not None and i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            out.append((v - pred) & 255)
        prev = row
    return bytes(out)


def _pack(samples, depth):
    """samples: (rows, n) array of integers below 2^depth -> list of packed byte rows (big-endian within the byte / sample)."""
    rows = []
    for r in samples:
        if depth == 8:
            rows.append(bytes(int(v) for v in r))
        elif depth == 16:
            rows.append(b"".join(struct.pack(">H", int(v)) for v in r))
        else:
            per = 8 // depth
            b = bytearray((len(r) + per - 1) // per)
            for i, v in enumerate(r):
                b[i // per] |= int(v) << ((per - 1 - i % per) * depth)
            rows.append(bytes(b))
    return rows


def make_png(samples, depth, ctype, palette=None, interlace=False, ftype=lambda y: y % 5):
    """samples: (rows, cols, channels) integers; channels must match the colour type."""
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        raw += _filter_rows(_pack(sub.reshape(sub.shape[0], -1), depth), bpp, ftype)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        png += _chunk(b"PLTE", bytes(int(v) for v in palette.reshape(-1)))
    comp = zlib.compress(raw, 6)
    half = len(comp) // 2                      # two IDAT chunks: the stream may be split anywhere
    return png + _chunk(b"IDAT", comp[:half]) + _chunk(b"IDAT", comp[half:]) + _chunk(b"tEXt", b"k\0v") + _chunk(b"IEND", b"")


def test_cv2_written_pngs_decode_like_imread_plus_cvtcolor(tool, tmp_path, fish_fixture):
    rng = np.random.default_rng(11)
    fish = fish_fixture["fish"]
    cases = {
        "gray8": fish,
        "bgr8": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
        "bgra8": rng.integers(0, 256, (20, 31, 4), dtype=np.uint8),
        "gray16": rng.integers(0, 65536, (19, 23), dtype=np.uint16),
        "bgr16": rng.integers(0, 65536, (9, 14, 3), dtype=np.uint16),
        "one_pixel": np.array([[200]], np.uint8),
    }
    for name, img in cases.items():
        ok, buf = cv2.imencode(".png", img)
        assert ok
        got = _decode(tool, buf.tobytes(), tmp_path, name)
        assert got is not None, name
        assert np.array_equal(got, _cv_gray(buf.tobytes())), name
    assert np.array_equal(_decode(tool, cv2.imencode(".png", fish)[1].tobytes(), tmp_path, "fish"), fish)


@pytest.mark.parametrize("interlace", [False, True])
def test_handwritten_pngs_every_type_depth_and_filter(tool, tmp_path, interlace):
    rng = np.random.default_rng(12 + interlace)
    h, w = 21, 19                                        # not multiples of 8: ragged Adam7 passes, ragged packed rows
    cases = []
    for depth in (1, 2, 4, 8, 16):
        cases.append(("gray%d" % depth, rng.integers(0, 1 << depth, (h, w, 1)), depth, 0, None))
    for depth in (1, 2, 4, 8):
        pal = rng.integers(0, 256, (1 << depth, 3))
        cases.append(("pal%d" % depth, rng.integers(0, 1 << depth, (h, w, 1)), depth, 3, pal))
    for depth in (8, 16):
        cases.append(("rgb%d" % depth, rng.integers(0, 1 << depth, (h, w, 3)), depth, 2, None))
        cases.append(("ga%d" % depth, rng.integers(0, 1 << depth, (h, w, 2)), depth, 4, None))
        cases.append(("rgba%d" % depth, rng.integers(0, 1 << depth, (h, w, 4)), depth, 6, None))
    for name, s, depth, ctype, pal in cases:
        png = make_png(s, depth, ctype, pal, interlace)
        want = _cv_gray(png)                              # OpenCV reads the file the helper wrote: the helper is sound
        got = _decode(tool, png, tmp_path, name)
        assert got is not None, name
        assert np.array_equal(got, want), (name, interlace)


def test_damaged_files_are_rejected(tool, tmp_path):
    ok, buf = cv2.imencode(".png", np.arange(64, dtype=np.uint8).reshape(8, 8))
    good = buf.tobytes()
    assert _decode(tool, good, tmp_path, "good") is not None
    bad_crc = bytearray(good)
    bad_crc[20] ^= 1                                       # inside IHDR: CRC mismatch
    assert _decode(tool, bytes(bad_crc), tmp_path, "crc") is None
    assert _decode(tool, good[: len(good) // 2], tmp_path, "cut") is None
    assert _decode(tool, b"P5\n2 2\n255\n\0\0\0\0", tmp_path, "notpng") is None
    # a header that promises far more pixels than the IDAT bytes could inflate to (found by fuzzing: it used to allocate
    # tens of GB before looking at the data) is rejected up front
    huge = bytearray(good)
    huge[16:24] = struct.pack(">II", 0x40000, 0x40000)
    ihdr = bytes(huge[12:29])
    huge[29:33] = struct.pack(">I", zlib.crc32(ihdr) & 0xFFFFFFFF)
    assert _decode(tool, bytes(huge), tmp_path, "huge") is None


def test_encoder_round_trips_through_opencv(tool, tmp_path, fish_fixture):
    for name, img in (("fish", fish_fixture["fish"]), ("noise", np.random.default_rng(13).integers(0, 256, (33, 1), dtype=np.uint8)),
                      ("wide", np.random.default_rng(14).integers(0, 256, (3, 700), dtype=np.uint8))):
        src, dst = tmp_path / (name + ".pgm"), tmp_path / (name + ".png")
        src.write_bytes(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]) + img.tobytes())
        assert subprocess.run([tool, "encode", str(src), str(dst)]).returncode == 0
        back = cv2.imread(str(dst), cv2.IMREAD_UNCHANGED)
        assert back is not None and back.dtype == np.uint8 and back.ndim == 2
        assert np.array_equal(back, img), name
        assert np.array_equal(_decode(tool, dst.read_bytes(), tmp_path, name + "_rt"), img)
