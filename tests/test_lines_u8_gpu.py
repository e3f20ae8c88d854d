"""'Next' rows f1/f2 of SURVEY section 8: the cvsteer-run per-file body as one call producing 8-bit maps, the device
float->u8 conversions (cv::normalize NORM_MINMAX / convertTo with gain), and the cvsteer-run CLI itself."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest
import torch

from cvsteer_b200 import capi
from oracle import cvsteer_ref as ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cvsteer_b200", "cli", "cvsteer-run")


def _oracle_maps(gray, gain=0.0):
    f, (g2, h2, e, mag, ph) = ref.g2_full(gray)
    maps = [ref.find_edges(mag, ph), ref.find_dark_lines(mag, ph), ref.find_bright_lines(mag, ph)]
    if gain > 0:
        return [cv2.convertScaleAbs(m, alpha=gain) if False else np.clip(np.rint(m * np.float32(gain)), 0, 255).astype(np.uint8) for m in maps]
    return [ref.normalize_minmax_u8(m) for m in maps]


def _assert_u8_close(got, want, name):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    # float maps agree to ~1e-6 of range; after x255/range scaling a pixel sitting on a rounding boundary may move by 1
    assert d.max() <= 1, (name, int(d.max()))
    assert (d > 0).mean() < 0.02, (name, float((d > 0).mean()))


def _lines(h, gray, gain):
    n, rows, cols = gray.shape
    outs = [np.zeros_like(gray) for _ in range(3)]
    capi.check(capi.lib().cvs_g2_lines_u8_host(h, gray.ctypes.data, n, rows, cols, cols, rows * cols, gain,
                                               outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, cols, rows * cols))
    return outs


def test_lines_u8_matches_reference_flow(fish_fixture):
    lib = capi.lib()
    h = C.c_void_p()
    capi.check(lib.cvs_g2_create(C.byref(h), 0, 4, 0.67))
    fish = fish_fixture["fish"]
    rng = np.random.default_rng(8)
    other = rng.integers(0, 256, fish.shape, dtype=np.uint8)
    batch = np.stack([fish, other])
    for gain in (0.0, 0.5):
        got = _lines(h, batch, gain)
        for i in range(2):
            want = _oracle_maps(batch[i], gain)
            for k, name in enumerate(("edges", "dark", "bright")):
                _assert_u8_close(got[k][i], want[k], f"{name} frame{i} gain{gain}")
    # the reference's acceptance test on these very outputs (test/test.cpp:97-103)
    got = _lines(h, fish[None], 0.0)
    for k, name in enumerate(("edges_gt", "lines_dark_gt", "lines_bright_gt")):
        ok, buf = cv2.imencode(".jpg", got[k][0])
        err = cv2.norm(cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE), fish_fixture[name], cv2.NORM_L1) / float(fish.size)
        assert err <= 1.0, (name, err)
    lib.cvs_g2_destroy(h)


def test_to_u8_dev_semantics():
    rs = np.random.default_rng(2)
    x = rs.normal(10, 40, (3, 50, 70)).astype(np.float32)
    x[1] = 5.0                                   # constant frame: max - min == 0 -> scale 0 -> all zeros (cv::normalize)
    t = torch.from_numpy(x).cuda()
    out = torch.empty((3, 50, 70), dtype=torch.uint8, device="cuda")
    for gain in (0.0, 2.0):
        capi.check(capi.lib().cvs_to_u8_dev(0, t.data_ptr(), 3, 50, 70, 70 * 4, 50 * 70 * 4, gain, out.data_ptr(), 70, 50 * 70,
                                            torch.cuda.current_stream().cuda_stream))
        got = out.cpu().numpy()
        for i in range(3):
            if gain > 0:
                want = np.clip(np.rint(x[i] * np.float32(gain)), 0, 255).astype(np.uint8)
            else:
                want = cv2.normalize(x[i], None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
            _assert_u8_close(got[i], want, f"to_u8 frame{i} gain{gain}")


def _write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n# test\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def _read_pgm(path):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    w, h = map(int, parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w)


def test_cvsteer_run_cli(fish_fixture, tmp_path):
    """The reference CLI's contract (example/steer.cpp:59-173): a list of image files in, three PNG maps per file out; PNG
    (gray and colour) and binary PGM inputs, an unreadable file skipped silently; --format=pgm for PGM outputs."""
    assert os.path.exists(CLI), "cvsteer-run not built (run __graft_entry__.build())"
    rng = np.random.default_rng(1)
    colour = rng.integers(0, 256, (60, 90, 3), dtype=np.uint8)
    imgs = {"fish": fish_fixture["fish"], "noise": rng.integers(0, 256, (97, 203), dtype=np.uint8),
            "colour": cv2.cvtColor(colour, cv2.COLOR_BGR2GRAY)}            # what the reference's imread + cvtColor sees
    _write_pgm(tmp_path / "fish.pgm", imgs["fish"])
    assert cv2.imwrite(str(tmp_path / "noise.png"), imgs["noise"]) and cv2.imwrite(str(tmp_path / "colour.png"), colour)
    (tmp_path / "text.png").write_text("not an image")
    lst = tmp_path / "files.txt"
    lst.write_text(f"{tmp_path}/fish.pgm\n{tmp_path}/noise.png\n{tmp_path}/colour.png\n{tmp_path}/missing.png\n{tmp_path}/text.png\n")
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([CLI, f"--input={lst}", f"--output={out}", "--verbose"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "3 of 5 files processed" in r.stdout            # unreadable files are skipped silently, like the reference
    for k, v in imgs.items():
        want = _oracle_maps(v)
        for j, suffix in enumerate(("edges", "lines_dark", "lines_bright")):
            got = cv2.imread(str(out / f"{k}_{suffix}.png"), cv2.IMREAD_UNCHANGED)   # the reference's output names
            assert got is not None and got.dtype == np.uint8 and got.shape == v.shape
            _assert_u8_close(got, want[j], f"{k}_{suffix}")
    # PGM outputs on request; a single image instead of a list
    out2 = tmp_path / "out2"
    out2.mkdir()
    r = subprocess.run([CLI, f"--input={tmp_path}/noise.png", f"--output={out2}", "--format=pgm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for suffix in ("edges", "lines_dark", "lines_bright"):
        assert np.array_equal(_read_pgm(out2 / f"noise_{suffix}.pgm"), cv2.imread(str(out / f"noise_{suffix}.png"), cv2.IMREAD_UNCHANGED))
    assert "Usage" in subprocess.run([CLI, "--help"], capture_output=True, text=True).stdout


def test_lines_u8_every_device_of_one_process(fish_fixture):
    """One process driving several GPUs (what cvsteer-run does: worker w -> device w % ndev).  The 8-bit TMA kernel asks
    for more than 48 KB of dynamic shared memory, a per-DEVICE function attribute: it must be set on each GPU, not once
    per process.  On a single-GPU box this still exercises the per-device bookkeeping on device 0."""
    lib = capi.lib()
    fish = fish_fixture["fish"]
    batch = np.stack([fish, fish[::-1].copy()])
    want = None
    for dev in range(torch.cuda.device_count()):
        h = C.c_void_p()
        capi.check(lib.cvs_g2_create(C.byref(h), dev, 4, 0.67))
        got = _lines(h, batch, 0.0)
        lib.cvs_g2_destroy(h)
        if want is None:
            want = got
            ref_maps = _oracle_maps(batch[1])
            for k, name in enumerate(("edges", "dark", "bright")):
                _assert_u8_close(got[k][1], ref_maps[k], name)
        else:
            for k in range(3):
                assert np.array_equal(got[k], want[k]), f"device {dev} differs from device 0 (plane {k})"
