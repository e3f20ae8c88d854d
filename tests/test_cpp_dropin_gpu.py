"""The drop-in C++ classes (include/cvsteer/*.h over the C ABI) run the reference's own test flow; every output is
compared with the oracle.  The C++ program is built by __graft_entry__.build() (tests/cpp/Makefile)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cvsteer_ref as ref
from tests.util import assert_angle_close, assert_close_range, basis_range

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "test_dropin")


def _load(d, name, shape):
    return np.fromfile(os.path.join(d, name + ".f32"), np.float32).reshape(shape)


def test_cpp_dropin_reference_flow(fish_fixture, tmp_path):
    assert os.path.exists(EXE), "tests/cpp/test_dropin not built (run __graft_entry__.build())"
    fish = fish_fixture["fish"]
    src = tmp_path / "fish.u8"
    fish.tofile(src)
    r = subprocess.run([EXE, str(src), str(fish.shape[0]), str(fish.shape[1]), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    shp = fish.shape
    o = ref.SteerableFiltersG2(fish)
    rng = basis_range([getattr(o, k) for k in o.PLANES])
    theta = _load(tmp_path, "theta", shp)
    assert_angle_close(theta, o.theta, o.strength, np.pi, "theta")
    assert_close_range(_load(tmp_path, "strength", shp), o.strength, ("own", rng), "strength")
    assert_close_range(_load(tmp_path, "g2a", shp), o.g2a, rng, "m_g2a via subclass")
    assert_close_range(_load(tmp_path, "c1", shp), o.c1, ("own", rng), "m_c1 via subclass")
    assert np.array_equal(_load(tmp_path, "taps_g1", (9,)), o.g1[0])
    w = o.steer_map_full(theta)
    for name, want, scale in (("g2", w[0], rng), ("h2", w[1], rng), ("e", w[2], ("own", rng)), ("magnitude", w[3], rng)):
        assert_close_range(_load(tmp_path, name, shp), want, scale, name)
    phase = _load(tmp_path, "phase", shp)
    assert_angle_close(phase, w[4], w[3], 2 * np.pi, "phase")
    mag = _load(tmp_path, "magnitude", shp)
    for name, fn in (("edges", ref.find_edges), ("linesDark", ref.find_dark_lines), ("linesBright", ref.find_bright_lines)):
        assert_close_range(_load(tmp_path, name, shp), fn(mag, phase), rng, name)
    ws = o.steer_scalar(0.3)
    assert_close_range(_load(tmp_path, "g2_s03", shp), ws[0], rng, "g2 scalar")
    assert_close_range(_load(tmp_path, "h2_s03", shp), ws[1], rng, "h2 scalar")
    pt = _load(tmp_path, "point", (5,))
    wp = o.steer_point((17, 5), 0.3, full=True)
    assert abs(pt[0] - wp[0]) <= 1e-4 * rng and abs(pt[1] - wp[1]) <= 1e-4 * rng and abs(pt[3] - wp[3]) <= 1e-4 * rng
    assert abs(pt[2] - wp[2]) <= 1e-4 * float(o.c1.max() - o.c1.min())
    assert float(np.max(np.abs(_load(tmp_path, "lambda", shp) - ref.phase_weights(phase, 1.0, True)))) <= 1e-5
    o4 = ref.SteerableFiltersG4(fish)
    rng4 = basis_range([getattr(o4, k) for k in o4.PLANES])
    w4 = o4.steer_map_full(theta)
    assert_close_range(_load(tmp_path, "g4", shp), w4[0], rng4, "g4")
    assert_close_range(_load(tmp_path, "h4", shp), w4[1], rng4, "h4")
    assert_close_range(_load(tmp_path, "mag4", shp), w4[2], rng4, "mag4")
    assert_angle_close(_load(tmp_path, "phase4", shp), w4[3], w4[2], 2 * np.pi, "phase4")
    assert_close_range(_load(tmp_path, "g4_s03", shp), o4.steer_scalar(0.3)[0], rng4, "g4 scalar")
    # extension: computeDominantOrientation() == the Python class's getters (same C-ABI call)
    import cvsteer_b200 as cb
    f4 = cb.SteerableFiltersG4(fish.astype(np.float32))
    assert np.array_equal(_load(tmp_path, "theta4", shp), f4.getDominantOrientationAngle())
    assert np.array_equal(_load(tmp_path, "strength4", shp), f4.getDominantOrientationStrength())


def test_cpp_dropin_eager_mirrors(fish_fixture, tmp_path):
    """-DCVSTEER_EAGER_HOST_MIRRORS: setup() fills every protected member itself, so subclass code written against the
    reference (reads m_g2a ... without asking) works with no source change; same outputs as the lazy build."""
    exe = os.path.join(HERE, "cpp", "test_dropin_eager")
    assert os.path.exists(exe), "tests/cpp/test_dropin_eager not built (run __graft_entry__.build())"
    fish = fish_fixture["fish"]
    src = tmp_path / "fish.u8"
    fish.tofile(src)
    r = subprocess.run([exe, str(src), str(fish.shape[0]), str(fish.shape[1]), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    o = ref.SteerableFiltersG2(fish)
    rng = basis_range([getattr(o, k) for k in o.PLANES])
    assert_close_range(_load(tmp_path, "g2a", fish.shape), o.g2a, rng, "eager m_g2a")


def test_plain_c_caller_on_gpu():
    exe = os.path.join(HERE, "cpp", "abi_c_check")
    assert os.path.exists(exe), "tests/cpp/abi_c_check not built (run __graft_entry__.build())"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "abi_c ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
