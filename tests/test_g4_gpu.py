"""G4/H4: basis, scalar / map steering, magnitude + phase -- class surface and fused batch path vs the oracle."""
import numpy as np
import pytest
import torch

import cvsteer_b200 as cb
from cvsteer_b200 import capi
from cvsteer_b200.batch import G4Batch
from oracle import cvsteer_ref as ref
from tests.util import assert_angle_close, assert_close_range, basis_range, synth

pytestmark = pytest.mark.gpu
P = ref.SteerableFiltersG4.PLANES


def test_fish_golden(fish_fixture, fish_oracle):
    f = cb.SteerableFiltersG4(fish_fixture["fish"].astype(np.float32))
    rng = basis_range([fish_oracle[k] for k in P])
    for k in P:
        assert_close_range(getattr(f, k), fish_oracle[k], rng, k)
    g4, h4 = f.steer(0.3)
    assert_close_range(g4, fish_oracle["g4_s03"], rng, "g4(0.3)")
    assert_close_range(h4, fish_oracle["h4_s03"], rng, "h4(0.3)")
    assert abs(float(g4[92, 128]) - 191.86337) < 2e-2 and abs(float(h4[92, 128]) - 16.558273) < 2e-2
    g4, h4, mag, ph = f.steer(fish_oracle["theta"], with_phase=True)
    assert_close_range(g4, fish_oracle["g4_map"], rng, "g4 map")
    assert_close_range(h4, fish_oracle["h4_map"], rng, "h4 map")
    assert_close_range(mag, fish_oracle["mag4"], rng, "mag4")
    assert_angle_close(ph, fish_oracle["phase4"], fish_oracle["mag4"], 2 * np.pi, "phase4")


@pytest.mark.parametrize("shape", [(131, 257), (64, 128), (13, 13), (7, 7), (3, 5), (1, 1), (5, 300)])
def test_class_vs_oracle(shape):
    img = synth(4000 + shape[1], *shape)
    o = ref.SteerableFiltersG4(img)
    f = cb.SteerableFiltersG4(img)
    rng = max(basis_range([getattr(o, k) for k in P]), 1.0)
    for k in P:
        assert_close_range(getattr(f, k), getattr(o, k), rng, k)
    for th in (0.3, -1.9):
        g4, h4 = f.steer(th)
        w = o.steer_scalar(th)
        assert_close_range(g4, w[0], rng, "g4 scalar")
        assert_close_range(h4, w[1], rng, "h4 scalar")
    th = np.random.default_rng(1).uniform(-3.5, 3.5, shape).astype(np.float32)
    g4, h4, mag, ph = f.steer(th, with_phase=True)
    w = o.steer_map_full(th)
    assert_close_range(g4, w[0], rng, "g4 map")
    assert_close_range(h4, w[1], rng, "h4 map")
    assert_close_range(mag, w[2], rng, "mag")
    assert_angle_close(ph, w[3], w[2], 2 * np.pi, "phase")


@pytest.mark.parametrize("width,spacing", [(4, 0.7), (8, 0.4)])
def test_generic_width(width, spacing):
    img = synth(4100, 60, 77)
    o = ref.SteerableFiltersG4(img, width, spacing)
    f = cb.SteerableFiltersG4(img, width, spacing)
    rng = basis_range([getattr(o, k) for k in P])
    for k in P:
        assert_close_range(getattr(f, k), getattr(o, k), rng, k)


def test_fused_batch_steer_phase():
    """Config 4: steer with a per-pixel theta map -> g4, h4, magnitude, phase in one launch."""
    fr = np.stack([synth(4200 + i, 150, 260) for i in range(2)])
    th = np.stack([ref.SteerableFiltersG2(f).theta for f in fr])
    g = G4Batch()
    r = g.run(torch.from_numpy(fr).cuda(), capi.G4_MASK_STEER | capi.G4_MASK_BASIS, steer=capi.STEER_MAP,
              theta_map=torch.from_numpy(th).cuda())
    for i in range(2):
        o = ref.SteerableFiltersG4(fr[i])
        rng = basis_range([getattr(o, k) for k in P])
        for k in P:
            assert_close_range(r[k][i].cpu().numpy(), getattr(o, k), rng, k)
        w = o.steer_map_full(th[i])
        assert_close_range(r["g4"][i].cpu().numpy(), w[0], rng, "g4")
        assert_close_range(r["h4"][i].cpu().numpy(), w[1], rng, "h4")
        assert_close_range(r["magnitude"][i].cpu().numpy(), w[2], rng, "mag")
        assert_angle_close(r["phase"][i].cpu().numpy(), w[3], w[2], 2 * np.pi, "phase")
    r2 = g.run(torch.from_numpy(fr).cuda(), capi.G4_MASK_STEER, steer=capi.STEER_SCALAR, theta=0.3)
    w = ref.SteerableFiltersG4(fr[0]).steer_scalar(0.3)
    assert_close_range(r2["g4"][0].cpu().numpy(), w[0], 1000.0, "g4 scalar fused")


def test_g4_band_equals_whole_and_u8_input():
    from cvsteer_b200.batch import Band
    H, W = 260, 300
    img = synth(4300, H, W)
    g = G4Batch()
    th = np.random.default_rng(2).uniform(-1.5, 1.5, (1, H, W)).astype(np.float32)
    mask = capi.G4_MASK_STEER
    whole = g.run(torch.from_numpy(img[None]).cuda(), mask, steer=capi.STEER_MAP, theta_map=torch.from_numpy(th).cuda())
    for (r0, r1) in ((0, 64), (64, 200), (200, 260)):
        lo, hi = max(0, r0 - 6), min(H, r1 + 6)
        part = g.run(torch.from_numpy(img[None, lo:hi].copy()).cuda(), mask, steer=capi.STEER_MAP,
                     theta_map=torch.from_numpy(th[:, r0:r1].copy()).cuda(), band=Band(full_rows=H, y_origin=lo, row_begin=r0, row_end=r1))
        for k in whole:
            assert torch.equal(part[k][0], whole[k][0, r0:r1]), (k, r0)
    img8 = np.random.default_rng(3).integers(0, 256, (1, 90, 140), dtype=np.uint8)
    a = g.run(torch.from_numpy(img8).cuda(), capi.G4_MASK_BASIS)
    b = g.run(torch.from_numpy(img8.astype(np.float32)).cuda(), capi.G4_MASK_BASIS)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def _g4_orientation_bruteforce(o):
    """C2, C3 of E(theta) = G4(theta)^2 + H4(theta)^2 by steering the ORACLE at 32 angles and projecting on cos/sin 2theta
    (exact for this degree-10 trigonometric polynomial), then theta_d / strength as the reference defines them for G2."""
    n = 32
    c2 = np.zeros(o.g4a.shape, np.float64)
    c3 = np.zeros_like(c2)
    for t in range(n):
        th = 2 * np.pi * t / n
        g4, h4 = o.steer_scalar(th)
        e = g4.astype(np.float64) ** 2 + h4.astype(np.float64) ** 2
        c2 += 2 * e * np.cos(2 * th) / n
        c3 += 2 * e * np.sin(2 * th) / n
    return 0.5 * np.arctan2(c3, c2), np.hypot(c2, c3)


def test_g4_dominant_orientation_extension():
    """Row f4: theta_d / strength for G4 (absent in the reference), class API + fused batch + dominant-angle steering."""
    img = synth(4400, 120, 170)
    o = ref.SteerableFiltersG4(img)
    theta_bf, strength_bf = _g4_orientation_bruteforce(o)
    rng = basis_range([getattr(o, k) for k in P])
    f = cb.SteerableFiltersG4(img)
    th, sg = f.getDominantOrientationAngle(), f.getDominantOrientationStrength()
    assert_close_range(sg, strength_bf.astype(np.float32), ("own", rng), "G4 strength", rtol=2e-4)
    assert_angle_close(th, theta_bf.astype(np.float32), strength_bf, np.pi, "G4 theta_d", thresh_frac=1e-2)
    g = G4Batch()
    x = torch.from_numpy(img[None]).cuda()
    mask = capi.bit(capi.G4_THETA) | capi.bit(capi.G4_STRENGTH) | capi.G4_MASK_STEER
    r = g.run(x, mask, steer=capi.STEER_DOMINANT)
    assert_angle_close(r["theta"][0].cpu().numpy(), th, strength_bf, np.pi, "fused theta == class theta", tol=1e-5, thresh_frac=1e-2)
    w = o.steer_map_full(r["theta"][0].cpu().numpy())
    assert_close_range(r["g4"][0].cpu().numpy(), w[0], rng, "g4 at theta_d")
    assert_close_range(r["h4"][0].cpu().numpy(), w[1], rng, "h4 at theta_d")
    g4c, h4c = f.steer(None)                     # class API: steer at the handle's own dominant map
    wc = o.steer_map(th)
    assert_close_range(g4c, wc[0], rng, "class g4 at theta_d")
    # an oriented grating inside the G4/H4 pass band (1.5 rad/px; far below it the 13-tap sampled filters are no longer
    # exactly steerable and the G4 estimate drifts, 0.13 rad at 0.5 rad/px -- a property of the reference's taps):
    # G4's dominant orientation must agree with G2's
    y, xx = np.mgrid[0:128, 0:160].astype(np.float32)
    for ang in (0.3, 1.2, -0.6):
        gr = (127 + 100 * np.cos(1.5 * (xx * np.cos(ang) + y * np.sin(ang)))).astype(np.float32)
        t4 = cb.SteerableFiltersG4(gr).getDominantOrientationAngle()[30:-30, 30:-30]
        t2 = cb.SteerableFiltersG2(gr).getDominantOrientationAngle()[30:-30, 30:-30]
        d = np.abs(t4 - t2) % np.pi
        assert float(np.median(np.minimum(d, np.pi - d))) < 0.02, ang


def test_g4_flip_symmetry_and_crop_parity_4k():
    """Config-4 frame size (3840x2160): size-independent properties of the G4/H4 basis (a horizontal flip mirrors the
    planes whose row filter is even and negates those whose row filter is odd, exactly), plus oracle parity on a crop."""
    rs = np.random.default_rng(321)
    img = rs.uniform(0, 255, (1, 2160, 3840)).astype(np.float32)
    g = G4Batch()
    x = torch.from_numpy(img).cuda()
    a = g.run(x, capi.G4_MASK_BASIS)
    b = g.run(torch.flip(x, dims=[2]).contiguous(), capi.G4_MASK_BASIS)
    sign = {"g4a": 1, "g4b": -1, "g4c": 1, "g4d": -1, "g4e": 1, "h4a": -1, "h4b": 1, "h4c": -1, "h4d": 1, "h4e": -1, "h4f": 1}
    for k, s in sign.items():
        assert torch.equal(torch.flip(b[k], dims=[2]), s * a[k]), k
    o = ref.SteerableFiltersG4(img[0, 900:1100, 1500:1800])
    rng = basis_range([getattr(o, k) for k in P])
    for k in P:
        assert_close_range(a[k][0, 912:1088, 1512:1788].cpu().numpy(), getattr(o, k)[12:-12, 12:-12], rng, "crop " + k)
    th = torch.rand((1, 2160, 3840), device="cuda") * 3 - 1.5
    r = g.run(x, capi.G4_MASK_STEER, steer=capi.STEER_MAP, theta_map=th)
    w = o.steer_map_full(th[0, 900:1100, 1500:1800].cpu().numpy())
    assert_close_range(r["g4"][0, 912:1088, 1512:1788].cpu().numpy(), w[0][12:-12, 12:-12], rng, "crop g4")
    assert_close_range(r["magnitude"][0, 912:1088, 1512:1788].cpu().numpy(), w[2][12:-12, 12:-12], rng, "crop magnitude")


def test_g4_host_batch_api_matches_device_path():
    """cvs_g4_run_batch_host: chunked H2D / kernel / D2H pipeline == one device launch (scalar angle and theta_d)."""
    # (192 columns: rows are 16-byte aligned on both paths, so both take the same TMA kernel -- the steer-only kernels with
    # baked taps fold the binomial factors into the column taps, which rounds differently from the constant-bank variant)
    fr = np.stack([synth(4700 + i, 110, 192) for i in range(6)])
    g = G4Batch()
    for theta, steer in ((0.45, capi.STEER_SCALAR), (None, capi.STEER_DOMINANT)):
        mask = capi.G4_MASK_STEER | (capi.bit(capi.G4_THETA) | capi.bit(capi.G4_STRENGTH) if theta is None else 0)
        dev = g.run(torch.from_numpy(fr).cuda(), mask, steer=steer, theta=theta or 0.0)
        planes = [p for p in range(capi.G4_NPLANES) if mask >> p & 1]
        oh = {p: torch.empty((6, 110, 192), dtype=torch.float32).pin_memory() for p in planes}
        g.run_host(torch.from_numpy(fr).pin_memory(), mask, oh, theta=theta)
        for p in planes:
            assert torch.equal(oh[p], dev[capi.G4_PLANE_NAMES[p]].cpu()), capi.G4_PLANE_NAMES[p]
    w = ref.SteerableFiltersG4(fr[2]).steer_scalar(0.45)          # and the oracle, on one frame
    oh = {capi.G4T: torch.empty((6, 110, 192), dtype=torch.float32)}
    g.run_host(torch.from_numpy(fr), capi.bit(capi.G4T), oh, theta=0.45)
    assert_close_range(oh[capi.G4T][2].numpy(), w[0], 1000.0, "g4 host scalar")
    with pytest.raises(capi.CvsError):                             # a G2 handle is refused
        from cvsteer_b200.batch import G2Batch
        g2 = G2Batch()
        capi.check(capi.lib().cvs_g4_run_batch_host(g2._h, 0, 1, 1, 1, 4, 4, 1, 0, 0.0, None, 4, 4))


@pytest.mark.parametrize("seed", range(10))
def test_g4_random_sweep_vs_oracle(seed):
    """Seeded sweep: frame size (down to smaller than the 13-tap filter), batch size, 8-bit / float input, padded views,
    random plane masks, steering from a random per-pixel angle map."""
    r = np.random.default_rng(9500 + seed)
    rows = int(r.choice([r.integers(1, 16), r.integers(16, 140), r.integers(140, 300)]))
    cols = int(r.choice([r.integers(1, 16), r.integers(16, 140), r.integers(140, 400)]))
    n = int(r.integers(1, 4))
    u8 = bool(r.integers(0, 2))
    fr = r.integers(0, 256, (n, rows, cols), dtype=np.uint8) if u8 else r.uniform(0, 255, (n, rows, cols)).astype(np.float32)
    x = torch.from_numpy(fr).cuda()
    if r.integers(0, 2):
        off, extra = int(r.integers(0, 4)), int(r.integers(0, 9))
        buf = torch.zeros((n, rows + 1, cols + off + extra), dtype=x.dtype, device="cuda")
        buf[:, :rows, off:off + cols] = x
        x = buf[:, :rows, off:off + cols]
    th = r.uniform(-4, 4, (n, rows, cols)).astype(np.float32)
    steerable = (1 << capi.G4_THETA) - 1                    # basis + g4 / h4 / magnitude / phase
    mask = int(r.integers(1, steerable + 1))
    if r.integers(0, 3) == 0:
        mask = int(r.choice([capi.G4_MASK_BASIS, capi.G4_MASK_STEER]))
    g = G4Batch()
    res = g.run(x, mask, steer=capi.STEER_MAP, theta_map=torch.from_numpy(th).cuda())
    assert sorted(res) == sorted(capi.G4_PLANE_NAMES[p] for p in range(capi.G4_NPLANES) if mask >> p & 1)
    for i in range(n):
        o = ref.SteerableFiltersG4(fr[i].astype(np.float32))
        rng = max(basis_range([getattr(o, k) for k in P]), 1e-3)
        for k in P:
            if k in res:
                assert_close_range(res[k][i].cpu().numpy(), getattr(o, k), rng, f"{k} seed{seed}")
        w = o.steer_map_full(th[i])
        for k, j in (("g4", 0), ("h4", 1), ("magnitude", 2)):
            if k in res:
                assert_close_range(res[k][i].cpu().numpy(), w[j], rng, f"{k} seed{seed}")
        if "phase" in res:
            assert_angle_close(res["phase"][i].cpu().numpy(), w[3], w[2], 2 * np.pi, f"phase seed{seed}")
