"""Device-resident fused batch path (cvs_g2_run_batch_dev) vs the oracle, plus structural properties."""
import numpy as np
import pytest
import torch

from cvsteer_b200 import capi
from cvsteer_b200.batch import Band, G2Batch, pyr_down
from oracle import cvsteer_ref as ref
from tests.util import assert_angle_close, assert_close_range, basis_range, synth

pytestmark = pytest.mark.gpu
STATE = ("g2a", "g2b", "g2c", "h2a", "h2b", "h2c", "h2d")


def _frames(seed0, n, rows, cols):
    return np.stack([synth(seed0 + i, rows, cols) for i in range(n)])


def _check_full(res, i, o, rng, tag=""):
    th = res["theta"][i].cpu().numpy()
    assert_angle_close(th, o.theta, o.strength, np.pi, tag + "theta")
    assert_close_range(res["strength"][i].cpu().numpy(), o.strength, ("own", rng), tag + "strength")
    w = o.steer_map_full(th)
    assert_close_range(res["g2"][i].cpu().numpy(), w[0], rng, tag + "g2")
    assert_close_range(res["h2"][i].cpu().numpy(), w[1], rng, tag + "h2")
    assert_close_range(res["e"][i].cpu().numpy(), w[2], ("own", rng), tag + "e")
    assert_close_range(res["magnitude"][i].cpu().numpy(), w[3], rng, tag + "magnitude")
    assert_angle_close(res["phase"][i].cpu().numpy(), w[4], w[3], 2 * np.pi, tag + "phase")


@pytest.mark.parametrize("shape", [(72, 256), (130, 131), (200, 520), (64, 128), (1080, 1920)])
def test_modes_vs_oracle(shape):
    n = 3 if shape[0] < 1000 else 1
    fr = _frames(3000, n, *shape)
    x = torch.from_numpy(fr).cuda()
    g = G2Batch()
    m0 = g.run(x, capi.G2_MASK_STATE)
    m1 = g.run(x, capi.G2_MASK_ORIENT)
    m2 = g.run(x, capi.G2_MASK_FULL)
    assert "tma" in g.last_launch()["kernel"] or shape[1] % 4
    for i in range(n):
        o = ref.SteerableFiltersG2(fr[i])
        rng = basis_range([getattr(o, k) for k in STATE])
        for k in STATE:
            assert_close_range(m0[k][i].cpu().numpy(), getattr(o, k), rng, f"M0 {k}")
        for k in ("c1", "c2", "c3", "strength"):
            assert_close_range(m0[k][i].cpu().numpy(), getattr(o, k), ("own", rng), f"M0 {k}")
        assert_angle_close(m0["theta"][i].cpu().numpy(), o.theta, o.strength, np.pi, "M0 theta")
        # M1: energy at theta_d  (oracle: c1 + c2 cos 2t + c3 sin 2t at t = theta_d)
        _, _, e1 = ref.g2_orientation(fr[i])
        assert_close_range(m1["e"][i].cpu().numpy(), e1, ("own", rng), "M1 e")
        # M0 (the class state) runs the exact cv::cartToPolar sequence, M1/M2 the MUFU variants: same angle to rounding
        assert_angle_close(m1["theta"][i].cpu().numpy(), m0["theta"][i].cpu().numpy(), o.strength, np.pi, "M1 vs M0 theta", tol=2e-6)
        assert_close_range(m1["strength"][i].cpu().numpy(), m0["strength"][i].cpu().numpy(), float(o.strength.max()), "M1 vs M0 strength", rtol=1e-6)
        _check_full(m2, i, o, rng, "M2 ")
        assert torch.equal(m2["theta"][i], m1["theta"][i]) and torch.equal(m2["strength"][i], m1["strength"][i])


def test_all_planes_dynamic_mask_and_find_maps():
    fr = _frames(3100, 2, 96, 200)
    x = torch.from_numpy(fr).cuda()
    g = G2Batch()
    mask = (1 << capi.G2_NPLANES) - 1
    r = g.run(x, mask)
    for i in range(2):
        o = ref.SteerableFiltersG2(fr[i])
        rng = basis_range([getattr(o, k) for k in STATE])
        _check_full(r, i, o, rng, "dyn ")
        mag, ph = r["magnitude"][i].cpu().numpy(), r["phase"][i].cpu().numpy()
        for k, fn in (("edges", ref.find_edges), ("lines_dark", ref.find_dark_lines), ("lines_bright", ref.find_bright_lines)):
            assert_close_range(r[k][i].cpu().numpy(), fn(mag, ph), rng, k)


def test_scalar_and_map_steering_fused():
    fr = _frames(3200, 1, 80, 150)
    x = torch.from_numpy(fr).cuda()
    g = G2Batch()
    o = ref.SteerableFiltersG2(fr[0])
    rng = basis_range([getattr(o, k) for k in STATE])
    mask = capi.bit(capi.G2T) | capi.bit(capi.H2T) | capi.bit(capi.E) | capi.bit(capi.MAG) | capi.bit(capi.PHASE)
    r = g.run(x, mask, steer=capi.STEER_SCALAR, theta=0.7)
    w = o.steer_scalar_full(0.7)
    for k, j, s in (("g2", 0, rng), ("h2", 1, rng), ("e", 2, ("own", rng)), ("magnitude", 3, rng)):
        assert_close_range(r[k][0].cpu().numpy(), w[j], s, "scalar " + k)
    th = np.random.default_rng(9).uniform(-4, 4, (1, 80, 150)).astype(np.float32)
    r = g.run(x, mask, steer=capi.STEER_MAP, theta_map=torch.from_numpy(th).cuda())
    w = o.steer_map_full(th[0])
    for k, j, s in (("g2", 0, rng), ("h2", 1, rng), ("e", 2, ("own", rng)), ("magnitude", 3, rng)):
        assert_close_range(r[k][0].cpu().numpy(), w[j], s, "map " + k)
    assert_angle_close(r["phase"][0].cpu().numpy(), w[4], w[3], 2 * np.pi, "map phase")


def test_tma_and_ldg_loaders_agree_bitwise():
    """Same pixels through the TMA-staged tile and through the cooperative reflect-indexed loader."""
    fr = _frames(3300, 2, 150, 260)
    g = G2Batch()
    x = torch.from_numpy(fr).cuda()
    a = g.run(x, capi.G2_MASK_STATE)
    assert "/tma" in g.last_launch()["kernel"]
    pad = torch.zeros((2, 150, 263), dtype=torch.float32, device="cuda")
    pad[:, :, 1:261] = x
    b = g.run(pad[:, :, 1:261], capi.G2_MASK_STATE)   # misaligned base + pitch: TMA cannot describe it
    assert g.last_launch()["kernel"].endswith("/ldg")
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("mask_name", ["steer5", "full"])
def test_two_pixel_map_kernels_match_oracle_and_one_pixel_path(mask_name):
    """steer(theta map) through the two-pixel static kernels (angle pairs as 8-byte loads, one row ahead): same pixels as the
    one-pixel kernel that an odd-aligned view falls back to, bit for bit, and within tolerance of the oracle steered at the
    same angles (reference G2.cpp:167-177).  300 rows: interior bands, a ragged last band and both image borders."""
    fr = _frames(3310, 2, 300, 392)
    th = np.random.default_rng(3311).uniform(-4, 4, (2, 300, 392)).astype(np.float32)
    mask = capi.G2_MASK_STEER5 if mask_name == "steer5" else capi.G2_MASK_FULL
    g = G2Batch()
    x, t = torch.from_numpy(fr).cuda(), torch.from_numpy(th).cuda()
    a = g.run(x, mask, steer=capi.STEER_MAP, theta_map=t)
    assert g.last_launch()["kernel"].endswith("@map>/tma/imm-taps/2px"), g.last_launch()["kernel"]
    pad = torch.zeros((2, 300, 395), dtype=torch.float32, device="cuda")
    pad[:, :, 1:393] = x
    b = g.run(pad[:, :, 1:393], mask, steer=capi.STEER_MAP, theta_map=t)
    assert g.last_launch()["kernel"].endswith("/ldg")
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # an angle map at an address that is not a multiple of 8: the launcher must fall back to one pixel per thread
    tpad = torch.zeros(2 * 300 * 392 + 1, dtype=torch.float32, device="cuda")
    tpad[1:] = t.flatten()
    c = g.run(x, mask, steer=capi.STEER_MAP, theta_map=tpad[1:].view(2, 300, 392))
    assert g.last_launch()["kernel"].endswith("@map>/tma/imm-taps"), g.last_launch()["kernel"]
    for k in a:
        assert torch.equal(a[k], c[k]), k
    o = ref.SteerableFiltersG2(fr[1])
    rng = basis_range([getattr(o, k) for k in STATE])
    w = o.steer_map_full(th[1])
    for k, j, sc in (("g2", 0, rng), ("h2", 1, rng), ("e", 2, ("own", rng)), ("magnitude", 3, rng)):
        assert_close_range(a[k][1].cpu().numpy(), w[j], sc, "map/2px " + k)
    assert_angle_close(a["phase"][1].cpu().numpy(), w[4], w[3], 2 * np.pi, "map/2px phase")


def test_u8_input_matches_float_input():
    img8 = np.random.default_rng(4).integers(0, 256, (2, 100, 300), dtype=np.uint8)
    g = G2Batch()
    a = g.run(torch.from_numpy(img8).cuda(), capi.G2_MASK_FULL)
    assert "u8" in g.last_launch()["kernel"]
    b = g.run(torch.from_numpy(img8.astype(np.float32)).cuda(), capi.G2_MASK_FULL)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_frame_independence_and_batch_consistency():
    fr = _frames(3400, 5, 70, 140)
    g = G2Batch()
    full = g.run(torch.from_numpy(fr).cuda(), capi.G2_MASK_FULL)
    one = g.run(torch.from_numpy(fr[3:4]).cuda(), capi.G2_MASK_FULL)
    for k in full:
        assert torch.equal(full[k][3], one[k][0]), k


def test_band_equals_whole_bitwise():
    """Row-band launch (band + halo rows in the buffer) must reproduce the whole-image result exactly."""
    H, W = 300, 260
    img = synth(3500, H, W)
    g = G2Batch()
    whole = g.run(torch.from_numpy(img[None]).cuda(), capi.G2_MASK_FULL)
    for (r0, r1) in ((0, 100), (100, 228), (228, 300)):
        lo, hi = max(0, r0 - 4), min(H, r1 + 4)
        buf = torch.from_numpy(img[None, lo:hi].copy()).cuda()
        part = g.run(buf, capi.G2_MASK_FULL, band=Band(full_rows=H, y_origin=lo, row_begin=r0, row_end=r1))
        for k in whole:
            assert torch.equal(part[k][0], whole[k][0, r0:r1]), (k, r0, r1)


def test_flip_symmetry_and_scaling_4k():
    """Size-independent properties at full frame size (3840x2160): a horizontal flip mirrors even-x planes and
    negates odd-x planes; scaling the input by 2 scales the basis exactly by 2 (power-of-two, no rounding)."""
    rs = np.random.default_rng(123)
    img = rs.uniform(0, 255, (1, 2160, 3840)).astype(np.float32)
    g = G2Batch()
    x = torch.from_numpy(img).cuda()
    a = g.run(x, capi.G2_MASK_STATE)
    b = g.run(torch.flip(x, dims=[2]).contiguous(), capi.G2_MASK_STATE)
    # x-parity of each basis plane = parity of its row (kernelX) tap set: g1 e, g3 o, g2 e, h1 o, h4 e, h3 o, h2 e
    sign = {"g2a": 1, "g2b": -1, "g2c": 1, "h2a": -1, "h2b": 1, "h2c": -1, "h2d": 1}
    for k, s in sign.items():
        assert torch.equal(torch.flip(b[k], dims=[2]), s * a[k]), k
    c = g.run(x * 2, capi.G2_MASK_STATE)
    for k in sign:
        assert torch.equal(c[k], 2 * a[k]), k
    assert torch.equal(c["theta"], a["theta"])
    # spot parity on a crop against the oracle (crop far from borders => identical support)
    o = ref.SteerableFiltersG2(img[0, 1000:1200, 2000:2300])
    rng = basis_range([getattr(o, k) for k in STATE])
    for k in STATE:
        assert_close_range(a[k][0, 1010:1190, 2010:2290].cpu().numpy(), getattr(o, k)[10:-10, 10:-10], rng, "crop " + k)


def test_pyramid_levels_vs_oracle():
    fr = _frames(3600, 2, 135, 241)
    x = torch.from_numpy(fr).cuda()
    g = G2Batch()
    levels = g.run_pyramid(x, 5, capi.G2_MASK_ORIENT)
    assert [tuple(l["theta"].shape[1:]) for l in levels] == [(135, 241), (68, 121), (34, 61), (17, 31), (9, 16)]
    for i in range(2):
        pyr = ref.pyramid(fr[i], 5)
        cur = x[i:i + 1]
        for l in range(5):
            if l:
                cur = pyr_down(cur)
                assert float(np.max(np.abs(cur[0].cpu().numpy() - pyr[l]))) <= 1e-4 * 255, f"pyr level {l}"
            # each level's analysis vs the oracle run on the ORACLE's level (tolerances absorb the 1e-5 level diff)
            o = ref.SteerableFiltersG2(pyr[l])
            rng = basis_range([getattr(o, k) for k in STATE])
            assert_close_range(levels[l]["strength"][i].cpu().numpy(), o.strength, ("own", rng), f"L{l} strength")
            assert_angle_close(levels[l]["theta"][i].cpu().numpy(), o.theta, o.strength, np.pi, f"L{l} theta", thresh_frac=1e-2)


@pytest.mark.parametrize("shape", [(1, 7), (3, 3), (2, 9), (37, 51), (64, 64)])
def test_pyr_down_edge_sizes(shape):
    img = synth(3700, *shape)
    out = pyr_down(torch.from_numpy(img[None]).cuda())[0].cpu().numpy()
    want = ref.pyr_down(img)
    assert out.shape == want.shape
    assert float(np.max(np.abs(out - want))) <= 1e-4 * 255


def test_band_pyramid_equals_whole_bitwise():
    """The row-band pipeline of cvsteer_b200.multi (bands + halos through a 4-level pyramid, band-mode pyr_down and
    band-mode fused kernel) reproduces the whole-image pyramid result exactly.  Ranks are emulated sequentially."""
    from cvsteer_b200 import multi
    H, W, L = 400, 300, 4
    img = synth(3800, H, W)
    x = torch.from_numpy(img[None]).cuda()
    g = G2Batch()
    whole = g.run_pyramid(x, L, capi.G2_MASK_FULL)
    process, down, _ = multi.cuda_callables(capi.G2_MASK_FULL)
    for world in (2, 3):
        for plan in multi.plan_bands(H, world, L):
            if plan.empty:
                continue
            buf = x[0, plan.have[0][0]:plan.have[0][1]].contiguous()
            for l in range(L):
                res = process(buf, l, plan)
                lo, hi = plan.out[l]
                for k, v in res.items():
                    assert torch.equal(v, whole[l][k][0, lo:hi]), (world, plan.rank, l, k)
                if l + 1 < L:
                    buf = down(buf, l, plan)


def test_host_batch_api_matches_device_path():
    """cvs_g2_run_batch_host (the e2e call of bench.py): chunked H2D / kernel / D2H pipeline == one device launch."""
    fr = _frames(3900, 7, 120, 200)                      # 7 frames: uneven chunks
    g = G2Batch()
    dev = g.run(torch.from_numpy(fr).cuda(), capi.G2_MASK_FULL)
    xh = torch.from_numpy(fr).pin_memory()
    planes = [p for p in range(capi.G2_NPLANES) if capi.G2_MASK_FULL >> p & 1]
    oh = {p: torch.empty((7, 120, 200), dtype=torch.float32).pin_memory() for p in planes}
    g.run_host(xh, capi.G2_MASK_FULL, oh)
    for p in planes:
        assert torch.equal(oh[p], dev[capi.G2_PLANE_NAMES[p]].cpu()), capi.G2_PLANE_NAMES[p]
    # pageable host memory works too (slower)
    oh2 = {p: torch.empty((7, 120, 200), dtype=torch.float32) for p in planes}
    g.run_host(torch.from_numpy(fr), capi.G2_MASK_FULL, oh2)
    assert torch.equal(oh2[capi.PHASE], oh[capi.PHASE])


def test_host_batch_dense_rows_take_linear_copies_and_match():
    """Rows that are a multiple of 128 bytes (1920 and 3840 columns are) make frames dense on the device: the pipeline
    then moves each chunk with ONE linear copy per plane.  Dense input + dense outputs, dense input + strided outputs and
    strided input + dense outputs must all equal the device launch bit for bit."""
    n, R, Cc = 5, 96, 256
    fr = _frames(3901, n, R, Cc)
    g = G2Batch()
    dev = g.run(torch.from_numpy(fr).cuda(), capi.G2_MASK_FULL)
    planes = [p for p in range(capi.G2_NPLANES) if capi.G2_MASK_FULL >> p & 1]
    dense_in = torch.from_numpy(fr).pin_memory()
    wide_in = torch.zeros((n, R + 3, Cc + 32), dtype=torch.float32).pin_memory()
    wide_in[:, :R, :Cc] = dense_in
    for xin, strided_out in ((dense_in, False), (dense_in, True), (wide_in[:, :R, :Cc], False)):
        if strided_out:
            big = {p: torch.full((n, R + 2, Cc + 64), -7.0, dtype=torch.float32).pin_memory() for p in planes}
            oh = {p: big[p][:, :R, :Cc] for p in planes}
        else:
            oh = {p: torch.empty((n, R, Cc), dtype=torch.float32).pin_memory() for p in planes}
        g.run_host(xin, capi.G2_MASK_FULL, oh)
        for p in planes:
            assert torch.equal(oh[p], dev[capi.G2_PLANE_NAMES[p]].cpu()), (capi.G2_PLANE_NAMES[p], strided_out)
        if strided_out:  # nothing outside the views was touched
            assert all(float(big[p][:, R:, :].max()) == -7.0 and float(big[p][:, :, Cc:].max()) == -7.0 for p in planes)


def test_batch_api_errors():
    import ctypes as C
    g = G2Batch()
    x = torch.zeros((1, 16, 16), device="cuda")
    with pytest.raises(capi.CvsError) as e:
        g.run(x, 0)
    assert e.value.code == capi.ERR_INVALID_ARG
    with pytest.raises(capi.CvsError):
        g.run(x, 1 << 25)
    with pytest.raises(capi.CvsError):
        g.run(x, capi.G2_MASK_FULL, steer=capi.STEER_MAP)          # no theta map
    with pytest.raises(capi.CvsError):
        g.run(torch.zeros((1, 16, 16)), capi.G2_MASK_FULL)          # host tensor on the device path
    b = capi.Batch()
    assert capi.lib().cvs_g2_run_batch_dev(g._h, C.byref(b), 1, 0, 0.0, None, None, None) == capi.ERR_INVALID_ARG
    assert b"null" in capi.lib().cvs_last_error()


def test_multi_device_host_api():
    """cvs_g2_run_batch_host_multi: frames sharded over every visible GPU from ONE process (threads, no collective)."""
    import ctypes as C
    ndev = torch.cuda.device_count()
    fr = _frames(3950, 5, 90, 130)
    want = G2Batch().run(torch.from_numpy(fr).cuda(), capi.G2_MASK_ORIENT)
    planes = [p for p in range(capi.G2_NPLANES) if capi.G2_MASK_ORIENT >> p & 1]
    outs = {p: np.empty((5, 90, 130), np.float32) for p in planes}
    arr = (C.c_void_p * capi.G2_NPLANES)()
    for p in planes:
        arr[p] = outs[p].ctypes.data
    for nd in sorted({1, ndev}):
        for o in outs.values():
            o.fill(0)
        capi.check(capi.lib().cvs_g2_run_batch_host_multi(nd, None, 4, 0.67, fr.ctypes.data, 5, 90, 130, 130 * 4, 90 * 130 * 4,
                                                          capi.G2_MASK_ORIENT, arr, 130 * 4, 90 * 130 * 4))
        for p in planes:
            assert np.array_equal(outs[p], want[capi.G2_PLANE_NAMES[p]].cpu().numpy()), (nd, p)


def test_generic_width_host_batch_pipeline():
    """Non-default widths go through the two-pass generic kernels and a scratch buffer.  The host-batch pipeline keeps
    three chunks in flight on three streams: every stream needs its own scratch slice (was one shared buffer: a race)."""
    fr = _frames(3970, 9, 150, 230)                      # 9 frames -> 3-frame chunks round-robin on 3 streams
    g = G2Batch(width=6, spacing=0.45)
    dev = g.run(torch.from_numpy(fr).cuda(), capi.G2_MASK_FULL)
    assert g.last_launch()["kernel"].startswith("generic")
    planes = [p for p in range(capi.G2_NPLANES) if capi.G2_MASK_FULL >> p & 1]
    xh = torch.from_numpy(fr).pin_memory()
    for _ in range(3):                                   # a race does not show every time
        oh = {p: torch.zeros((9, 150, 230), dtype=torch.float32).pin_memory() for p in planes}
        g.run_host(xh, capi.G2_MASK_FULL, oh)
        for p in planes:
            assert torch.equal(oh[p], dev[capi.G2_PLANE_NAMES[p]].cpu()), capi.G2_PLANE_NAMES[p]


def test_generic_width_band_equals_whole():
    H, W = 150, 170
    img = synth(3960, H, W)
    g = G2Batch(width=6, spacing=0.45)
    whole = g.run(torch.from_numpy(img[None]).cuda(), capi.G2_MASK_FULL)
    assert g.last_launch()["kernel"].startswith("generic")
    for (r0, r1) in ((0, 70), (70, 150)):
        lo, hi = max(0, r0 - 6), min(H, r1 + 6)
        part = g.run(torch.from_numpy(img[None, lo:hi].copy()).cuda(), capi.G2_MASK_FULL, band=Band(full_rows=H, y_origin=lo, row_begin=r0, row_end=r1))
        for k in whole:
            assert torch.equal(part[k][0], whole[k][0, r0:r1]), (k, r0)


def test_huge_single_image_64bit_indexing():
    """One 32768 x 20000 frame (2.6 GB in, 7.9 GB out): plane offsets exceed 2^32 bytes; crops are checked against the
    oracle run on the same crops (far from borders the supports coincide)."""
    H, W = 32768, 20000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    x = torch.rand((1, H, W), device="cuda", generator=gen) * 255
    g = G2Batch()
    r = g.run(x, capi.G2_MASK_ORIENT)
    for (y0, x0) in ((100, 100), (32768 - 300, 20000 - 400), (30000, 123), (16384, 9000)):
        crop = x[0, y0:y0 + 160, x0:x0 + 200].cpu().numpy()
        o = ref.SteerableFiltersG2(crop)
        rng = basis_range([getattr(o, k) for k in STATE])
        got = r["strength"][0, y0 + 8:y0 + 152, x0 + 8:x0 + 192].cpu().numpy()
        assert_close_range(got, o.strength[8:-8, 8:-8], ("own", rng), f"strength crop {(y0, x0)}")
        th = r["theta"][0, y0 + 8:y0 + 152, x0 + 8:x0 + 192].cpu().numpy()
        assert_angle_close(th, o.theta[8:-8, 8:-8], o.strength[8:-8, 8:-8], np.pi, f"theta crop {(y0, x0)}")
    # bottom border rows use reflect-101 of the true last rows
    o = ref.SteerableFiltersG2(x[0, H - 120:, 5000:5200].cpu().numpy())
    rng = basis_range([getattr(o, k) for k in STATE])
    assert_close_range(r["strength"][0, H - 100:, 5008:5192].cpu().numpy(), o.strength[20:, 8:-8], ("own", rng), "bottom border")
    del r, x
    torch.cuda.empty_cache()


@pytest.mark.parametrize("shape", [(135, 241), (64, 128), (200, 383), (1080, 1920), (37, 5)])
def test_fused_pyramid_emission_bitwise(shape):
    """next_level written by the fused basis kernel == the stand-alone pyr_down kernel, bit for bit; and the pyramid
    results do not depend on which of the two builds the levels."""
    fr = _frames(3990, 2, *shape)
    x = torch.from_numpy(fr).cuda()
    g = G2Batch()
    nxt = torch.empty((2, (shape[0] + 1) // 2, (shape[1] + 1) // 2), device="cuda")
    for mask in (capi.G2_MASK_ORIENT, capi.G2_MASK_STATE, (1 << capi.G2_NPLANES) - 1):
        nxt.fill_(-1)
        g.run(x, mask, next_level=nxt)
        assert torch.equal(nxt, pyr_down(x)), (shape, hex(mask))
    assert float(np.max(np.abs(nxt[1].cpu().numpy() - ref.pyr_down(fr[1])))) <= 1e-4 * 255
    a = g.run_pyramid(x, 4, capi.G2_MASK_ORIENT)
    b = g.run_pyramid(x, 4, capi.G2_MASK_ORIENT, fuse_pyramid=False)
    for la, lb in zip(a, b):
        for k in la:
            assert torch.equal(la[k], lb[k]), k
    # misaligned input -> LDG loader; generic width -> stand-alone kernel behind the same argument
    pad = torch.zeros((2, shape[0], shape[1] + 3), device="cuda")
    pad[:, :, 1:shape[1] + 1] = x
    nxt.fill_(-1)
    g.run(pad[:, :, 1:shape[1] + 1], capi.G2_MASK_ORIENT, next_level=nxt)
    assert torch.equal(nxt, pyr_down(x))
    nxt.fill_(-1)
    G2Batch(width=5, spacing=0.5).run(x, capi.G2_MASK_ORIENT, next_level=nxt)
    assert torch.equal(nxt, pyr_down(x))


def test_bands_host_multi_equals_whole():
    """cvs_g2_run_bands_host_multi: one image, row bands over every visible GPU (and over 3 'virtual' bands on one GPU by
    listing device 0 three times), 4-level pyramid, results identical to the whole-image pyramid."""
    import ctypes as C
    H, W, L = 333, 270, 4
    img = synth(4001, H, W)
    want = G2Batch().run_pyramid(torch.from_numpy(img[None]).cuda(), L, capi.G2_MASK_ORIENT)
    planes = [p for p in range(capi.G2_NPLANES) if capi.G2_MASK_ORIENT >> p & 1]
    shapes = [(H, W)]
    for _ in range(L - 1):
        shapes.append(((shapes[-1][0] + 1) // 2, (shapes[-1][1] + 1) // 2))
    for devs in ([0, 0, 0], list(range(torch.cuda.device_count()))):
        outs = [{p: np.zeros(s, np.float32) for p in planes} for s in shapes]
        lvl = (C.POINTER(C.c_void_p) * L)()
        keep = []
        for l in range(L):
            arr = (C.c_void_p * capi.G2_NPLANES)()
            for p in planes:
                arr[p] = outs[l][p].ctypes.data
            keep.append(arr)
            lvl[l] = C.cast(arr, C.POINTER(C.c_void_p))
        steps = (C.c_size_t * L)(*[s[1] * 4 for s in shapes])
        dv = (C.c_int * len(devs))(*devs)
        capi.check(capi.lib().cvs_g2_run_bands_host_multi(len(devs), dv, 4, 0.67, img.ctypes.data, H, W, W * 4, L, capi.G2_MASK_ORIENT, lvl, steps))
        for l in range(L):
            for p in planes:
                assert np.array_equal(outs[l][p], want[l][capi.G2_PLANE_NAMES[p]][0].cpu().numpy()), (devs, l, p)


def test_more_than_65535_frames():
    """gridDim.z limit: 70000 tiny frames go out as two launches and still match per-frame results."""
    n = 70000
    x = torch.rand((n, 6, 8), device="cuda") * 255
    g = G2Batch()
    r = g.run(x, capi.G2_MASK_ORIENT)
    for i in (0, 65534, 65535, 65536, n - 1):
        one = g.run(x[i:i + 1].contiguous(), capi.G2_MASK_ORIENT)
        for k in r:
            assert torch.equal(r[k][i], one[k][0]), (k, i)


def test_u8_inputs_on_secondary_paths():
    """8-bit input through the stand-alone pyr_down kernel and through the generic-width path."""
    img8 = np.random.default_rng(6).integers(0, 256, (2, 77, 130), dtype=np.uint8)
    x8, xf = torch.from_numpy(img8).cuda(), torch.from_numpy(img8.astype(np.float32)).cuda()
    assert torch.equal(pyr_down(x8), pyr_down(xf))
    g = G2Batch(width=3, spacing=0.8)
    a, b = g.run(x8, capi.G2_MASK_FULL), g.run(xf, capi.G2_MASK_FULL)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # non-contiguous (strided) float input view: rows with a pitch larger than cols*4
    big = torch.rand((2, 77, 200), device="cuda") * 255
    view = big[:, :, 10:140]
    r1 = G2Batch().run(view, capi.G2_MASK_ORIENT)
    r2 = G2Batch().run(view.contiguous(), capi.G2_MASK_ORIENT)
    for k in r1:
        assert torch.equal(r1[k], r2[k]), k


def test_fuzz_shapes_masks_vs_oracle():
    """Seeded fuzz over image shapes around every tiling boundary (strip width 128, band height 64, row groups of 9),
    output masks and steering modes; every selected plane is compared with the oracle."""
    rs = np.random.default_rng(20261017)
    edge_rows = [1, 2, 4, 5, 8, 9, 10, 55, 56, 57, 63, 64, 65, 72, 73, 127, 128, 129, 136]
    edge_cols = [1, 3, 4, 5, 9, 120, 124, 127, 128, 129, 132, 133, 255, 256, 257, 260, 391]
    g = G2Batch()
    for it in range(28):
        rows = int(rs.choice(edge_rows)) if it % 3 else int(rs.integers(1, 200))
        cols = int(rs.choice(edge_cols)) if it % 2 else int(rs.integers(1, 420))
        n = int(rs.integers(1, 4))
        fr = np.stack([synth(9000 + 10 * it + i, rows, cols) for i in range(n)])
        mode = it % 4
        x = torch.from_numpy(fr).cuda()
        if mode == 0:
            mask, kw = capi.G2_MASK_FULL, {}
        elif mode == 1:
            mask, kw = capi.G2_MASK_STATE, {}
        elif mode == 2:
            mask = capi.G2_MASK_ORIENT | capi.bit(capi.G2A) | capi.bit(capi.H2D) | capi.bit(capi.G2T) | capi.bit(capi.PHASE) | capi.bit(capi.MAG)
            kw = {}
        else:
            mask = capi.bit(capi.G2T) | capi.bit(capi.H2T) | capi.bit(capi.E) | capi.bit(capi.MAG)
            kw = dict(steer=capi.STEER_SCALAR, theta=float(rs.uniform(-3, 3)))
        r = g.run(x, mask, **kw)
        for i in range(n):
            o = ref.SteerableFiltersG2(fr[i])
            rng = max(basis_range([getattr(o, k) for k in STATE]), 1.0)
            tag = f"it{it} {rows}x{cols} f{i} "
            for k in STATE + ("c1", "c2", "c3", "strength"):
                if k in r:
                    assert_close_range(r[k][i].cpu().numpy(), getattr(o, k), rng if k in STATE else ("own", rng), tag + k)
            if "theta" in r:
                assert_angle_close(r["theta"][i].cpu().numpy(), o.theta, o.strength, np.pi, tag + "theta")
            if mode == 3:
                w = o.steer_scalar_full(kw["theta"])
            elif "g2" in r or "phase" in r or "magnitude" in r:
                w = o.steer_map_full(r["theta"][i].cpu().numpy())
            else:
                continue
            for k, j, s in (("g2", 0, rng), ("h2", 1, rng), ("e", 2, ("own", rng)), ("magnitude", 3, rng)):
                if k in r and not (k == "e" and mode == 1):
                    assert_close_range(r[k][i].cpu().numpy(), w[j], s, tag + k)
            if "phase" in r:
                assert_angle_close(r["phase"][i].cpu().numpy(), w[4], w[3], 2 * np.pi, tag + "phase")


@pytest.mark.parametrize("seed", range(24))
def test_random_sweep_vs_oracle(seed):
    """Seeded sweep over what a caller can vary at once: frame size (incl. smaller than the filter and not a multiple of the
    tile), batch size, 8-bit or float input, a padded / misaligned input view, and a random plane mask."""
    r = np.random.default_rng(9000 + seed)
    rows = int(r.choice([r.integers(1, 12), r.integers(12, 140), r.integers(140, 330)]))
    cols = int(r.choice([r.integers(1, 12), r.integers(12, 140), r.integers(140, 420)]))
    n = int(r.integers(1, 4))
    u8 = bool(r.integers(0, 2))
    if u8:
        fr = r.integers(0, 256, (n, rows, cols), dtype=np.uint8)
    else:
        fr = r.uniform(0, 255, (n, rows, cols)).astype(np.float32)
    x = torch.from_numpy(fr).cuda()
    if r.integers(0, 2):                                   # non-dense view: odd column offset and a longer pitch
        off, extra = int(r.integers(0, 4)), int(r.integers(0, 9))
        buf = torch.zeros((n, rows + 2, cols + off + extra), dtype=x.dtype, device="cuda")
        buf[:, 1:rows + 1, off:off + cols] = x
        x = buf[:, 1:rows + 1, off:off + cols]
    mask = int(r.integers(1, 1 << capi.G2_NPLANES))
    if r.integers(0, 3) == 0:
        mask = int(r.choice([capi.G2_MASK_STATE, capi.G2_MASK_ORIENT, capi.G2_MASK_FULL]))
    g = G2Batch()
    res = g.run(x, mask)
    full = g.run(x, (1 << capi.G2_NPLANES) - 1)              # every plane, same launch geometry
    names = [capi.G2_PLANE_NAMES[p] for p in range(capi.G2_NPLANES) if mask >> p & 1]
    assert sorted(res) == sorted(names)
    for k in names:                                        # a mask selects planes, it never changes their values...
        a, b = res[k], full[k]
        if k in ("g2", "h2", "e", "magnitude", "phase", "edges", "lines_dark", "lines_bright", "theta", "strength") and \
                mask in (capi.G2_MASK_STATE, capi.G2_MASK_ORIENT, capi.G2_MASK_FULL):
            continue                                       # ...except that the static fast-math masks round differently
        assert torch.equal(a, b), (k, hex(mask))
    for i in range(n):                                     # and the values are the oracle's
        o = ref.SteerableFiltersG2(fr[i].astype(np.float32))
        rng = max(basis_range([getattr(o, k) for k in STATE]), 1e-3)
        for k in STATE:
            if k in res:
                assert_close_range(res[k][i].cpu().numpy(), getattr(o, k), rng, f"{k} seed{seed}")
        for k in ("c1", "c2", "c3", "strength"):
            if k in res:
                assert_close_range(res[k][i].cpu().numpy(), getattr(o, k), ("own", rng), f"{k} seed{seed}")
        if "theta" in res:
            assert_angle_close(res["theta"][i].cpu().numpy(), o.theta, o.strength, np.pi, f"theta seed{seed}")
        th = full["theta"][i].cpu().numpy()
        w = o.steer_map_full(th)
        for k, j, s in (("g2", 0, rng), ("h2", 1, rng), ("e", 2, ("own", rng)), ("magnitude", 3, rng)):
            if k in res:
                assert_close_range(res[k][i].cpu().numpy(), w[j], s, f"{k} seed{seed}")
        if "phase" in res:
            assert_angle_close(res["phase"][i].cpu().numpy(), w[4], w[3], 2 * np.pi, f"phase seed{seed}")


def test_u8_tma_loader_matches_float_and_ldg_bitwise():
    """8-bit frames TMA can describe (16-byte aligned base and pitch) are staged as bytes and expanded in shared memory:
    same results, bit for bit, as the fp32 input and as the cooperative 8-bit loader; G4 too."""
    from cvsteer_b200.batch import G4Batch
    img8 = np.random.default_rng(41).integers(0, 256, (3, 150, 256), dtype=np.uint8)
    x8 = torch.from_numpy(img8).cuda()
    xf = x8.float()
    pad = torch.zeros((3, 150, 259), dtype=torch.uint8, device="cuda")
    pad[:, :, 1:257] = x8
    for g, mask in ((G2Batch(), capi.G2_MASK_FULL), (G2Batch(), capi.G2_MASK_LINES), (G2Batch(), (1 << capi.G2_NPLANES) - 1),
                    (G4Batch(), capi.G4_MASK_BASIS)):
        a = g.run(x8, mask)
        assert g.last_launch()["kernel"].endswith("/tma-u8"), g.last_launch()["kernel"]
        b = g.run(xf, mask)
        c = g.run(pad[:, :, 1:257], mask)
        assert g.last_launch()["kernel"].endswith("/ldg-u8")
        for k in a:
            assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), (k, hex(mask))


@pytest.mark.parametrize("shape", [(360, 640), (185, 256)])
def test_fast_static_path_end_to_end_vs_oracle_own_theta(shape, fish_fixture):
    """The fused static M2 kernel (MUFU rcp / sqrt / sin / cos, its OWN theta_d) against the oracle's full flow steered at
    the ORACLE's theta_d -- no shared angles.  Pixels are compared where the comparison is well posed (SURVEY section 7,
    hard part 2c): away from the branch cut |theta_d| = pi/2, where a last-bit difference flips the sign of theta_d and
    with it H2 and the phase, and where the orientation is defined (strength above 1 % of its maximum: below that
    theta_d = atan2(c3, c2) / 2 of two rounding-level numbers, and d(g2, h2)/d(theta) x its error exceeds any tolerance
    for ANY two implementations, two OpenCV builds included)."""
    img = fish_fixture["fish"].astype(np.float32) if shape == (185, 256) else synth(8100, *shape)
    o, (g2, h2, e, mag, ph) = ref.g2_full(img)
    rng = basis_range([getattr(o, k) for k in STATE])
    g = G2Batch()
    r = g.run(torch.from_numpy(img[None]).cuda(), capi.G2_MASK_FULL)
    assert "M2" in g.last_launch()["kernel"]
    ok = (np.abs(o.theta) < np.pi / 2 - 1e-3) & (o.strength > 1e-2 * float(o.strength.max()))
    assert ok.mean() > 0.2, float(ok.mean())          # (the fish image is mostly flat background)
    got = {k: r[k][0].cpu().numpy() for k in ("theta", "strength", "g2", "h2", "e", "magnitude", "phase")}
    assert_angle_close(got["theta"], o.theta, o.strength, np.pi, "e2e theta_d")
    assert_close_range(got["strength"], o.strength, ("own", rng), "e2e strength")
    assert_close_range(got["e"], e, ("own", rng), "e2e e")            # e = c1 + strength: no angle sensitivity, every pixel
    for name, want in (("g2", g2), ("h2", h2), ("magnitude", mag)):
        assert_close_range(got[name][ok], want[ok], rng, "e2e " + name)
    assert_angle_close(got["phase"][ok], ph[ok], mag[ok], 2 * np.pi, "e2e phase")
    # and at the cut itself the only disagreement is the documented symmetry: theta -> -theta, H2 -> -H2, phase -> -phase
    cut = (np.abs(o.theta) >= np.pi / 2 - 1e-3) & (o.strength > 1e-2 * float(o.strength.max()))
    if cut.any():
        flip = np.sign(got["theta"][cut]) != np.sign(o.theta[cut])
        assert_close_range(np.abs(got["h2"][cut]), np.abs(h2[cut]), rng, "e2e |h2| at the cut")
        assert_close_range(got["g2"][cut], g2[cut], rng, "e2e g2 at the cut")
        assert_close_range(got["h2"][cut][~flip], h2[cut][~flip], rng, "e2e h2 at the cut, same sign")
