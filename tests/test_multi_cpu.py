"""Host-side logic of the multi-GPU dispatch, on CPU with the gloo backend (world_size 2 and 3): frame sharding,
band planning (halo sufficiency through a 5-level pyramid) and the batched send/recv gather.  The compute callables
are simple numpy stand-ins with the SAME support as the real kernels (radius-4 vertical stencil, 5-tap even-sample
downsample, reflect-101 at true image borders only) and they assert that every row they touch is present in the band
buffer -- so a wrong halo computation fails loudly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cvsteer_b200 import multi


def _reflect(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * (n - 1) - p
    return p


def _stencil_rows(buf, origin, full_rows, lo, hi, radius):
    """out[y] = sum_{i=-r..r} (i+r+1) * img[reflect(y+i)] for y in [lo,hi), reading rows from buf (rows origin..)."""
    out = np.zeros((hi - lo, buf.shape[1]), np.float64)
    for y in range(lo, hi):
        for i in range(-radius, radius + 1):
            r = _reflect(y + i, full_rows) - origin
            assert 0 <= r < buf.shape[0], f"row {y + i} of level not in band buffer [{origin},{origin + buf.shape[0]})"
            out[y - lo] += (i + radius + 1) * buf[r]
    return out


def _down_rows(buf, origin, full_rows, lo, hi):
    w = np.array([1, 4, 6, 4, 1], np.float64) / 16
    cols = buf.shape[1]
    oc = (cols + 1) // 2
    out = np.zeros((hi - lo, oc), np.float64)
    xs = [[_reflect(2 * x + j - 2, cols) for j in range(5)] for x in range(oc)]
    for y in range(lo, hi):
        for i in range(5):
            r = _reflect(2 * y + i - 2, full_rows) - origin
            assert 0 <= r < buf.shape[0], "pyr_down source row missing from band buffer"
            row = buf[r]
            out[y - lo] += w[i] * np.array([sum(w[j] * row[xs[x][j]] for j in range(5)) for x in range(oc)])
    return out


def _whole(img, levels):
    res, cur = [], img.astype(np.float64)
    for l in range(levels):
        res.append(_stencil_rows(cur, 0, cur.shape[0], 0, cur.shape[0], 4))
        if l + 1 < levels:
            cur = _down_rows(cur, 0, cur.shape[0], 0, (cur.shape[0] + 1) // 2)
    return res


def _worker(rank, world, port, rows, cols, levels, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = np.random.default_rng(7).uniform(0, 255, (rows, cols))

    def load(lo, hi):
        return torch.from_numpy(img[lo:hi].copy())

    def process(buf, level, plan):
        lo, hi = plan.out[level]
        o = _stencil_rows(buf.numpy(), plan.have[level][0], plan.rows[level], lo, hi, 4)
        return {"a": torch.from_numpy(o.astype(np.float32)), "b": torch.from_numpy((2 * o).astype(np.float32))}

    def down(buf, level, plan):
        a, b = plan.have[level + 1]
        return torch.from_numpy(_down_rows(buf.numpy(), plan.have[level][0], plan.rows[level], a, b))

    full, local, plan = multi.run_bands(load, rows, cols, levels, process, down, root=0)
    if rank == 0:
        want = _whole(img, levels)
        ok = all(np.array_equal(full[l]["a"].numpy(), want[l].astype(np.float32)) and
                 np.array_equal(full[l]["b"].numpy(), (2 * want[l]).astype(np.float32)) for l in range(levels))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,rows,cols,levels", [(2, 96, 21, 3), (3, 150, 17, 4), (2, 40, 9, 5)])
def test_band_pipeline_gloo(world, rows, cols, levels):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, cols, levels, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_plan_properties():
    for rows, world, levels in ((32768, 8, 5), (300, 2, 5), (1000, 3, 4), (17, 4, 3), (2160, 8, 5)):
        plans = multi.plan_bands(rows, world, levels)
        hl = multi.level_rows(rows, levels)
        for l in range(levels):
            cover = []
            for p in plans:
                lo, hi = p.out[l]
                cover += list(range(lo, hi))
                if lo < hi:  # the kernel's support is inside what the rank holds
                    assert p.have[l][0] <= max(0, lo - 4) and p.have[l][1] >= min(hl[l], hi + 4)
                if l + 1 < levels and p.have[l + 1][0] < p.have[l + 1][1]:
                    a, b = p.have[l + 1]
                    assert p.have[l][0] <= max(0, 2 * a - 2) and p.have[l][1] >= min(hl[l], 2 * b + 1)
            assert cover == list(range(hl[l])), (rows, world, l)   # bands tile every level exactly once
    p = multi.plan_bands(32768, 8, 5)[1]
    assert p.out[0] == (4096, 8192) and p.have[0] == (4096 - 94, 8192 + 79)


def test_shard_frames():
    for n, w in ((256, 8), (64, 1), (5, 4), (3, 8)):
        got = [multi.shard_frames(n, w, r) for r in range(w)]
        flat = [i for lo, hi in got for i in range(lo, hi)]
        assert flat == list(range(n))
        assert max(hi - lo for lo, hi in got) == -(-n // w)


def test_plane_layout_for_peer_planes():
    """Flat buffer behind PeerPlanes: every (level, plane) present once, 128-byte aligned, no overlap, level sizes halve."""
    names = ["theta", "strength", "e"]
    rows = multi.level_rows(1001, 5)
    lay, total = multi.plane_layout(names, rows, 701)
    assert set(lay) == {(l, n) for l in range(5) for n in names}
    spans = sorted((off, off + r * c) for off, r, c in lay.values())
    assert all(off % 32 == 0 for off, _ in spans)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total
    assert lay[(0, "theta")][1:] == (1001, 701) and lay[(1, "e")][1:] == (501, 351) and lay[(4, "e")][1:] == (63, 44)


def test_c_band_plan_equals_python_plan():
    """cvs_plan_bands (what cvs_g2_run_bands_host_multi shards by, exported for C/C++ callers) == multi.plan_bands."""
    import ctypes as C

    from cvsteer_b200 import capi
    lib = capi.lib()
    n = 0
    for rows in (5, 17, 64, 100, 777, 1000, 4097, 32768):
        for world in (1, 2, 3, 8):
            for levels in (1, 2, 4, 5):
                for radius in (4, 6):
                    plan = (C.c_int * (world * levels * 4))()
                    hl = (C.c_int * levels)()
                    capi.check(lib.cvs_plan_bands(rows, world, levels, radius, plan, hl))
                    assert list(hl) == multi.level_rows(rows, levels)
                    for p in multi.plan_bands(rows, world, levels, radius):
                        for l in range(levels):
                            q = plan[(p.rank * levels + l) * 4:(p.rank * levels + l) * 4 + 4]
                            for got, want in ((tuple(q[:2]), p.out[l]), (tuple(q[2:]), p.have[l])):
                                if want[0] >= want[1]:
                                    assert got[0] >= got[1], (rows, world, levels, radius, p.rank, l, got, want)
                                else:
                                    assert got == tuple(want), (rows, world, levels, radius, p.rank, l, got, want)
                            n += 1
    assert n > 1000
    assert lib.cvs_plan_bands(0, 1, 1, 4, plan, None) == capi.ERR_INVALID_ARG
