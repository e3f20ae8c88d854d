"""Row-band mode behind the C ABI (cvs_bands_*, cvs_g2_run_bands_dev_multi): every gather mode must reproduce the single-GPU
whole-image pyramid BIT FOR BIT.  Single-GPU boxes emulate the ranks on one device (several contexts in one process, or two
processes sharing GPU 0 over CUDA IPC); the NCCL mode needs two GPUs and is skipped otherwise."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

from cvsteer_b200 import capi
from cvsteer_b200.bands import BandRun, _view
from cvsteer_b200.batch import G2Batch
from tests.util import synth

pytestmark = pytest.mark.gpu
L = 5
NAMES = ("theta", "strength", "e")


def _whole(img):
    return G2Batch().run_pyramid(torch.from_numpy(img[None]).cuda(), L, capi.G2_MASK_ORIENT)


@pytest.mark.parametrize("world,mode,H,W", [(1, capi.GATHER_NONE, 1000, 700), (3, capi.GATHER_PEER_STORE, 1000, 700),
                                            (3, capi.GATHER_PEER_COPY, 1000, 700), (5, capi.GATHER_PEER_STORE, 1000, 700),
                                            (2, capi.GATHER_PEER_COPY, 2300, 520)])   # 1152-row bands: copied in 3 chunks
def test_contexts_of_one_process_equal_whole_image(world, mode, H, W):
    """`world` band contexts on ONE device in ONE process: the root exports its block, the others attach the pointer."""
    img = synth(7100, H, W)
    whole = _whole(img)
    lib = capi.lib()
    ctx = []
    for r in range(world):
        h = C.c_void_p()
        capi.check(lib.cvs_bands_create(C.byref(h), 0, r, world, 0, H, W, L, capi.G2_MASK_ORIENT, 4, 0.67))
        ctx.append(h)
    base = C.c_void_p()
    capi.check(lib.cvs_bands_root_export(ctx[0], None, C.byref(base)))
    for h in ctx[1:]:
        capi.check(lib.cvs_bands_root_attach(h, base))
    himg = torch.from_numpy(img)
    s = torch.cuda.current_stream().cuda_stream
    for h in ctx:
        capi.check(lib.cvs_bands_upload_host(h, himg.data_ptr(), W * 4, C.c_void_p(s)))
    for _ in range(2):                                   # reusable across steps
        for h in ctx:
            capi.check(lib.cvs_bands_run(h, mode, C.c_void_p(s)))
    torch.cuda.synchronize()
    for l in range(L):
        rows, cols = whole[l]["theta"].shape[1:]
        for p, name in ((capi.THETA, "theta"), (capi.STRENGTH, "strength"), (capi.E, "e")):
            ptr, pitch = C.c_void_p(), C.c_size_t()
            capi.check(lib.cvs_bands_root_plane(ctx[0], l, p, C.byref(ptr), C.byref(pitch)))
            got = _view(ptr.value, rows, cols, pitch.value, 0)
            assert torch.equal(got, whole[l][name][0]), (world, mode, l, name)
    for h in reversed(ctx):
        lib.cvs_bands_destroy(h)


def test_geometry_matches_python_planner():
    from cvsteer_b200 import multi
    lib = capi.lib()
    for rows, world in ((1000, 3), (32768, 8), (77, 4), (16, 8)):
        plans = multi.plan_bands(rows, world, L)
        for r in range(world):
            h = C.c_void_p()
            capi.check(lib.cvs_bands_create(C.byref(h), 0, r, world, 0, rows, 64, L, capi.G2_MASK_ORIENT, 4, 0.67))
            for l in range(L):
                v = [C.c_int() for _ in range(6)]
                capi.check(lib.cvs_bands_geometry(h, -1, l, *[C.byref(x) for x in v], None))
                assert (v[2].value, v[3].value) == plans[r].out[l] and (v[4].value, v[5].value) == plans[r].have[l], (rows, world, r, l)
            lib.cvs_bands_destroy(h)


@pytest.mark.parametrize("mode", [capi.GATHER_PEER_STORE, capi.GATHER_PEER_COPY])
def test_run_bands_dev_multi_one_process(mode):
    """cvs_g2_run_bands_dev_multi: host image in, caller-owned device planes on devices[0] out.  Uses every GPU of the box;
    on a single-GPU box the device list repeats GPU 0 (three bands on one device)."""
    H, W = 900, 640
    img = synth(7200, H, W)
    whole = _whole(img)
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    lib = capi.lib()
    shapes = [whole[l]["theta"].shape[1:] for l in range(L)]
    outs = [{p: torch.full(shapes[l], -7.0, device="cuda:0") for p in (capi.THETA, capi.STRENGTH, capi.E)} for l in range(L)]
    lvl = (C.POINTER(C.c_void_p) * L)()
    keep = []
    for l in range(L):
        arr = (C.c_void_p * capi.G2_NPLANES)()
        for p, t in outs[l].items():
            arr[p] = t.data_ptr()
        keep.append(arr)
        lvl[l] = C.cast(arr, C.POINTER(C.c_void_p))
    pitches = (C.c_size_t * L)(*[s[1] * 4 for s in shapes])
    darr = (C.c_int * len(devs))(*devs)
    himg = torch.from_numpy(img)
    torch.cuda.synchronize()      # the call runs on the library's own (non-blocking) streams: the fills above must have landed
    capi.check(lib.cvs_g2_run_bands_dev_multi(len(devs), darr, 4, 0.67, himg.data_ptr(), H, W, W * 4, L, capi.G2_MASK_ORIENT, mode, lvl, pitches))
    torch.cuda.synchronize()
    for l in range(L):
        for p, name in ((capi.THETA, "theta"), (capi.STRENGTH, "strength"), (capi.E, "e")):
            assert torch.equal(outs[l][p], whole[l][name][0].to("cuda:0")), (l, name)


# ---- one process per rank ---------------------------------------------------------------------------------------------
def _worker(rank, world, port, q, backend):
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(150, exit=True)
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dev = rank if backend == "nccl" else 0
        torch.cuda.set_device(dev)
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        H, W = 1000, 700
        img = synth(7300, H, W)
        run = BandRun(H, W, L, device=dev, world=world, rank=rank, nccl=(backend == "nccl"))
        run.load_image_rows(torch.from_numpy(img))
        whole = _whole(img) if rank == 0 else None
        ok = True
        modes = ("peer", "copy") + (("nccl",) if backend == "nccl" else ())
        for mode in modes:
            if rank == 0:
                for l in range(L):
                    for t in run.root_planes(l).values():
                        t.fill_(-3.0)
            torch.cuda.synchronize()
            dist.barrier()
            run.step(mode)
            torch.cuda.synchronize()
            dist.barrier()            # (the gloo variant has no on-stream barrier: stream sync + control-plane barrier)
            if rank == 0:
                for l in range(L):
                    got = run.root_planes(l)
                    ok = ok and all(torch.equal(got[k], whole[l][k][0]) for k in NAMES)
        # compute only: local planes hold this rank's rows
        run.step("none")
        torch.cuda.synchronize()
        if rank != 0:
            g = run.geometry(0)
            ok = ok and run.local_planes(0)["theta"].shape[0] == g["out_hi"] - g["out_lo"]
        if rank == 0:
            q.put(bool(ok))
        dist.barrier()
        run.close()
        dist.destroy_process_group()
    except Exception:
        q.put("rank %d: %s" % (rank, traceback.format_exc()))
        os._exit(3)


def _run_ranks(world, backend):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, backend)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = q.get(timeout=170)
    finally:
        for p in procs:
            p.join(20)
            if p.is_alive():
                p.kill()
    assert res is True, res


def test_two_processes_share_one_gpu_over_ipc():
    """Peer stores / peer copies through a CUDA IPC mapping of the root's block; gloo carries the 64-byte handle."""
    _run_ranks(2, "gloo")


def test_two_gpus_nccl_and_peer():
    """Two GPUs, one process each: NCCL gather (library-owned communicator), peer stores and peer copies over NVLink."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_ranks(2, "nccl")
