"""Two-rank NCCL run of the row-band pipeline (needs >= 2 GPUs; skipped otherwise): gathered result on the root is
bit-identical to the single-GPU whole-image pyramid -- through the NCCL gather and through direct peer-memory stores."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, backend):
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(120, exit=True)   # a rank that blocks (peer died) must not outlive the test
    try:
        _worker_body(rank, world, port, q, backend)
    except Exception:
        q.put("rank %d: %s" % (rank, traceback.format_exc()))
        os._exit(3)


def _worker_body(rank, world, port, q, backend):
    import torch.distributed as dist
    from cvsteer_b200 import capi, multi
    from cvsteer_b200.batch import G2Batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:   # both ranks on GPU 0: CUDA IPC works between processes on one device; gloo carries the control plane
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W, L = 1000, 700, 5
    img = np.random.default_rng(11).uniform(0, 255, (H, W)).astype(np.float32)
    process, down, _ = multi.cuda_callables(capi.G2_MASK_ORIENT)
    load = lambda lo, hi: torch.from_numpy(img[lo:hi].copy()).cuda()  # noqa: E731
    ok, whole = True, None
    if rank == 0:
        whole = G2Batch().run_pyramid(torch.from_numpy(img[None]).cuda(), L, capi.G2_MASK_ORIENT)
    if backend == "nccl":
        full, local, plan = multi.run_bands(load, H, W, L, process, down)
        if rank == 0:
            ok = all(torch.equal(full[l][k], whole[l][k][0]) for l in range(L) for k in full[l])
    # compute + gather as ONE kernel per level: every rank's fused kernel stores its rows into the root's planes (peer memory)
    peer = multi.PeerPlanes(["theta", "strength", "e"], H, W, L)
    for _ in range(2):   # the mapping is reusable across steps
        full2, _, _ = multi.run_bands(load, H, W, L, process, down, gather="direct", peer=peer)
    if rank == 0:
        ok = ok and all(torch.equal(full2[l][k], whole[l][k][0]) for l in range(L) for k in full2[l])
        q.put(ok)
    dist.barrier()
    del full2
    peer.close()
    dist.destroy_process_group()


def _run_two_ranks(backend):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = q.get(timeout=150)          # rank 0's verdict, or the first traceback
    finally:
        for p in procs:
            p.join(20)
            if p.is_alive():
                p.kill()
    assert res is True, res


def test_bands_nccl_two_ranks():
    """NCCL gather and direct peer-memory stores over NVLink, two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_two_ranks("nccl")


def test_bands_direct_peer_stores_one_gpu():
    """The direct gather (CUDA IPC mapping of the root's planes, stores from the fused kernel, fence) with two processes
    sharing GPU 0 -- everything but the NVLink hop, runnable on a single-GPU box."""
    _run_two_ranks("gloo")
