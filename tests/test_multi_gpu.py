"""Two-rank NCCL run of the row-band pipeline (needs >= 2 GPUs; skipped otherwise): gathered result on the root is
bit-identical to the single-GPU whole-image pyramid."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from cvsteer_b200 import capi, multi
    from cvsteer_b200.batch import G2Batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    H, W, L = 1000, 700, 5
    img = np.random.default_rng(11).uniform(0, 255, (H, W)).astype(np.float32)
    process, down, _ = multi.cuda_callables(capi.G2_MASK_ORIENT)
    full, local, plan = multi.run_bands(lambda lo, hi: torch.from_numpy(img[lo:hi].copy()).cuda(), H, W, L, process, down)
    ok = True
    if rank == 0:
        whole = G2Batch().run_pyramid(torch.from_numpy(img[None]).cuda(), L, capi.G2_MASK_ORIENT)
        ok = all(torch.equal(full[l][k], whole[l][k][0]) for l in range(L) for k in full[l])
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_bands_nccl_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
