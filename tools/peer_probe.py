"""2-rank probe of multi.PeerPlanes (direct peer-memory gather): prints progress per rank; dumps tracebacks and exits
if anything blocks for 60 s.  torchrun --nproc-per-node 2 tools/peer_probe.py"""
import faulthandler
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi, multi  # noqa: E402
from cvsteer_b200.batch import G2Batch  # noqa: E402

faulthandler.dump_traceback_later(60, exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])


def say(*a):
    print(f"[rank {rank}]", *a, flush=True)


try:
    backend = os.environ.get("PROBE_BACKEND", "nccl")
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    else:   # one GPU, two processes: CUDA IPC works between processes on the same device; gloo carries the control plane
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    H, W, L = 1000, 700, 5
    img = np.random.default_rng(11).uniform(0, 255, (H, W)).astype(np.float32)
    process, down, _ = multi.cuda_callables(capi.G2_MASK_ORIENT)
    say("creating PeerPlanes")
    peer = multi.PeerPlanes(["theta", "strength", "e"], H, W, L)
    say("mapped; flat device", peer._flat.device, "ptr", hex(peer._flat.data_ptr()))
    for it in range(2):
        full, _, plan = multi.run_bands(lambda lo, hi: torch.from_numpy(img[lo:hi].copy()).cuda(), H, W, L, process, down,
                                        gather="direct", peer=peer)
        say("step", it, "done; out rows", plan.out[0])
    if rank == 0:
        whole = G2Batch().run_pyramid(torch.from_numpy(img[None]).cuda(), L, capi.G2_MASK_ORIENT)
        ok = all(torch.equal(full[l][k], whole[l][k][0]) for l in range(L) for k in full[l])
        say("bitwise equal to the whole-image run:", ok)
    dist.barrier()
    del full
    peer.close()
    say("teardown")
    dist.destroy_process_group()
    say("exit")
except Exception:
    traceback.print_exc()
    sys.stdout.flush()
    os._exit(3)
