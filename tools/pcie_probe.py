"""Host <-> device copy ceiling of this box, with 1..N GPUs streaming at once.

    python tools/pcie_probe.py                      # one process, GPU 0
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py

Every rank copies `--mb` MiB blocks between pinned host memory and its GPU for `--seconds`: H2D only, D2H only, and both
directions at once (what the e2e leg of bench.py does: 1 plane up, 7 planes down).  Each case runs twice: with the process
bound to the CPUs (and therefore the memory) of the GPU's NUMA node, and unbound.  Rank 0 prints one JSON object: per-rank
and whole-box GB/s.  This is the denominator of `e2e.frac_of_ceiling` in bench.py's line: if eight GPUs together cannot pull
more than the box's host-memory / root-complex limit, no pipelining inside the library can either."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import hostmem  # noqa: E402


def run_case(dev, nbytes, seconds, h2d, d2h, dist, world):
    hb_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory() if h2d else None
    hb_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory() if d2h else None
    db_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    db_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        if h2d:
            with torch.cuda.stream(s1):
                db_in.copy_(hb_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hb_out.copy_(db_out, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(4):
            once()
        s1.synchronize()
        s2.synchronize()
        n += 4
    dt = time.perf_counter() - t0
    return {"h2d_GB_s": round(n * nbytes / 1e9 / dt, 2) if h2d else 0.0, "d2h_GB_s": round(n * nbytes / 1e9 / dt, 2) if d2h else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=1.0)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    nbytes = a.mb << 20
    allowed = os.sched_getaffinity(0)
    out = {"world": world, "block_MiB": a.mb, "host_cpus": len(allowed), "cases": {}}
    for bound in (True, False):
        os.sched_setaffinity(0, allowed)
        numa = hostmem.bind_to_gpu_numa(local) if bound else {"node": hostmem.gpu_numa_node(local), "bound": False}
        for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
            r = run_case(dev, nbytes, a.seconds, h2d, d2h, dist, world)
            r["numa_node"] = numa.get("node")
            rows = [r]
            if world > 1:
                rows = [None] * world
                dist.all_gather_object(rows, r)
            if rank == 0:
                key = f"{name}_{'numa_bound' if bound else 'unbound'}"
                out["cases"][key] = {"per_rank": rows, "total_h2d_GB_s": round(sum(x["h2d_GB_s"] for x in rows), 1),
                                     "total_d2h_GB_s": round(sum(x["d2h_GB_s"] for x in rows), 1)}
    os.sched_setaffinity(0, allowed)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
