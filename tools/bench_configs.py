"""Secondary workloads of BASELINE.json (configs[2..4]); the contract bench (bench.py) covers configs[1], and
`python bench.py --cfg1` configs[0] (the cvsteer-run per-file body with its CPU leg).

  cfg3  G2/H2 orientation (M1) over a 5-level pyramid, 3840x2160 frames, batch 256 TOTAL, sharded by frame (strong scaling)
  cfg4  G4/H4 steer at a per-pixel theta map + magnitude + phase, 3840x2160, batch 256 total, sharded by frame
  cfg5  one 32768x32768 image, G2/H2 M1 + 5-level pyramid, row bands with halo, NCCL gather of the outputs to rank 0

Run with torchrun for N > 1.  Prints one JSON line per config on rank 0.  Device-timed (CUDA events), max over ranks.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi, multi  # noqa: E402
from cvsteer_b200.batch import G2Batch, G4Batch, pyr_down  # noqa: E402


def timed(fn, warm, steps, world, dev):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg3,cfg4")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--big", type=int, default=32768)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, C, L = 2160, 3840, 5
    lo, hi = multi.shard_frames(a.frames, world, rank)
    n = hi - lo
    for cfg in a.configs.split(","):
        if cfg == "cfg3":
            x = torch.rand((n, R, C), device=dev) * 255
            g = G2Batch(device=local)
            # pre-allocate every level's outputs so that the timed region holds kernels only
            shapes = [(R, C)]
            for _ in range(L - 1):
                shapes.append(((shapes[-1][0] + 1) // 2, (shapes[-1][1] + 1) // 2))
            outs = [{p: torch.empty((n,) + s, device=dev) for p in (capi.THETA, capi.STRENGTH, capi.E)} for s in shapes]
            lv = [x] + [torch.empty((n,) + s, device=dev) for s in shapes[1:]]   # every level stays resident

            def step():
                # one launch per level: the fused kernel also emits the next level from its staged tile
                for l in range(L):
                    g.run(lv[l], capi.G2_MASK_ORIENT, outs=outs[l], next_level=lv[l + 1] if l + 1 < L else None)
            ms = timed(step, a.warmup, a.steps, world, dev)
            px0 = a.frames * R * C
            line = {"config": "cfg3", "what": "G2/H2 M1 over a 5-level pyramid, 3840x2160, %d frames total" % a.frames, "n_gpus": world,
                    "ms_per_step": round(ms, 3), "Mpix_s_level0": round(px0 / 1e6 / (ms / 1e3), 1), "scaling": "strong",
                    "launches_per_step": L,
                    "algorithmic_GB_s": round(px0 * (1.332 * 16 + 1.328 * 1.25) / 1e9 / (ms / 1e3), 1)}
            del x, outs
        elif cfg == "cfg4":
            x = torch.rand((n, R, C), device=dev) * 255
            g2, g4 = G2Batch(device=local), G4Batch(device=local)
            th = g2.run(x, capi.bit(capi.THETA))["theta"]          # precomputed, resident theta_d of the same frames
            outs = {p: torch.empty((n, R, C), device=dev) for p in (capi.G4T, capi.H4T, capi.MAG4, capi.PHASE4)}
            ms = timed(lambda: g4.run(x, capi.G4_MASK_STEER, steer=capi.STEER_MAP, theta_map=th, outs=outs), a.warmup, a.steps, world, dev)
            px0 = a.frames * R * C
            line = {"config": "cfg4", "what": "G4/H4 steer(theta map)+magnitude+phase, 3840x2160, %d frames total" % a.frames, "n_gpus": world,
                    "ms_per_step": round(ms, 3), "Mpix_s": round(px0 / 1e6 / (ms / 1e3), 1), "scaling": "strong",
                    "algorithmic_GB_s": round(px0 * 24 / 1e9 / (ms / 1e3), 1),
                    "algorithmic_Tinstr_s": round(px0 * 323 / 1e12 / (ms / 1e3), 2), "kernel": g4.last_launch()["kernel"]}
            del x, th, outs
        elif cfg == "cfg5":
            H = W = a.big
            process, down, _ = multi.cuda_callables(capi.G2_MASK_ORIENT)
            plan = multi.plan_bands(H, world, L)[rank]
            band = torch.rand((plan.have[0][1] - plan.have[0][0], W), device=dev) * 255   # this rank's band + halo, resident
            t_c = timed(lambda: multi.run_bands(lambda lo_, hi_: band, H, W, L, process, down, gather=False), 1, a.steps, world, dev)
            t_g = timed(lambda: multi.run_bands(lambda lo_, hi_: band, H, W, L, process, down, gather=True), 1, a.steps, world, dev)
            t_d = None
            if world > 1:   # compute and gather fused: each rank's kernels store into the root's planes over NVLink peer memory
                peer = multi.PeerPlanes(["theta", "strength", "e"], H, W, L)
                t_d = timed(lambda: multi.run_bands(lambda lo_, hi_: band, H, W, L, process, down, gather="direct", peer=peer), 1, a.steps, world, dev)
                peer.close()
            line = {"config": "cfg5", "what": "%dx%d image, G2/H2 M1 + 5-level pyramid, %d row bands + halo" % (H, W, world), "n_gpus": world,
                    "compute_only_ms": round(t_c, 3), "compute_plus_gather_ms": round(t_g, 3),
                    "fused_peer_store_ms": None if t_d is None else round(t_d, 3),
                    "Mpix_s_fused_peer_store": None if t_d is None else round(H * W / 1e6 / (t_d / 1e3), 1),
                    "Mpix_s_compute": round(H * W / 1e6 / (t_c / 1e3), 1), "Mpix_s_with_gather": round(H * W / 1e6 / (t_g / 1e3), 1),
                    "gather_bytes_to_root": int(3 * 4 * 1.332 * H * W * (world - 1) / max(world, 1))}
            del band
        else:
            continue
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
