"""Secondary workloads of BASELINE.json (configs[0], configs[2..4]); the contract bench (bench.py) covers configs[1].

  cfg1  the cvsteer-run per-file body on the bundled 256x185 test image: latency of ONE file through the C ABI
        (cvs_g2_lines_u8_host, n=1), throughput of a 2048-file batch, and the cv2 oracle on one host core beside them

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


def cfg1(device):
    """configs[0]: host-timed (the call synchronises): 8-bit gray in host memory -> three 8-bit maps in host memory."""
    import ctypes as C

    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fish = np.load(os.path.join(root, "tests", "golden", "fish_fixture.npz"))["fish"]
    rows, cols = fish.shape
    lib = capi.lib()
    h = C.c_void_p()
    capi.check(lib.cvs_g2_create(C.byref(h), device, 4, 0.67))

    def run(batch, outs):
        n = batch.shape[0]
        capi.check(lib.cvs_g2_lines_u8_host(h, batch.data_ptr(), n, rows, cols, cols, rows * cols, 0.0, outs[0].data_ptr(),
                                            outs[1].data_ptr(), outs[2].data_ptr(), cols, rows * cols))

    res = {}
    for n, reps in ((1, 200), (2048, 10)):
        batch = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(fish, (n, rows, cols)))).pin_memory()
        outs = [torch.empty_like(batch).pin_memory() for _ in range(3)]
        for _ in range(3):
            run(batch, outs)
        t0 = time.perf_counter()
        for _ in range(reps):
            run(batch, outs)
        res[n] = (time.perf_counter() - t0) / reps
    from oracle import cvsteer_ref as ref          # the CPU leg of a bench: the one place tools may time the oracle
    import cv2
    cv2.setNumThreads(1)

    def cpu_one():
        _, (g2, h2, e, mag, ph) = ref.g2_full(fish)
        return [ref.normalize_minmax_u8(m) for m in (ref.find_edges(mag, ph), ref.find_dark_lines(mag, ph), ref.find_bright_lines(mag, ph))]
    cpu_one()
    t0 = time.perf_counter()
    for _ in range(20):
        cpu_one()
    t_cpu = (time.perf_counter() - t0) / 20
    lib.cvs_g2_destroy(h)
    px = rows * cols
    return {"config": "cfg1", "what": "cvsteer-run per-file body (gray u8 -> edges / dark lines / bright lines u8) on the bundled %dx%d image" % (cols, rows),
            "n_gpus": 1, "one_file_ms": round(res[1] * 1e3, 4), "one_file_Mpix_s": round(px / 1e6 / res[1], 1),
            "batch_2048_files_ms": round(res[2048] * 1e3, 3), "batch_Mpix_s": round(2048 * px / 1e6 / res[2048], 1),
            "batch_files_per_s": round(2048 / res[2048], 0), "timing": "host wall clock around the synchronous C-ABI call, pinned host buffers",
            "cpu_oracle_one_core_ms": round(t_cpu * 1e3, 3), "cpu_oracle_Mpix_s": round(px / 1e6 / t_cpu, 1)}


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
        elif cfg == "cfg1":
            if rank == 0:
                print(json.dumps(cfg1(local)), flush=True)
            continue
        elif cfg == "cfg5":
            H = W = a.big
            process, down, _ = multi.cuda_callables(capi.G2_MASK_ORIENT)
            plan = multi.plan_bands(H, world, L)[rank]
            band = torch.rand((plan.have[0][1] - plan.have[0][0], W), device=dev) * 255   # this rank's band + halo, resident
            t_c = timed(lambda: multi.run_bands(lambda lo_, hi_: band, H, W, L, process, down, gather=False), 1, a.steps, world, dev)
            t_g = timed(lambda: multi.run_bands(lambda lo_, hi_: band, H, W, L, process, down, gather=True), 1, a.steps, world, dev)
            line = {"config": "cfg5", "what": "%dx%d image, G2/H2 M1 + 5-level pyramid, %d row bands + halo" % (H, W, world), "n_gpus": world,
                    "compute_only_ms": round(t_c, 3), "compute_plus_gather_ms": round(t_g, 3),
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
