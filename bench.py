#!/usr/bin/env python
"""Contract benchmark for the cvsteer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode M0|M1|M2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused G2/H2 basis+steer+orientation kernel over one batch of synthetic frames
(BASELINE.json configs[1]: 64 frames of 1920x1080, fp32, single scale) already resident in HBM.  Rank 0 prints
ONE JSON line.  See DESIGN.md "Measurement" for how every field is derived.

  value      whole-job Mpix/s, device-timed (CUDA events on the launching stream, max over ranks)
  e2e        the same metric through the host-buffer C-ABI call (cvs_g2_run_batch_host): pinned host frames in,
             pinned host planes out, H2D + D2H inside the timed region
  roofline   algorithmic bytes of the mode / measured launch duration vs MEASURED_PEAKS.json (HBM), plus the FP32
             side against an FFMA-saturation microbenchmark run in the same process
  cpu_baseline  the oracle (the reference's glue over the same OpenCV primitives, cv2) on this box's host cores,
             frame-parallel like the reference's cv::parallel_for_ (example/steer.cpp:169), bounded sample

--impl reference runs only that CPU path and prints the same line shape with "impl": "reference".
--cfg1 measures BASELINE.json configs[0] instead (the cvsteer-run per-file body on the bundled test image: one-file
latency and file-batch throughput through cvs_g2_lines_u8_host, with the oracle on one host core beside them).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, FRAMES = 1080, 1920, 64          # BASELINE.json configs[1]
WORKLOAD = "cfg2: G2/H2 (width 4, spacing 0.67) on 64 synthetic 1920x1080 fp32 frames per GPU, single scale"
# algorithmic bytes / fp32 instructions per pixel per mode (DESIGN.md "Roofline")
MODES = {
    "M0": dict(bpp=52, ipp=167, what="materialise class state: 7 basis planes + c1..c3 + theta_d + strength"),
    "M1": dict(bpp=16, ipp=167, what="orientation: theta_d, strength, energy at theta_d"),
    "M2": dict(bpp=32, ipp=217, what="basis+steer+orientation: theta_d, strength, g2, h2, e, magnitude, phase"),
}


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle over cv2), frame-parallel
# ------------------------------------------------------------------------------------------------
def _cpu_frames(n, seed0):
    import numpy as np
    return [np.random.default_rng(seed0 + i).uniform(0, 255, (ROWS, COLS)).astype(np.float32) for i in range(n)]


def _cpu_one(mode, img):
    from oracle import cvsteer_ref as ref
    if mode == "M2":
        ref.g2_full(img)            # ctor + steer(theta_d map, g2,h2,e,magnitude,phase): what both reference callers run
    elif mode == "M1":
        ref.g2_orientation(img)
    else:
        ref.SteerableFiltersG2(img)


def cpu_pass(mode, frames, workers):
    """One frame-parallel pass; returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(lambda f: _cpu_one(mode, f), frames))
    return time.perf_counter() - t0


def cpu_setup():
    import cv2
    cv2.setNumThreads(1)            # one worker thread per core, OpenCV single-threaded inside each
    cores = os.cpu_count() or 1
    workers = min(cores, 64)
    return cores, workers


def cpu_baseline(mode):
    cores, workers = cpu_setup()
    nframes = max(16, 2 * workers)
    nframes = min(nframes, 96)
    frames = _cpu_frames(nframes, 2000)
    cpu_pass(mode, frames[:workers], workers)          # warm-up
    dt = cpu_pass(mode, frames, workers)
    return {"value": round(nframes * ROWS * COLS / 1e6 / dt, 2), "unit": "Mpix/s", "cores": workers, "kind": "port",
            "sample": f"{nframes} of the workload's 1920x1080 frames, mode {mode}, {workers} frame-parallel threads "
                      f"(host has {cores} logical cores), cv2.setNumThreads(1), 1 warm-up pass"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores, workers = cpu_setup()
    per_step = max(8, workers)
    per_step = min(per_step, 64)
    frames = _cpu_frames(per_step, 2000)
    for _ in range(args.warmup):
        cpu_pass(args.mode, frames, workers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(args.mode, frames, workers)
    dt = (time.perf_counter() - t0) / args.steps
    val = round(per_step * ROWS * COLS / 1e6 / dt, 2)
    sample = (f"each step = {per_step} of the workload's 64 frames, {workers} frame-parallel threads "
              f"(host has {cores} logical cores), cv2.setNumThreads(1)")
    line = {"impl": "reference", "metric": "G2/H2 basis+steer throughput", "value": val, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": args.mode, "what": MODES[args.mode]["what"],
                       "arm": "oracle/cvsteer_ref.py: the reference's glue over cv2 (the reference C++ cannot be "
                              "built here: no OpenCV C++ headers)"},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML, polled during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.stop_flag, self.ok = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        win = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-5:]
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        bits = 0
        for s in win:
            bits |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in win) if win else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in names.items() if bits & k), "samples": len(win)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from cvsteer_b200 import capi
    from cvsteer_b200.batch import G2Batch, ffma_peak

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: cvsteer_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode
    mask = {"M0": capi.G2_MASK_STATE, "M1": capi.G2_MASK_ORIENT, "M2": capi.G2_MASK_FULL}[mode]
    planes = [p for p in range(capi.G2_NPLANES) if mask >> p & 1]
    px = FRAMES * ROWS * COLS

    gen = torch.Generator(device=dev)
    gen.manual_seed(2000 + rank)
    x = torch.rand((FRAMES, ROWS, COLS), device=dev, generator=gen) * 255.0       # 531 MB > 126 MB L2
    outs = {p: torch.empty((FRAMES, ROWS, COLS), device=dev) for p in planes}
    g = G2Batch(device=local_rank)
    lib = capi.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        g.run(x, mask, outs=outs)
    barrier()
    l0 = lib.cvs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        g.run(x, mask, outs=outs)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.cvs_launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.summary(t0, t1)
    launch = g.last_launch()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * px / 1e6 / (ms / 1e3)

    # ---- other modes, kernel-only, for the roofline table (rank 0, N=1 only: keep multi-rank runs short)
    extra = {}
    if world == 1 and not args.quick:
        for m2, mk in (("M0", capi.G2_MASK_STATE), ("M1", capi.G2_MASK_ORIENT), ("M2", capi.G2_MASK_FULL)):
            if m2 == mode:
                continue
            o2 = {p: (outs[p] if p in outs else torch.empty((FRAMES, ROWS, COLS), device=dev))
                  for p in range(capi.G2_NPLANES) if mk >> p & 1}
            for _ in range(3):
                g.run(x, mk, outs=o2)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(20):
                g.run(x, mk, outs=o2)
            a1.record()
            torch.cuda.synchronize()
            t_ms = a0.elapsed_time(a1) / 20
            extra[m2] = {"ms": round(t_ms, 4), "Mpix_s": round(px / 1e6 / (t_ms / 1e3), 1),
                         "GB_s": round(px * MODES[m2]["bpp"] / 1e9 / (t_ms / 1e3), 1)}
            del o2

    # ---- roofline of the dominant (only) kernel of the step
    hbm_peak, peak_src = load_peaks()
    bpp, ipp = MODES[mode]["bpp"], MODES[mode]["ipp"]
    ach_gbs = px * bpp / 1e9 / (ms / 1e3)
    fp32 = None
    if rank == 0:
        try:
            peak_i = max(ffma_peak(0, 20000, local_rank)[0], ffma_peak(2, 20000, local_rank)[0])
            ach_i = px * ipp / (ms / 1e3)
            fp32 = {"achieved_Tinstr_s": round(ach_i / 1e12, 2), "peak_Tinstr_s": round(peak_i / 1e12, 2),
                    "frac": round(ach_i / peak_i, 3),
                    "peak_how": "FFMA-saturation microbenchmark (cvs_bench_ffma, best of immediate / constant-bank operand), same process"}
        except Exception as ex:  # pragma: no cover
            fp32 = {"error": str(ex)}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(mode)
        except Exception:
            traffic = None
    t_hbm = px * bpp / (hbm_peak * 1e9)
    t_fp = (px * ipp / (fp32["peak_Tinstr_s"] * 1e12)) if fp32 and "peak_Tinstr_s" in fp32 else 0.0
    roof = {"bound": "hbm", "achieved": round(ach_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
            "frac": round(ach_gbs / hbm_peak, 3), "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_px": bpp, "algorithmic_fp32_instr_per_px": ipp, "kernel": launch["kernel"],
            "kernel_ms": round(ms, 4), "fp32": fp32,
            "binding": "fp32" if t_fp > t_hbm else "hbm",
            "frac_of_binding_roofline": round(max(t_hbm, t_fp) / (ms / 1e3), 3)}

    # ---- end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        xh = torch.empty((FRAMES, ROWS, COLS), dtype=torch.float32).pin_memory()
        xh.copy_(x)
        oh = {p: torch.empty((FRAMES, ROWS, COLS), dtype=torch.float32).pin_memory() for p in planes}
        g.run_host(xh, mask, oh)                                   # warm-up (allocates the device staging ring)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            g.run_host(xh, mask, oh)                               # synchronous: returns when results are in host memory
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        chk = float(oh[planes[0]][0, 100, 100])                    # touch a result on the host
        e2e = {"value": round(world * px / 1e6 / dt, 1), "unit": "Mpix/s", "h2d_bytes_per_step": px * 4,
               "d2h_bytes_per_step": px * 4 * len(planes), "steps": e2e_steps, "ms_per_step": round(dt * 1e3, 2),
               "api": "cvs_g2_run_batch_host (pinned host buffers, 3-stage H2D/kernel/D2H pipeline)",
               "probe": chk}
        del xh, oh

    sampler.stop_flag = True
    cpu = cpu_baseline(mode) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {"metric": "G2/H2 basis+steer throughput", "value": round(value, 1), "unit": "Mpix/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "mode": mode, "what": MODES[mode]["what"],
                           "frames_per_gpu": FRAMES, "rows": ROWS, "cols": COLS, "sharding": "frames, no collective",
                           "l2": "inputs (531 MB/GPU) larger than the 126 MB L2; no flush needed",
                           "grid": launch["grid"], "block": launch["block"], "smem": launch["smem"]},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "other_modes_kernel_only": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cfg1_line(device=0):
    """configs[0]: host-timed (the call synchronises): 8-bit gray in host memory -> three 8-bit maps in host memory."""
    import ctypes as C

    import numpy as np
    import torch
    from cvsteer_b200 import capi
    root = os.path.dirname(os.path.abspath(__file__))
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
        batch = torch.from_numpy(np.repeat(fish[None], n, axis=0)).pin_memory()
        outs = [torch.empty_like(batch).pin_memory() for _ in range(3)]
        for _ in range(3):
            run(batch, outs)
        t0 = time.perf_counter()
        for _ in range(reps):
            run(batch, outs)
        res[n] = (time.perf_counter() - t0) / reps
    from oracle import cvsteer_ref as ref          # CPU leg (cpu_baseline): the oracle, timed beside the GPU path
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
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="M2", choices=list(MODES))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the other-modes table")
    ap.add_argument("--cfg1", action="store_true", help="configs[0] instead: cvsteer-run per-file body on the bundled test image")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.cfg1:
        if rank == 0:
            print(json.dumps(cfg1_line(local_rank)), flush=True)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # `python bench.py --gpus N` without torchrun: re-launch under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
