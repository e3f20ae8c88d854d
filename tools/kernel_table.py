"""Kernel-only table of every public mask / steering-source combination named in include/cvsteer_c.h (CUDA events on the
launching stream, inputs resident in HBM and larger than L2).  One JSON object per line; bench.py carries the contract
numbers, this tool feeds the table in DESIGN.md and profiles/r02_kernel_table.jsonl.

    python tools/kernel_table.py [--sizes 1080p,4k] [--iters 20] [--only g4]

Roofline per row: algorithmic bytes / fp32 instructions per pixel (the extra rows are derived the
same way: 4 B x (planes read + planes written), basis FMAs + point-wise allowance) against MEASURED_PEAKS.json's HBM
figure and the nominal FP32 issue rate 148 SM x 128 lanes x 1.965 GHz = 37.2 T instr/s."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402
from cvsteer_b200.batch import G2Batch, G4Batch, pyr_down  # noqa: E402

FP32_NOMINAL = 148 * 128 * 1.965e9


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_ms(fn, warm=3, it=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1080p,4k")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    hbm = hbm_peak()
    sizes = {"1080p": (64, 1080, 1920), "4k": (32, 2160, 3840)}
    dyn_mask = capi.G2_MASK_FULL | capi.bit(capi.G2A)      # not one of the named masks -> run-time-mask kernel
    # name, family, mask, steer, B/px, instr/px
    rows = [
        ("g2 M0 state", 2, capi.G2_MASK_STATE, capi.STEER_DOMINANT, 52, 167),
        ("g2 M1 orient", 2, capi.G2_MASK_ORIENT, capi.STEER_DOMINANT, 16, 167),
        ("g2 M2 full", 2, capi.G2_MASK_FULL, capi.STEER_DOMINANT, 32, 217),
        ("g2 lines", 2, capi.G2_MASK_LINES, capi.STEER_DOMINANT, 16, 217),
        ("g2 lines u8-in", 2, capi.G2_MASK_LINES, "u8", 13, 217),
        ("g2 steer5@scalar", 2, capi.G2_MASK_STEER5, capi.STEER_SCALAR, 24, 117 + 16 + 16 + 50),
        ("g2 steer5@map", 2, capi.G2_MASK_STEER5, capi.STEER_MAP, 28, 117 + 16 + 16 + 60),
        ("g2 M2@map", 2, capi.G2_MASK_FULL, capi.STEER_MAP, 36, 227),
        ("g2 lines@map", 2, capi.G2_MASK_LINES, capi.STEER_MAP, 20, 227),
        ("g2 dyn (full+g2a)", 2, dyn_mask, capi.STEER_DOMINANT, 36, 217),
        ("g2 dyn steer5+g2a@map", 2, capi.G2_MASK_STEER5 | capi.bit(capi.G2A), capi.STEER_MAP, 32, 209),
        ("g4 basis", 4, capi.G4_MASK_BASIS, capi.STEER_DOMINANT, 48, 273),
        ("g4 steer@map", 4, capi.G4_MASK_STEER, capi.STEER_MAP, 24, 323),
        ("g4 steer@scalar (dyn)", 4, capi.G4_MASK_STEER, capi.STEER_SCALAR, 20, 313),
        ("g4 steer@theta_d (dyn)", 4, capi.G4_MASK_STEER, capi.STEER_DOMINANT, 20, 273 + 120),
    ]
    for sz in a.sizes.split(","):
        n, R, C = sizes[sz]
        x = torch.rand((n, R, C), device="cuda") * 255
        xu8 = x.to(torch.uint8)
        th = (torch.rand((n, R, C), device="cuda") - 0.5) * 3.1
        px = n * R * C
        g2, g4 = G2Batch(), G4Batch()
        for name, fam, mask, steer, bpp, ipp in rows:
            if a.only and a.only not in name:
                continue
            g = g2 if fam == 2 else g4
            np_ = capi.G2_NPLANES if fam == 2 else capi.G4_NPLANES
            outs = {p: torch.empty((n, R, C), device="cuda") for p in range(np_) if mask >> p & 1}
            xin, kw = x, {}
            if steer == "u8":
                xin = xu8
            elif steer == capi.STEER_MAP:
                kw = dict(steer=steer, theta_map=th)
            elif steer == capi.STEER_SCALAR:
                kw = dict(steer=steer, theta=0.3)
            ms = time_ms(lambda: g.run(xin, mask, outs=outs, **kw), it=a.iters)
            gpix = px / 1e9 / (ms / 1e3)
            t_h, t_f = px * bpp / (hbm * 1e9), px * ipp / FP32_NOMINAL
            print(json.dumps({"size": sz, "frames": n, "kernel": name, "launched": g.last_launch()["kernel"], "ms": round(ms, 4),
                              "Gpix_s": round(gpix, 2), "B_px": bpp, "instr_px": ipp, "GB_s": round(gpix * bpp, 1),
                              "Tinstr_s": round(gpix * ipp / 1e3, 2), "frac_hbm_measured": round(gpix * bpp / hbm, 3),
                              "frac_fp32_nominal": round(gpix * ipp * 1e9 / FP32_NOMINAL, 3),
                              "binding": "fp32" if t_f > t_h else "hbm",
                              "frac_of_binding": round(max(t_h, t_f) / (ms / 1e3), 3)}), flush=True)
            del outs
        if not a.only or "pyr" in a.only:
            ms = time_ms(lambda: pyr_down(x), it=a.iters)
            print(json.dumps({"size": sz, "frames": n, "kernel": "pyr_down", "ms": round(ms, 4), "GB_s": round(px * 5 / 1e6 / ms, 1),
                              "frac_hbm_measured": round(px * 5 / 1e6 / ms / hbm, 3)}), flush=True)
        del x, xu8, th
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
