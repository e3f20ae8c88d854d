"""Development timing of the fused kernels (kernel-only, CUDA events).  Not the contract bench (see bench.py)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402
from cvsteer_b200.batch import G2Batch, G4Batch, ffma_peak, pyr_down  # noqa: E402


def time_ms(fn, warm=3, it=10):
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
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--rows", type=int, default=1080)
    ap.add_argument("--cols", type=int, default=1920)
    ap.add_argument("--g4", action="store_true")
    ap.add_argument("--ffma", action="store_true")
    ap.add_argument("--only-g4", action="store_true")
    ap.add_argument("--lines", action="store_true", help="only the cvsteer-run mask (edges / dark / bright), float and 8-bit input")
    a = ap.parse_args()
    out = {}
    if a.ffma:
        for form, nm in ((0, "imm"), (1, "reg"), (2, "const")):
            v, ms = ffma_peak(form, 20000)
            out[f"ffma_{nm}_Tinstr_s"] = round(v / 1e12, 3)
    x = torch.rand((a.n, a.rows, a.cols), device="cuda") * 255
    mpix = a.n * a.rows * a.cols / 1e6
    g = G2Batch()
    if a.lines:
        outs = {p: torch.empty((a.n, a.rows, a.cols), device="cuda") for p in (capi.EDGES, capi.DARK, capi.BRIGHT)}
        for name, xin, bpp in (("lines_f32", x, 16), ("lines_u8", x.to(torch.uint8), 13)):
            ms = time_ms(lambda: g.run(xin, capi.G2_MASK_LINES, outs=outs))
            out[name] = {"ms": round(ms, 4), "Gpix_s": round(mpix / ms, 2), "GB_s": round(mpix * bpp / ms, 1), "k": g.last_launch()["kernel"]}
        print(json.dumps(out))
        return
    for name, mask, bpp in (() if a.only_g4 else (("M0", capi.G2_MASK_STATE, 52), ("M1", capi.G2_MASK_ORIENT, 16), ("M2", capi.G2_MASK_FULL, 32))):
        outs = {p: torch.empty((a.n, a.rows, a.cols), device="cuda") for p in range(capi.G2_NPLANES) if mask >> p & 1}
        ms = time_ms(lambda: g.run(x, mask, outs=outs))
        out[name] = {"ms": round(ms, 4), "Gpix_s": round(mpix / ms, 2), "GB_s": round(mpix * bpp / ms, 1), "k": g.last_launch()["kernel"]}
        del outs
    ms = time_ms(lambda: pyr_down(x))
    out["pyr_down"] = {"ms": round(ms, 4), "GB_s": round(mpix * 5 / ms, 1)}
    if a.g4:
        g4 = G4Batch()
        th = torch.rand((a.n, a.rows, a.cols), device="cuda")
        for name, mask, bpp, kw in (("G4_M0", capi.G4_MASK_BASIS, 48, {}),
                                    ("G4_M2", capi.G4_MASK_STEER, 24, dict(steer=capi.STEER_MAP, theta_map=th))):
            outs = {p: torch.empty((a.n, a.rows, a.cols), device="cuda") for p in range(capi.G4_NPLANES) if mask >> p & 1}
            ms = time_ms(lambda: g4.run(x, mask, outs=outs, **kw))
            out[name] = {"ms": round(ms, 4), "Gpix_s": round(mpix / ms, 2), "GB_s": round(mpix * bpp / ms, 1)}
            del outs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
