"""Launch ONE fused kernel variant three times (for `ncu -s 2 -c 1 --set full ...`): python tools/prof_one.py g4s --n 4 --size 4k"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402
from cvsteer_b200.batch import G2Batch, G4Batch, pyr_down  # noqa: E402

KERNELS = {
    "m0": (2, capi.G2_MASK_STATE, capi.STEER_DOMINANT), "m1": (2, capi.G2_MASK_ORIENT, capi.STEER_DOMINANT),
    "m2": (2, capi.G2_MASK_FULL, capi.STEER_DOMINANT), "lines": (2, capi.G2_MASK_LINES, capi.STEER_DOMINANT),
    "steer5s": (2, capi.G2_MASK_STEER5, capi.STEER_SCALAR), "steer5m": (2, capi.G2_MASK_STEER5, capi.STEER_MAP),
    "dyn": (2, capi.G2_MASK_FULL | capi.bit(capi.G2A), capi.STEER_DOMINANT),
    "g4b": (4, capi.G4_MASK_BASIS, capi.STEER_DOMINANT), "g4s": (4, capi.G4_MASK_STEER, capi.STEER_MAP),
    "g4ss": (4, capi.G4_MASK_STEER, capi.STEER_SCALAR), "g4sd": (4, capi.G4_MASK_STEER, capi.STEER_DOMINANT),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", choices=list(KERNELS) + ["pyr"])
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--size", default="4k")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    R, C = {"4k": (2160, 3840), "1080p": (1080, 1920)}[a.size]
    x = torch.rand((a.n, R, C), device="cuda") * 255
    if a.kernel == "pyr":
        for _ in range(a.reps):
            pyr_down(x)
        torch.cuda.synchronize()
        return
    fam, mask, steer = KERNELS[a.kernel]
    g = G2Batch() if fam == 2 else G4Batch()
    npl = capi.G2_NPLANES if fam == 2 else capi.G4_NPLANES
    outs = {p: torch.empty((a.n, R, C), device="cuda") for p in range(npl) if mask >> p & 1}
    kw = {}
    if steer == capi.STEER_MAP:
        kw = dict(steer=steer, theta_map=(torch.rand((a.n, R, C), device="cuda") - 0.5) * 3.1)
    elif steer == capi.STEER_SCALAR:
        kw = dict(steer=steer, theta=0.3)
    for _ in range(a.reps):
        g.run(x, mask, outs=outs, **kw)
    torch.cuda.synchronize()
    print(g.last_launch())


if __name__ == "__main__":
    main()
