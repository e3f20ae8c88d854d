"""ONE process, 2 GPUs: cvs_g2_run_bands_dev_multi with fused peer stores -- GPU 1's fused kernel writes its band straight into
GPU 0's planes over NVLink.  For `ncu --devices 1 -k regex:k_march -c 1` (the level-0 launch of GPU 1) and as a plain timing."""
import argparse
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--mode", type=int, default=capi.GATHER_PEER_STORE)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    L, H, W = 5, a.rows, a.cols
    assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
    lib = capi.lib()
    img = (torch.rand((H, W)) * 255).pin_memory()
    shapes = [(H, W)]
    for _ in range(L - 1):
        shapes.append(((shapes[-1][0] + 1) // 2, (shapes[-1][1] + 1) // 2))
    outs = [{p: torch.empty(s, device="cuda:0") for p in (capi.THETA, capi.STRENGTH, capi.E)} for s in shapes]
    lvl = (C.POINTER(C.c_void_p) * L)()
    keep = []
    for l in range(L):
        arr = (C.c_void_p * capi.G2_NPLANES)()
        for p, t in outs[l].items():
            arr[p] = t.data_ptr()
        keep.append(arr)
        lvl[l] = C.cast(arr, C.POINTER(C.c_void_p))
    pitches = (C.c_size_t * L)(*[s[1] * 4 for s in shapes])
    devs = (C.c_int * 2)(0, 1)
    for i in range(a.reps):
        t0 = time.perf_counter()
        capi.check(lib.cvs_g2_run_bands_dev_multi(2, devs, 4, 0.67, img.data_ptr(), H, W, W * 4, L, capi.G2_MASK_ORIENT, a.mode, lvl, pitches))
        print(f"rep {i}: {1e3 * (time.perf_counter() - t0):.2f} ms (upload + compute + gather, host clock)")
    torch.cuda.synchronize()
    print("checksum", float(outs[0][capi.STRENGTH][H // 2 + 5, 100]))


if __name__ == "__main__":
    main()
