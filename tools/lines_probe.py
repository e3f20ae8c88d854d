"""Where does the time of cvs_g2_lines_u8_host go?  (scratch probe: PCIe copies alone, kernels alone, the whole call)"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402
from cvsteer_b200.batch import G2Batch  # noqa: E402

n, rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 185, 256
x = torch.randint(0, 256, (n, rows, cols), dtype=torch.uint8).pin_memory()
outs = [torch.empty_like(x).pin_memory() for _ in range(3)]
lib = capi.lib()
h = C.c_void_p()
capi.check(lib.cvs_g2_create(C.byref(h), 0, 4, 0.67))


def wall(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def call(gain):
    capi.check(lib.cvs_g2_lines_u8_host(h, x.data_ptr(), n, rows, cols, cols, rows * cols, gain, outs[0].data_ptr(), outs[1].data_ptr(),
                                        outs[2].data_ptr(), cols, rows * cols))


d = torch.empty_like(x, device="cuda")
dd = [torch.empty_like(x, device="cuda") for _ in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def copies():
    with torch.cuda.stream(s1):
        d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2):
        for i in range(3):
            outs[i].copy_(dd[i], non_blocking=True)


g = G2Batch()
mask = capi.bit(capi.EDGES) | capi.bit(capi.DARK) | capi.bit(capi.BRIGHT)
fo = {p: torch.empty((n, rows, cols), device="cuda") for p in (capi.EDGES, capi.DARK, capi.BRIGHT)}
print("whole call, minmax      ms", round(wall(lambda: call(0.0)), 3))
print("whole call, gain        ms", round(wall(lambda: call(0.5)), 3))
print("PCIe copies alone       ms", round(wall(copies), 3))
print("h2d alone               ms", round(wall(lambda: d.copy_(x, non_blocking=True)), 3))
print("d2h alone (3 maps)      ms", round(wall(lambda: [outs[i].copy_(dd[i], non_blocking=True) for i in range(3)]), 3))
print("fused kernel alone      ms", round(wall(lambda: g.run(d, mask, outs=fo)), 3))
m8 = torch.empty((3 * n, rows, cols), dtype=torch.uint8, device="cuda")
fm = torch.empty((3 * n, rows, cols), device="cuda")
print("to_u8 minmax (3n frames) ms", round(wall(lambda: capi.check(lib.cvs_to_u8_dev(0, fm.data_ptr(), 3 * n, rows, cols, cols * 4, rows * cols * 4, 0.0,
                                                                                      m8.data_ptr(), cols, rows * cols, None))), 3))
print("to_u8 gain   (3n frames) ms", round(wall(lambda: capi.check(lib.cvs_to_u8_dev(0, fm.data_ptr(), 3 * n, rows, cols, cols * 4, rows * cols * 4, 0.5,
                                                                                      m8.data_ptr(), cols, rows * cols, None))), 3))
