"""Launch the M2 kernel a few times on config-2 frames (for a single-kernel ncu capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvsteer_b200 import capi  # noqa: E402
from cvsteer_b200.batch import G2Batch  # noqa: E402

x = torch.rand((64, 1080, 1920), device="cuda") * 255
g = G2Batch()
outs = {p: torch.empty((64, 1080, 1920), device="cuda") for p in range(capi.G2_NPLANES) if capi.G2_MASK_FULL >> p & 1}
for _ in range(5):
    g.run(x, capi.G2_MASK_FULL, outs=outs)
torch.cuda.synchronize()
print(g.last_launch()["kernel"])
