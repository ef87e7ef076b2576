"""One stencil aggregation at a cfg 3 shape (bf16, F = 512, B = 2) for ncu captures of k_grid_stencil."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

h, w, b, f = 1158, 774, 2, 512
dev = torch.device("cuda:0")
g = gw.build_graph(gw.grid(h, w, dev), h * w)
x = torch.randn(b, h * w, f, device=dev).bfloat16()
out = torch.empty_like(x)
for _ in range(6):
    ops.aggregate(g, x, kernel="stencil", out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.aggregate(g, x, kernel="stencil", out=out)
e1.record()
torch.cuda.synchronize()
print("ok us/launch %.1f" % (e0.elapsed_time(e1) * 100))
