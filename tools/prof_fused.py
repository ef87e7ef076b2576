"""One fused layer launch at a cfg 3 shape (for ncu captures of k_gcn_fused)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

h, w, b = 1158, 774, 2
k, n = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
g = gw.build_graph(gw.grid(h, w, dev), h * w)
x = torch.randn(b, h * w, k, device=dev).bfloat16()
wt = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
bias = torch.randn(n, device=dev)
y = torch.empty(b, h * w, n, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.gcn_fused(g, x, wt, bias, relu=True, out=y)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
