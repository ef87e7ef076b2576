"""One tensor-core projection at a GWEN layer shape (for ncu captures of k_linear_tc2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gwen_b200 import ops  # noqa: E402

m, k, n = 896292, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
x = torch.randn(m, k, device=dev).bfloat16()
w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
b = torch.randn(n, device=dev)
for _ in range(6):
    y = ops.linear(x, w, b, relu=True)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
