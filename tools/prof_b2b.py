"""A few back-to-back projection launches at a cfg 3 shape (for ncu captures of k_linear_b2b).
  python tools/prof_b2b.py K1 N1 N2 [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402,F401
from gwen_b200 import ops  # noqa: E402

k1, n1, n2 = [int(v) for v in sys.argv[1:4]]
m = int(sys.argv[4]) if len(sys.argv) > 4 else 2 * 1158 * 774
dev = torch.device("cuda:0")
x = torch.randn(m, k1, device=dev).bfloat16()
w1 = (torch.randn(n1, k1, device=dev) / k1 ** 0.5).bfloat16()
w2 = (torch.randn(n2, n1, device=dev) / n1 ** 0.5).bfloat16()
b1 = torch.randn(n1, device=dev) * 0.1
y = torch.empty(m, n2, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.linear_b2b(x, w1, b1, True, w2, out=y)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
