"""Launches for ncu: the masked dgrad (512 -> 1024 at M = 896 292: conv2's dgrad with conv1's ReLU mask in the
epilogue) and the wgrad that also delivers the bias sums (64 -> 1024: conv1).   python tools/prof_bwd.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gwen_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
m = 896292
torch.manual_seed(1)
dh = torch.randn(m, 512, device=dev).bfloat16()
w = torch.randn(512, 1024, device=dev) * 0.03
y1 = torch.relu(torch.randn(m, 1024, device=dev)).bfloat16()
for _ in range(3):
    dx = ops.linear_bwd_data_masked(dh, w, y1)
assert dx is not None
dy = dx
h = torch.randn(m, 64, device=dev).bfloat16()
for _ in range(3):
    both = ops.linear_bwd_weight_bias(dy, h)
assert both is not None
torch.cuda.synchronize()
print("done")
