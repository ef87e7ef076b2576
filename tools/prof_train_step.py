"""Three gwen_b200.train_step calls at the cfg 5 member shape (for an ncu launch list of the training step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402

dev = torch.device("cuda:0")
h, wd, c = 1158, 774, 64
n = h * wd
torch.manual_seed(23)
ei = gw.grid(h, wd, dev)
cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
model = gw.GNNModel(cfg).to(dev)
x1 = torch.randn(1, n, c, device=dev).to(torch.bfloat16)
mask = (torch.arange(n, device=dev) % 125) == 124
for _ in range(3):
    gw.train_step(model, x1, ei, mask)
torch.cuda.synchronize()
print("done")
