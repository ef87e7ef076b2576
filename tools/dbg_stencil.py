import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw
from gwen_b200 import ops
dev = torch.device("cuda:0")
h, w, f = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = gw.build_graph(gw.grid(h, w, dev), h * w)
x = torch.randn(h * w, f, device=dev)
ref = ops.aggregate(g, x, kernel="rows")
torch.cuda.synchronize()
try:
    out = ops.aggregate(g, x, kernel="stencil")
    torch.cuda.synchronize()
    print("ok", ((out - ref).abs().max() / ref.abs().max()).item())
except Exception as e:
    print("ERR", str(e)[:300])
