"""Fused layer kernel (gwen_gcn_fused_fwd) vs stencil + GEMM at one shape (developer tool).
  python tools/bench_fused.py H W B K N"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

h, w, b, k, n = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda:0")
g = gw.build_graph(gw.grid(h, w, dev), h * w)
x = torch.randn(b, h * w, k, device=dev).bfloat16()
wt = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
bias = torch.randn(n, device=dev)
y = torch.empty(b, h * w, n, device=dev, dtype=torch.bfloat16)
y2 = torch.empty_like(y)
hbuf = torch.empty_like(x)


def fused():
    ops.gcn_fused(g, x, wt, bias, relu=True, out=y)


def unfused():
    ops.aggregate(g, x, kernel="stencil", out=hbuf)
    ops.linear(hbuf, wt, bias, relu=True, out=y2)


res = {}
for name, fn in (("fused", fused), ("unfused", unfused)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 5
print("case", h, w, b, k, n, "equal", torch.equal(y, y2), "fused %.3f ms" % res["fused"], "unfused %.3f ms" % res["unfused"], flush=True)
