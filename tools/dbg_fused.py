import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw
from gwen_b200 import ops
h, w, b, k, n = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda:0")
g = gw.build_graph(gw.grid(h, w, dev), h * w)
x = torch.randn(b, h * w, k, device=dev).bfloat16()
wt = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
bias = torch.randn(n, device=dev)
torch.cuda.synchronize()
t0 = time.time()
y = ops.gcn_fused(g, x, wt, bias, relu=True)
torch.cuda.synchronize()
t1 = time.time()
y2 = ops.linear(ops.aggregate(g, x, kernel="stencil"), wt, bias, relu=True)
torch.cuda.synchronize()
hbuf = torch.empty_like(x)
def unfused():
    ops.aggregate(g, x, kernel="stencil", out=hbuf)
    ops.linear(hbuf, wt, bias, relu=True, out=y2)
for fn in (lambda: ops.gcn_fused(g, x, wt, bias, relu=True, out=y), unfused):
    for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.gcn_fused(g, x, wt, bias, relu=True, out=y)
e1.record(); torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(5): unfused()
f1.record(); torch.cuda.synchronize()
print("case", h, w, b, k, n, "equal", torch.equal(y, y2), "first %.3fs" % (t1 - t0), "fused %.3f ms" % (e0.elapsed_time(e1) / 5), "unfused %.3f ms" % (f0.elapsed_time(f1) / 5), flush=True)
