"""Back-to-back projection kernel (gwen_linear_b2b_fwd) vs two GEMM launches at one shape (developer tool).
  python tools/bench_b2b.py M K1 N1 N2 [relu2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402,F401
from gwen_b200 import ops  # noqa: E402

m, k1, n1, n2 = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(m, k1, device=dev).bfloat16()
w1 = (torch.randn(n1, k1, device=dev) * (1.0 / k1 ** 0.5)).bfloat16()
b1 = torch.randn(n1, device=dev) * 0.1
w2 = (torch.randn(n2, n1, device=dev) * (1.0 / n1 ** 0.5)).bfloat16()
y = torch.empty(m, n2, device=dev, dtype=torch.bfloat16)
y2 = torch.empty_like(y)
hbuf = torch.empty(m, n1, device=dev, dtype=torch.bfloat16)


def fused():
    ops.linear_b2b(x, w1, b1, True, w2, out=y)


def unfused():
    ops.linear(x, w1, b1, relu=True, out=hbuf)
    ops.linear(hbuf, w2, None, out=y2)


res = {}
for name, fn in (("b2b", fused), ("two", unfused)):
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
rows = min(m, 4096)
ref = (torch.relu(x[:rows].float() @ w1.float().t() + b1).bfloat16().float() @ w2.float().t())
err = ((y[:rows].float() - ref).abs().max() / ref.abs().max()).item()
diff = (y.float() - y2.float()).abs().max().item()
flops = 2.0 * m * (k1 * n1 + n1 * n2)
print("case", m, k1, n1, n2, "equal", torch.equal(y, y2), "maxdiff %.3g" % diff, "err_vs_fp32 %.3g" % err,
      "b2b %.3f ms (%.0f TFLOP/s)" % (res["b2b"], flops / res["b2b"] * 1e-9), "two %.3f ms" % res["two"], flush=True)
