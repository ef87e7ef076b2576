"""Aggregation on BASELINE config 2's mesh (582 x 390, F = 256, fp32) with a fraction of the nodes cut out (a
land/sea mask): the masked-mesh fast path (stencil with dis zeroed at the cut-out nodes + gwen_rows_self_fwd)
against the general CSR kernels on the same graph, and a grid-numbered NON-mesh graph (improved=True) through
kernel="auto" (tiled CSR kernel).  Algorithmic bytes as SURVEY 8(d).  Developer tool; output -> profiles/.
  python tools/bench_masked.py [frac]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
h, w, f = 582, 390, 256
dev = torch.device("cuda:0")
n = h * w
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6650.0


def timeit(fn, iters=200, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


ei = gw.grid(h, w, dev)
g = torch.Generator(device="cpu").manual_seed(23)
cut = (torch.rand(n, generator=g) < frac).to(dev)
keep = ~(cut[ei[0]] | cut[ei[1]])
eim = ei[:, keep].contiguous()
gm = gw.build_graph(eim, n)
xs = [torch.randn(n, f, device=dev) for _ in range(3)]      # rotate buffers: 1.4 GB >> L2
outs = [torch.empty(n, f, device=dev) for _ in range(3)]
bias = torch.randn(f, device=dev) * 0.1
i = [0]


def run(graph, kernel):
    def fn():
        k = i[0] % 3
        i[0] += 1
        ops.aggregate(graph, xs[k], bias, kernel=kernel, out=outs[k])
    return fn


msgs = gm.num_messages
alg = 2 * n * f * 4 + 4 * (n + 1) + 4 * msgs + 4 * n
res = {"mesh": "%dx%d" % (h, w), "feat": f, "cut_fraction": frac, "cut_nodes": int(cut.sum()), "messages": msgs,
       "mesh_kind": gm.mesh_kind, "algorithmic_bytes": alg, "peak_gbs": peak}
for name, kern in (("masked_stencil (auto)", "auto"), ("tiled CSR", "tiled"), ("rows CSR", "rows")):
    us = timeit(run(gm, kern))
    res[name] = {"us": round(us, 1), "GBs": round(alg / us / 1e3, 1), "frac_of_copy_peak": round(alg / us / 1e3 / peak, 3)}
a = ops.aggregate(gm, xs[0], bias)
b = ops.aggregate(gm, xs[0], bias, kernel="rows")
res["max_abs_diff_vs_rows_over_max"] = ((a - b).abs().max() / b.abs().max()).item()
gi = gw.build_graph(ei, n, improved=True)
alg_i = 2 * n * f * 4 + 4 * (n + 1) + 4 * gi.num_messages + 4 * n + 4 * gi.num_messages
for name, kern in (("improved=True mesh, auto (tiled CSR)", "auto"), ("improved=True mesh, rows CSR", "rows")):
    us = timeit(run(gi, kern))
    res[name] = {"us": round(us, 1), "GBs": round(alg_i / us / 1e3, 1), "frac_of_copy_peak": round(alg_i / us / 1e3 / peak, 3)}
print(json.dumps(res, indent=1))
