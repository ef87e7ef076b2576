"""Aggregation on BASELINE config 2's mesh (582 x 390, F = 256, fp32 and bf16) with RANDOMLY PERMUTED node ids (a
graph whose numbering carries no locality, as an unstructured grid in file order): the row kernel against the staged
kernel over locality tiles (csrc/locality.cu: K0r finds compact patches from the CSR alone; the producer warp
gathers a tile's source rows one by one with cp.async).  x and out stay in the caller's (permuted) numbering.
Algorithmic bytes as SURVEY 8(d) with stored per-edge weights (a general CSR graph reads src AND w).
Developer tool; output -> profiles/.
  python tools/bench_permuted.py [radius [cap_rows [merge_rows]]]     (radius 0 = search)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402
from gwen_b200.graph import GraphCSR  # noqa: E402

radius = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if len(sys.argv) > 2:
    GraphCSR.LOCALITY_CAP_ROWS = int(sys.argv[2])
if len(sys.argv) > 3:
    GraphCSR.LOCALITY_MERGE_ROWS = int(sys.argv[3])
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


tri = os.environ.get("GWEN_BENCH_TRI") == "1"     # triangular cells (three neighbours), same node count
if tri:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.graphs import tri_mesh_edges
    h, w = 291, 390
    ei = tri_mesh_edges(h, w).to(dev)
    n = 2 * h * w
else:
    ei = gw.grid(h, w, dev)
g = torch.Generator(device="cpu").manual_seed(23)
perm = torch.randperm(n, generator=g).to(dev)
eip = perm[ei].contiguous()
gp = gw.build_graph(eip, n)
assert gp.grid_shape is None and gp.mesh_kind is None
torch.cuda.synchronize()
t0 = time.time()
plan = gp.locality_plan(radius or None)
torch.cuda.synchronize()
t_plan = time.time() - t0
msgs = gp.num_messages
res = {"mesh": ("triangular cells of a %dx%d lattice (3 neighbours), node ids randomly permuted" if tri else "%dx%d, node ids randomly permuted") % (h, w), "feat": f, "messages": msgs, "peak_gbs": peak,
       "locality_plan": None if plan is None else {
           "radius": radius or getattr(gp, "locality_radius", None), "tiles": plan.num_tiles,
           "max_tile_rows": plan.max_tile_rows, "max_tile_sources": plan.max_tile_runs,
           "staged_rows_per_destination_row": round(plan.amplification, 3), "build_s": round(t_plan, 3)}}
bias = torch.randn(f, device=dev) * 0.1
for dt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
    esz = 4 if dt == torch.float32 else 2
    xs = [torch.randn(n, f, device=dev).to(dt) for _ in range(3)]      # rotate buffers
    outs = [torch.empty(n, f, device=dev, dtype=dt) for _ in range(3)]
    i = [0]

    def run(kernel):
        def fn():
            k = i[0] % 3
            i[0] += 1
            ops.aggregate(gp, xs[k], bias, kernel=kernel, out=outs[k])
        return fn

    alg = 2 * n * f * esz + 4 * (n + 1) + 8 * msgs + 4 * n
    r = {"algorithmic_bytes": alg}
    for label, kern in (("locality tiles (auto)", "auto"), ("rows CSR", "rows")):
        us = timeit(run(kern))
        r[label] = {"us": round(us, 1), "GBs": round(alg / us / 1e3, 1), "frac_of_copy_peak": round(alg / us / 1e3 / peak, 3)}
    a = ops.aggregate(gp, xs[0], bias)
    b = ops.aggregate(gp, xs[0], bias, kernel="rows")
    r["bitwise_equal_to_rows"] = bool(torch.equal(a, b))
    res[name] = r
print(json.dumps(res, indent=1))
