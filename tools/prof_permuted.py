"""A few launches of the aggregation over locality tiles on the permuted cfg 2 mesh (for ncu) and a small sweep that
separates the memory system from the kernel: the same staged kernel on the NATURAL numbering (identical code, ordered
addresses), slab widths, and a plain permuted row copy (gwen_rows_gather) as the random-row bandwidth of the machine.
  python tools/prof_permuted.py [ncu]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

h, w, f = 582, 390, 256
dev = torch.device("cuda:0")
n = h * w
ei = gw.grid(h, w, dev)
perm = torch.randperm(n, generator=torch.Generator().manual_seed(23)).to(dev)
gp = gw.build_graph(perm[ei].contiguous(), n)
x = torch.randn(n, f, device=dev)
out = torch.empty(n, f, device=dev)
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    for _ in range(4):
        ops.aggregate(gp, x, kernel="locality", out=out.unsqueeze(0))
    torch.cuda.synchronize()
    sys.exit(0)


def timeit(fn, iters=100, warm=5):
    return _timeit(fn, iters, warm)


def _timeit(fn, iters, warm):
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


if len(sys.argv) > 1 and sys.argv[1] == "quick":
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if len(sys.argv) > 3:
        type(gp).LOCALITY_CAP_ROWS = int(sys.argv[3])
    if len(sys.argv) > 4:
        type(gp).LOCALITY_MERGE_ROWS = int(sys.argv[4])
    pl = gp.locality_plan(r or None)
    res = []
    for dt in (torch.float32, torch.bfloat16):
        xs = [torch.randn(n, f, device=dev).to(dt) for _ in range(3)]
        outs = [torch.empty(n, f, device=dev, dtype=dt) for _ in range(3)]
        i = [0]

        def fn():
            k = i[0] % 3
            i[0] += 1
            ops.aggregate(gp, xs[k], kernel="locality", out=outs[k].unsqueeze(0), plan=pl)
        us = timeit(fn)
        ok = torch.equal(outs[0], ops.aggregate(gp, xs[0], kernel="rows"))
        res.append("%s %.1f us %s" % ("fp32" if dt == torch.float32 else "bf16", us, "ok" if ok else "MISMATCH"))
    print("mode=%s warps=%s stages=%s radius=%s tiles=%d maxsrc=%d amp=%.2f | %s" % (
        os.environ.get("GWEN_GATHER_MODE", "-"), os.environ.get("GWEN_GATHER_WARPS", "-"),
        os.environ.get("GWEN_TILED_STAGES", "-"), r, pl.num_tiles, pl.max_tile_runs, pl.amplification, " | ".join(res)))
    sys.exit(0)

xs = [torch.randn(n, f, device=dev) for _ in range(3)]
outs = [torch.empty(n, f, device=dev) for _ in range(3)]
i = [0]


def rot(fn):
    def g():
        k = i[0] % 3
        i[0] += 1
        fn(xs[k], outs[k])
    return g


print("locality tiles, permuted ids       %.1f us" % timeit(rot(lambda a, o: ops.aggregate(gp, a, kernel="locality", out=o.unsqueeze(0)))))
for slab in (32, 64, 128):
    try:
        print("  slab %3d floats                   %.1f us" % (slab, timeit(rot(lambda a, o: ops.aggregate(gp, a, kernel="locality", out=o.unsqueeze(0), slab=slab)))))
    except RuntimeError as e:
        print("  slab", slab, "->", str(e)[:100])
# the same graph in its natural numbering through the SAME gather kernel: locality tiles of the unpermuted mesh
gn = gw.build_graph(ei, n, grid_shape=None)
pl = gn.locality_plan()
print("locality tiles, natural ids         %.1f us  (amp %.2f)" % (timeit(rot(lambda a, o: ops.aggregate(gn, a, kernel="locality", out=o.unsqueeze(0)))), pl.amplification))
gg = gw.build_graph(ei, n)
print("2-D tiles + TMA runs, natural ids   %.1f us" % timeit(rot(lambda a, o: ops.aggregate(gg, a, kernel="tiled", out=o.unsqueeze(0)))))
print("rows, permuted ids                  %.1f us" % timeit(rot(lambda a, o: ops.aggregate(gp, a, kernel="rows", out=o.unsqueeze(0)))))
print("rows, natural ids                   %.1f us" % timeit(rot(lambda a, o: ops.aggregate(gg, a, kernel="rows", out=o.unsqueeze(0)))))
idx = perm.to(torch.int32)
print("row gather out[i] = x[perm[i]]      %.1f us" % timeit(rot(lambda a, o: ops.rows_gather(a.unsqueeze(0), idx, o.unsqueeze(0)))))
print("torch copy                          %.1f us" % timeit(rot(lambda a, o: o.copy_(a))))
