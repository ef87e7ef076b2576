"""K1 aggregation sweep on one GPU (developer tool): kernel x tile x slab at a given grid/F/dtype.
Prints one JSON line per variant with the CUDA-event time and the algorithmic-bytes bandwidth."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3  # median, min in us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=582)
    ap.add_argument("--w", type=int, default=390)
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    n = args.h * args.w
    g = gw.build_graph(gw.grid(args.h, args.w, dev), n)
    x = torch.randn(args.batch, n, args.feat, device=dev).to(dt)
    out = torch.empty_like(x)
    esz = x.element_size()
    alg = 2 * args.batch * n * args.feat * esz + 4 * (n + 1) + 8 * g.num_messages
    ref = ops.aggregate(g, x, kernel="rows")
    variants = [("rows", None, 0)]
    vn = 16 // esz
    tiles = ((8, 32), (8, 16), (4, 32), (16, 16), (4, 64), (2, 64), (2, 128), (4, 128), (128,))
    if args.quick:
        tiles = ((8, 16),)
    for tile in tiles:
        for chunks in (8, 16, 32):
            if chunks * vn <= args.feat:
                variants.append(("tiled", tile, chunks * vn))
    # rows kernel with the tile order as a locality hint
    if not args.quick:
        for tile in ((8, 32), (4, 16)):
            variants.append(("rows+order", tile, 0))
    for tw in (16, 32, 64):
        for chunks in (8, 16, 32):
            if chunks * vn <= args.feat:
                variants.append(("stencil", (tw,), chunks * vn))
    for kern, tile, slab in variants:
        try:
            if kern == "rows+order":
                g.order = g.tile_plan(tile).order
                fn = lambda: ops.aggregate(g, x, kernel="rows", out=out)
            else:
                g.order = None
                fn = lambda: ops.aggregate(g, x, kernel=kern, tile=tile, slab=slab, out=out)
            fn()
            ok = torch.equal(out, ref) if kern != "stencil" else \
                ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
            med, mn = timeit(fn, args.iters)
            extra = {}
            extra["stages"] = os.environ.get("GWEN_STENCIL_STAGES", "")
            if kern == "tiled":
                pl = g.tile_plan(tile)
                extra = {"amp": round(pl.amplification, 3), "runs": pl.max_tile_runs, "rl": pl.run_len}
            extra["env"] = "/".join(os.environ.get(k, "") for k in ("GWEN_TILED_THREADS", "GWEN_TILED_STAGES", "GWEN_TILED_U"))
            print(json.dumps({"kernel": kern, "tile": tile, "slab": slab, **extra, "us_med": round(med, 1),
                              "us_min": round(mn, 1), "GBs_alg": round(alg / med / 1e3, 1),
                              "bitwise_ok": ok}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"kernel": kern, "tile": tile, "slab": slab, "error": str(e)[:200]}), flush=True)
    g.order = None
    # context: plain device copy of the same bytes
    med, mn = timeit(lambda: out.copy_(x), args.iters)
    print(json.dumps({"kernel": "torch_copy", "us_med": round(med, 1), "GBs": round(2 * x.numel() * esz / med / 1e3, 1)}))


if __name__ == "__main__":
    main()
