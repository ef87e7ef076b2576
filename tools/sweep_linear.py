"""K2 projection sweep (developer tool): the six GWEN layer shapes at a COSMO-1E-sized M, our
tcgen05 GEMM vs torch.matmul (cuBLAS) for context, plus the full GNNModel forward at config 3."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
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
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=896292)
    ap.add_argument("--model", action="store_true")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if dt == torch.bfloat16 else 4
    for k, n in ((64, 1024), (1024, 512), (512, 256), (256, 512), (512, 1024), (1024, 64)):
        x = torch.randn(args.m, k, device=dev).to(dt)
        w = (torch.randn(n, k, device=dev) * 0.05).to(dt)
        b = torch.randn(n, device=dev)
        ms = timeit(lambda: ops.linear(x, w, b, relu=True))
        ms_t = timeit(lambda: torch.relu(torch.nn.functional.linear(x, w, b.to(dt))))
        ms_mm = timeit(lambda: torch.nn.functional.linear(x, w))
        fl = 2.0 * args.m * k * n
        by = (args.m * (k + n) + n * k) * esz
        print(json.dumps({"M": args.m, "K": k, "N": n, "ours_ms": round(ms, 3), "TFLOPs": round(fl / ms / 1e9, 1),
                          "GBs": round(by / ms / 1e6, 1), "torch_linear_relu_ms": round(ms_t, 3),
                          "cublas_matmul_ms": round(ms_mm, 3)}), flush=True)
        del x, w
    if args.model:
        h, wd, c = 1158, 774, 64
        n = h * wd
        ei = gw.grid(h, wd, dev)
        cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
        model = gw.GNNModel(cfg).to(dev).to(dt)
        x = torch.randn(args.batch, n, c, device=dev).to(dt)
        with torch.no_grad():
            ms = timeit(lambda: model(x, ei), iters=5, warm=2)
        e1 = gw.grid_edge_count(h, wd)
        print(json.dumps({"model_fwd_ms": round(ms, 2), "batch": args.batch, "nodes": n,
                          "grid_steps_per_s": round(1e3 / ms, 3), "member_steps_per_s": round(args.batch * 1e3 / ms, 2),
                          "edges_per_s": round(args.batch * e1 * 6 / ms * 1e3), "dtype": args.dtype}), flush=True)
        # per-layer breakdown with CUDA events
        g = gw.get_graph(ei, n)
        d, u = model.conv_layers.down_conv_layers, model.conv_layers.up_conv_layers
        hcur = x
        with torch.no_grad():
            for name, conv, relu in (("conv1", d.conv1, True), ("conv2", d.conv2, True), ("conv3", d.conv3, True),
                                     ("upconv3", u.upconv3, True), ("upconv4", u.upconv4, True), ("upconv5", u.upconv5, False)):
                ms = timeit(lambda: conv(hcur, g, relu=relu), iters=5, warm=1)
                agg_first = conv.in_channels < conv.out_channels
                fa = min(conv.in_channels, conv.out_channels)
                src = hcur if agg_first else torch.empty(hcur.shape[:-1] + (fa,), device=dev, dtype=dt)
                ms_a = timeit(lambda: ops.aggregate(g, src), iters=5, warm=1)
                print(json.dumps({"layer": name, "in": conv.in_channels, "out": conv.out_channels,
                                  "ms": round(ms, 3), "agg_ms": round(ms_a, 3), "agg_width": fa}), flush=True)
                hcur = conv(hcur, g, relu=relu)


if __name__ == "__main__":
    main()
