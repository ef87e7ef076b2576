"""Full GNNModel forward at BASELINE config 3 (COSMO-1E grid 1158 x 774, C = 64, bf16, B = 8 members):
our path (K0 once, K1 stencil, K2 tcgen05) against the torch-op sequence torch_geometric 2.3.1
executes for GCNConv on a GPU (add_remaining_self_loops + gcn_norm per call,
F.linear, index_select, multiply, scatter_add_, + bias, relu).  Developer tool; bench.py reports
the same numbers under "full_forward".

  python tools/bench_forward.py [--h 1158 --w 774 --batch 8 --dtype bf16] [--layers] [--no-torch]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops  # noqa: E402

WIDTHS = (1024, 512, 256, 512, 1024)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def torch_gcn_norm(edge_index, n, dtype):
    """add_remaining_self_loops + symmetric normalisation, the op sequence PyG runs per call."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop = torch.arange(n, device=edge_index.device)
    ei = torch.cat([edge_index[:, keep], torch.stack([loop, loop])], dim=1)
    ew = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    deg = torch.zeros(n, dtype=dtype, device=ei.device).scatter_add_(0, ei[1], ew)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return ei, dis[ei[0]] * ew * dis[ei[1]]


def torch_gcn_conv(x, edge_index, weight, bias, relu, cache=None):
    """One GCNConv.forward as PyG 2.3.1 executes it (x [N, F_in], one member)."""
    n = x.size(-2)
    ei, ew = cache if cache is not None else torch_gcn_norm(edge_index, n, x.dtype)
    h = torch.nn.functional.linear(x, weight)
    msg = h.index_select(-2, ei[0]) * ew.view(-1, 1)
    out = torch.zeros_like(h).scatter_add_(-2, ei[1].view(-1, 1).expand_as(msg), msg)
    out = out + bias
    return torch.relu(out) if relu else out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=1158)
    ap.add_argument("--w", type=int, default=774)
    ap.add_argument("--c", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--no-torch", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    h, wd, c, b = args.h, args.w, args.c, args.batch
    n = h * wd
    torch.manual_seed(23)
    ei = gw.grid(h, wd, dev)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
    model = gw.GNNModel(cfg).to(dev).to(dt)
    x = torch.randn(b, n, c, device=dev).to(dt)
    e1 = gw.grid_edge_count(h, wd)
    with torch.no_grad():
        ms = timeit(lambda: model(x, ei))
        y = model(x, ei)
    res = {"impl": "gwen_b200", "fwd_ms": round(ms, 3), "batch": b, "nodes": n, "messages": e1, "dtype": args.dtype,
           "grid_steps_per_s": round(1e3 / ms, 3), "member_steps_per_s": round(b * 1e3 / ms, 2),
           "edges_per_s": round(b * e1 * 6 / ms * 1e3)}
    print(json.dumps(res), flush=True)
    d, u = model.conv_layers.down_conv_layers, model.conv_layers.up_conv_layers
    convs = (("conv1", d.conv1, True), ("conv2", d.conv2, True), ("conv3", d.conv3, True),
             ("upconv3", u.upconv3, True), ("upconv4", u.upconv4, True), ("upconv5", u.upconv5, False))
    if args.layers:
        g = gw.get_graph(ei, n)
        hcur = x
        with torch.no_grad():
            for name, conv, relu in convs:
                ms_l = timeit(lambda: conv(hcur, g, relu=relu), iters=5, warm=1)
                fa = min(conv.in_channels, conv.out_channels)
                src = hcur if conv.in_channels < conv.out_channels else \
                    torch.empty(hcur.shape[:-1] + (fa,), device=dev, dtype=dt)
                ms_a = timeit(lambda: ops.aggregate(g, src), iters=5, warm=1)
                esz = src.element_size()
                print(json.dumps({"layer": name, "in": conv.in_channels, "out": conv.out_channels,
                                  "ms": round(ms_l, 3), "agg_ms": round(ms_a, 3), "agg_width": fa,
                                  "agg_GBs": round(2 * src.numel() * esz / ms_a / 1e6, 1),
                                  "gemm_ms": round(ms_l - ms_a, 3),
                                  "gemm_TFLOPs": round(2.0 * b * n * conv.in_channels * conv.out_channels / (ms_l - ms_a) / 1e9, 1)}),
                      flush=True)
                hcur = conv(hcur, g, relu=relu)
    if not args.no_torch:
        # the PyG op sequence, one member at a time (its [E', F] message tensor is 16.5 GB at F = 1024)
        def torch_forward(xm, cached):
            cache = torch_gcn_norm(ei, n, xm.dtype) if cached else None
            for _, conv, relu in convs:
                xm = torch_gcn_conv(xm, ei, conv.lin.weight, conv.bias, relu, cache)
            return xm
        with torch.no_grad():
            for cached in (False, True):
                ms_t = timeit(lambda: torch_forward(x[0], cached), iters=3, warm=1) * b
                print(json.dumps({"impl": "torch op sequence of PyG 2.3.1 GCNConv (GPU, %s)" %
                                  ("norm computed once per forward" if cached else "norm recomputed per layer call, as GWEN runs it"),
                                  "fwd_ms": round(ms_t, 2), "note": "%d members run one after the other" % b,
                                  "grid_steps_per_s": round(1e3 / ms_t, 3), "speedup_ours": round(ms_t / ms, 2)}), flush=True)
            yt = torch_forward(x[0], True)
            err = ((y[0].float() - yt.float()).abs().max() / yt.float().abs().max()).item()
            print(json.dumps({"max_abs_diff_over_max_abs_vs_torch_sequence": err}), flush=True)


if __name__ == "__main__":
    main()
