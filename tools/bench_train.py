"""Forward + backward of the six-layer GNNModel (BASELINE config 5 shape: COSMO-1E grid, C = 64, bf16
activations) with the masked L1 loss of the reference (models_gnn.py:261-265, target = input).
Developer tool: per-step time and the autograd kernel breakdown.

  python tools/bench_train.py [--h 1158 --w 774 --batch 1 --dtype bf16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=1158)
    ap.add_argument("--w", type=int, default=774)
    ap.add_argument("--c", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
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
    mask = (torch.arange(n, device=dev) % 125) == 124

    def step():
        for p in model.parameters():
            p.grad = None
        y = model(x, ei)
        loss = gw.loss_func(y[:, mask].float(), x[:, mask].float(), torch.ones(int(mask.sum()), dtype=torch.bool, device=dev)) \
            if b > 1 else gw.loss_func(y[0].float(), x[0].float(), mask)
        loss.backward()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    with torch.no_grad():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model(x, ei)
        e1.record()
        torch.cuda.synchronize()
        fwd = e0.elapsed_time(e1)
    print(json.dumps({"fwd_bwd_ms": round(ms, 2), "fwd_only_ms": round(fwd, 2), "batch": b, "nodes": n, "dtype": args.dtype,
                      "loss": float(loss), "member_steps_per_s": round(b * 1e3 / ms, 2)}), flush=True)
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))


if __name__ == "__main__":
    main()
