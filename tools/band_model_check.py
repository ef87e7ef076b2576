"""Partitioned six-layer model (BandGNNModel on PeerMeshBand) against the un-partitioned model:
parity (forward bitwise, gradients to fp32 rounding) on a small ragged mesh, then STRONG-scaling
timings at BASELINE config 3 (1158 x 774, B = 8, bf16) and config 4 (2048 x 2048, B = 1, bf16).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/band_model_check.py [--no-time]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import partition  # noqa: E402


def timed(fn, warm, iters, dev):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--profile", action="store_true", help="kernel breakdown of one cfg 5 step on rank 0")
    ap.add_argument("--cfg5-batch", type=int, default=0,
                    help="also time the config 5 training step (1158 x 774, this many members) on the bands")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {"world": world}
    # ---- parity (fp32, ragged bands, batch 2) ---------------------------------------------------
    h, w, c, hid, b = 8 * world + 3, 29, 16, 64, 2
    n = h * w
    torch.manual_seed(23)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg).to(dev)
    with torch.no_grad():
        for p in model.parameters():          # non-zero biases
            if p.dim() == 1:
                p.normal_(0, 0.1)
    ei = gw.grid(h, w, dev)
    g = gw.get_graph(ei, n)
    gen = torch.Generator(dev).manual_seed(7)
    x = torch.randn(b, n, c, device=dev, generator=gen)
    t = torch.randn(b, n, c, device=dev, generator=gen)
    y_full = model(x, ei)
    (y_full * t).sum().backward()
    ref_grads = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    band = partition.PeerMeshBand(h, w, g.dis)
    net = partition.BandGNNModel(model, band)
    sl = slice(band.r0 * w, (band.r0 + band.rows) * w)
    ok_f, gerr = True, 0.0
    for it in range(2):                      # twice: buffers reused, epochs advance
        for p in model.parameters():
            p.grad = None
        y_own = net(x[:, sl].contiguous())
        ok_f &= torch.equal(y_own, y_full[:, sl].detach())
        (y_own * t[:, sl]).sum().backward()
        net.allreduce_grads()
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        assert len(grads) == len(ref_grads)
        for ga, gr in zip(grads, ref_grads):
            gerr = max(gerr, ((ga - gr).abs().max() / gr.abs().max().clamp_min(1e-12)).item())
    # reference loss_func over the whole mesh == sum of the ranks' shares; its gradient after all-reduce
    mask = (torch.arange(n, device=dev) % 5) == 4
    for p in model.parameters():
        p.grad = None
    lf = gw.masked_l1_loss(model(x, ei), x, mask)
    lf.backward()
    ref_l = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    ls = net.loss(net(x[:, sl].contiguous()), x[:, sl].contiguous(), mask[sl].contiguous())
    ls.backward()
    net.allreduce_grads()
    lsum = ls.detach().clone().reshape(1)
    dist.all_reduce(lsum)
    lerr = abs(lsum.item() - lf.item()) / abs(lf.item())
    for ga, gr in zip([p.grad for p in model.parameters() if p.grad is not None], ref_l):
        gerr = max(gerr, ((ga - gr).abs().max() / gr.abs().max().clamp_min(1e-12)).item())
    res["loss_rel_err"] = lerr
    okt = torch.tensor([1.0 if (ok_f and lerr < 1e-5) else 0.0, -gerr], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    res["forward_bitwise_equal"] = bool(okt[0].item())
    res["grad_max_rel_err"] = -okt[1].item()
    ok = res["forward_bitwise_equal"] and res["grad_max_rel_err"] < 1e-4
    # ---- bf16: the ReLU masks folded into the next layer's dgrad epilogue (nn.ReluLink logic of the band model) on
    # vs off: the same gradients bit for bit (weights; biases: the same values summed by the same kernel) ------------
    from gwen_b200 import nn as gnn
    h2, w2, c2, hid2, b2 = 16 * world + 5, 64, 64, 256, 2
    n2 = h2 * w2
    cfg2 = gw.GNNConfig(nodes_in=n2, nodes_out=n2, channels_in=c2, channels_out=c2, hidden_feats=hid2)
    model2 = gw.GNNModel(cfg2).to(dev).to(torch.bfloat16)
    g2 = gw.get_graph(gw.grid(h2, w2, dev), n2)
    band2 = partition.PeerMeshBand(h2, w2, g2.dis)
    net2 = partition.BandGNNModel(model2, band2)
    xo2 = torch.randn(b2, band2.n_own, c2, device=dev, generator=gen).to(torch.bfloat16)
    to2 = torch.randn(b2, band2.n_own, c2, device=dev, generator=gen)
    got, used = {}, {}
    old_flag = gnn.BWD_MASK_FUSION
    for fused in (True, False):
        gnn.BWD_MASK_FUSION = fused
        for p in model2.parameters():
            p.grad = None
        (net2(xo2).float() * to2).sum().backward()
        used[fused] = len(net2._masked)            # flags are consumed by the producer layers: 0 left after backward
        net2.allreduce_grads()
        got[fused] = [(nm, p.grad.clone()) for nm, p in model2.named_parameters() if p.grad is not None]
    gnn.BWD_MASK_FUSION = old_flag
    same = used[True] == 0 and len(got[True]) == len(got[False]) > 0
    for (nm, ga), (_, gb) in zip(got[True], got[False]):
        if nm.endswith("bias"):
            same &= bool(((ga.float() - gb.float()).norm() <= 1e-6 * gb.float().norm().clamp_min(1e-30)).item())
        else:
            same &= bool(torch.equal(ga, gb))
    st = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(st, op=dist.ReduceOp.MIN)
    res["bf16_mask_fusion_bitwise"] = bool(st.item())
    ok = ok and res["bf16_mask_fusion_bitwise"]
    del model2, net2, band2
    # ---- strong-scaling timings -------------------------------------------------------------------
    if not args.no_time:
        del model, net, band
        for name, (h, w, b) in (("cfg3_1158x774_B8", (1158, 774, 8)), ("cfg4_2048x2048_B1", (2048, 2048, 1))):
            n, c = h * w, 64
            gw.clear_graph_cache()
            ei = gw.grid(h, w, dev)
            g = gw.get_graph(ei, n)
            cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
            torch.manual_seed(23)
            model = gw.GNNModel(cfg).to(dev).to(torch.bfloat16)
            band = partition.PeerMeshBand(h, w, g.dis)
            net = partition.BandGNNModel(model, band)
            del ei
            xo = torch.randn(b, band.n_own, c, device=dev).to(torch.bfloat16)
            with torch.no_grad():
                ms_f = timed(lambda: net(xo), 2, 5, dev)

            def train():
                for p in model.parameters():
                    p.grad = None
                net(xo).float().abs().mean().backward()
                net.allreduce_grads()
            ms_t = timed(train, 1, 3, dev) if b == 1 else None
            res[name] = {"fwd_ms": round(ms_f, 3), "fwd_bwd_allreduce_ms": None if ms_t is None else round(ms_t, 3)}
            del model, net, band, xo
            torch.cuda.empty_cache()
    if args.cfg5_batch > 0:
        h, w, b, c = 1158, 774, args.cfg5_batch, 64
        n = h * w
        gw.clear_graph_cache()
        ei = gw.grid(h, w, dev)
        g = gw.get_graph(ei, n)
        cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=1024)
        torch.manual_seed(23)
        model = gw.GNNModel(cfg).to(dev).to(torch.bfloat16)
        band = partition.PeerMeshBand(h, w, g.dis)
        net = partition.BandGNNModel(model, band)
        del ei
        xo = torch.randn(b, band.n_own, c, device=dev).to(torch.bfloat16)
        ids = torch.arange(band.r0 * w, (band.r0 + band.rows) * w, device=dev)
        mo = (ids % 125) == 124

        def train():
            for p in model.parameters():
                p.grad = None
            net.loss(net(xo), xo, mo).backward()
            net.allreduce_grads()
        ms_t = timed(train, 1, 3, dev)
        if args.profile and rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                train()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60), flush=True)
        elif args.profile:
            train()
            torch.cuda.synchronize()
        res["cfg5_1158x774_B%d" % b] = {"train_step_ms": round(ms_t, 3), "member_steps_per_s": round(b * 1e3 / ms_t, 2),
                                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    if rank == 0:
        print(json.dumps(res), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
