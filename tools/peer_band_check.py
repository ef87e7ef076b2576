"""Multi-GPU check + timing of PeerMeshBand (halo exchange inside the stencil kernel over NVLink
peer memory) against the single-GPU stencil and against MeshBand (NCCL send/recv).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_band_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops, partition  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    # ---- parity: small ragged mesh, batch 2, several epochs with changing inputs --------------
    for (h, w, f, b) in ((41, 50, 64, 2), (3 * world + 1, 37, 32, 1)):
        g = gw.build_graph(gw.grid(h, w, dev), h * w)
        band = partition.PeerMeshBand(h, w, g.dis)
        xl = band.alloc(b, f, torch.float32, dev)
        bias = torch.randn(f, device=dev, generator=torch.Generator(dev).manual_seed(5))
        for step in range(4):
            gen = torch.Generator(dev).manual_seed(100 + step)
            x = torch.randn(b, h * w, f, device=dev, generator=gen)          # same on all ranks
            full = ops.aggregate(g, x, bias, relu=True, kernel="stencil")
            band.owned(xl[:, :band.n_local]).copy_(x[:, band.r0 * w:(band.r0 + band.rows) * w])
            torch.cuda.synchronize()
            dist.barrier()                 # all ranks' x in place before anyone's kernel reads it
            out = band.aggregate(xl[:, :band.n_local] if False else xl, bias, relu=True)
            torch.cuda.synchronize()
            want = full[:, band.r0 * w:(band.r0 + band.rows) * w]
            same = torch.equal(out, want)
            ok &= same
            if not same:
                print("rank %d mesh %dx%d step %d MISMATCH max %g" % (rank, h, w, step, (out - want).abs().max().item()), flush=True)
            dist.barrier()
    # ---- timing at the bench shape (weak scaling: 582 x 390 rows per rank) ----------------------
    h, w, f = 582 * world, 390, 256
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    bias = torch.randn(f, device=dev)
    res = {}
    for name, cls in (("peer", partition.PeerMeshBand), ("nccl", partition.MeshBand)):
        band = cls(h, w, g.dis)
        xl = band.alloc(1, f, torch.float32, dev)
        band.owned(xl).normal_()
        out = torch.empty(1, band.n_own, f, device=dev)
        for _ in range(5):
            band.aggregate(xl, bias, out=out)
        torch.cuda.synchronize()
        dist.barrier()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        gobj = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            with torch.cuda.graph(gobj, stream=cap):
                band.aggregate(xl, bias, out=out)
        torch.cuda.current_stream().wait_stream(cap)
        for _ in range(5):
            gobj.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 300
        e0.record()
        for _ in range(steps):
            gobj.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name + "_us_per_step"] = round(ms.item() * 1e3, 2)
        dist.barrier()
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        res.update({"world": world, "bitwise_equal_to_single_gpu": bool(okt.item())})
        print(json.dumps(res), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0 if okt.item() else 1)


if __name__ == "__main__":
    main()
