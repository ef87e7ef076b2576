"""What the peer-halo variant of the stencil kernel costs BY ITSELF: one rank (no neighbour to wait for, no NVLink
traffic), the cfg 2 band, PeerMeshBand.aggregate against the plain stencil on the same rows.  Separates the kernel's
own overhead (extra warp, boundary tiles last, cooperative launch, end-of-kernel re-arming) from the coupling of
the ranks at N >= 2.   torchrun --nproc-per-node 1 tools/peer_overhead.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwen_b200 as gw  # noqa: E402
from gwen_b200 import ops, partition  # noqa: E402

rank = int(os.environ.get("RANK", 0))
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=rank, world_size=int(os.environ.get("WORLD_SIZE", 1)), device_id=dev)
h, w, f = 582, 390, 256
n = h * w
g = gw.get_graph(gw.grid(h, w, dev), n)
band = partition.PeerMeshBand(h, w, g.dis)
bufs = [band.alloc(1, f, torch.float32, dev) for _ in range(3)]
outs = [torch.empty(1, n, f, device=dev) for _ in range(3)]
xs = [torch.randn(1, n, f, device=dev) for _ in range(3)]
for b_, x in zip(bufs, xs):
    band.owned(b_).copy_(x)


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


i = [0]


def peer():
    k = i[0] % 3
    i[0] += 1
    band.aggregate(bufs[k], out=outs[k])


def plain():
    k = i[0] % 3
    i[0] += 1
    ops.aggregate(g, xs[k], out=outs[k])


t_plain = timeit(plain)
t_peer = timeit(peer)
t_plain2 = timeit(plain)
ok = torch.equal(band.aggregate(bufs[0]), ops.aggregate(g, xs[0]))
print("plain stencil %.2f us | peer variant, one rank %.2f us | plain again %.2f us | bitwise equal %s" % (t_plain, t_peer, t_plain2, ok))
dist.destroy_process_group()
