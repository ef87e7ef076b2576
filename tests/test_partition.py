"""Host logic of the row-band partition, incl. a world_size-2 gloo run of the halo exchange.
The feature-row pack/unpack kernels are CUDA-only, so the gloo test substitutes a torch index
stand-in for them (test-only; the product has no CPU path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gwen_b200 import partition
from gwen_b200.graph import GraphCSR
from oracle import gcn_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_ranges():
    r = partition.band_ranges(10, 7, 4)
    assert [len(x) // 7 for x in r] == [3, 3, 2, 2]
    assert r[0].start == 0 and r[-1].stop == 70
    for a, b in zip(r, r[1:]):
        assert a.stop == b.start
    assert partition.band_ranges(582 * 8, 390, 8)[3] == range(3 * 582 * 390, 4 * 582 * 390)


def cpu_graph(h, w):
    ei = orc.grid(h, w)
    rowptr, src, perm, dis = orc.dst_sorted_csr(ei, h * w)
    _, ew, _ = orc.gcn_norm(ei, h * w, dis_mode="exact")
    t = torch.from_numpy
    return GraphCSR(t(rowptr), t(src), ew[t(perm)], t(dis), t(perm), h * w, h * w, len(src), 1)


def local_aggregate(lg, x_local):
    """CSR-order reference aggregate on the local graph (CPU, test only)."""
    g = lg.graph
    out = torch.zeros(lg.n_own, x_local.shape[-1])
    rp, src, w = g.rowptr.tolist(), g.src.tolist(), g.w
    for i in range(lg.n_own):
        acc = torch.zeros(x_local.shape[-1])
        for s in range(rp[i], rp[i + 1]):
            acc = acc + w[s] * x_local[src[s]]
        out[i] = acc
    return out


def test_partition_graph_slices_global_csr():
    h, w, world = 9, 5, 3
    g = cpu_graph(h, w)
    ranges = partition.band_ranges(h, w, world)
    x = torch.randn(h * w, 4)
    ei2, ew, _ = orc.gcn_norm(orc.grid(h, w), h * w, dis_mode="exact")
    ref = orc.propagate(x, ei2, ew, h * w)
    for p in range(world):
        lg = partition.partition_graph(g, ranges[p], (len(ranges[p]) // w, w))
        lo, hi = ranges[p].start, ranges[p].stop
        want_halo = list(range(max(lo - w, 0), lo)) + list(range(hi, min(hi + w, h * w)))
        assert lg.halo_ids.tolist() == want_halo
        assert lg.graph.n_dst == hi - lo and lg.graph.n_src == hi - lo + len(want_halo)
        x_local = torch.cat([x[lo:hi], x[lg.halo_ids]])
        assert torch.equal(local_aggregate(lg, x_local), ref[lo:hi])      # bitwise == unpartitioned


def _worker(rank, world, port, h, w, f, batch):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gwen_b200 import ops as gops
    gops.rows_gather = lambda x, idx, out=None: x[:, idx.long()].contiguous()

    def _scatter(x, idx, buf):
        x[:, idx.long()] = buf
        return x
    gops.rows_scatter_ = _scatter
    g = cpu_graph(h, w)
    ranges = partition.band_ranges(h, w, world)
    lg = partition.partition_graph(g, ranges[rank], (len(ranges[rank]) // w, w))
    hx = partition.HaloExchange(lg, ranges)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(batch, h * w, f, generator=gen)
    lo, hi = ranges[rank].start, ranges[rank].stop
    x_local = torch.zeros(batch, lg.n_local, f)
    x_local[:, :lg.n_own] = x[:, lo:hi]
    hx.exchange(x_local if batch > 1 else x_local[0])
    assert torch.equal(x_local[:, lg.n_own:], x[:, lg.halo_ids]), "halo rows differ"
    ei2, ew, _ = orc.gcn_norm(orc.grid(h, w), h * w, dis_mode="exact")
    for b in range(batch):
        ref = orc.propagate(x[b], ei2, ew, h * w)
        assert torch.equal(local_aggregate(lg, x_local[b]), ref[lo:hi])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 1), (2, 3), (3, 1)])
def test_halo_exchange_gloo(world, batch):
    port = 29500 + (os.getpid() % 2000) + world * 7 + batch
    mp.spawn(_worker, args=(world, port, 8, 6, 4, batch), nprocs=world, join=True)


def _band_worker(rank, world, port, h, w, f, batch):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, _, dis = orc.gcn_norm(orc.grid(h, w), h * w, dis_mode="exact")
    band = partition.MeshBand(h, w, dis)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(batch, h * w, f, generator=gen)
    xl = band.alloc(batch, f, torch.float32, "cpu")
    band.owned(xl).copy_(x[:, band.r0 * w:(band.r0 + band.rows) * w])
    band.exchange(xl)
    want = torch.zeros(batch, (band.rows + 2) * w, f)
    lo, hi = max(band.r0 - 1, 0), min(band.r0 + band.rows + 1, h)
    want[:, (lo - (band.r0 - 1)) * w:(hi - (band.r0 - 1)) * w] = x[:, lo * w:hi * w]
    assert torch.equal(xl, want), "band halo rows differ"
    # bordered dis: element [r+1][c+1] = global dis of mesh row r0-1+r (0 outside the mesh)
    d2 = dis.view(h, w)
    for r in range(band.rows + 2):
        gr = band.r0 - 1 + r
        exp = d2[gr] if 0 <= gr < h else torch.zeros(w)
        assert torch.equal(band.dis[r + 1, 1:w + 1], exp)
    assert band.dis[0].abs().sum() == 0 and band.dis[:, 0].abs().sum() == 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 1), (3, 2)])
def test_mesh_band_exchange_gloo(world, batch):
    port = 31500 + (os.getpid() % 2000) + world * 5 + batch
    mp.spawn(_band_worker, args=(world, port, 9, 6, 4, batch), nprocs=world, join=True)


def _eval_gather_worker(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gwen_b200.train import gather_eval_results
    # two kept predictions per rank (output[1] of two samples), a rank-dependent loss
    preds = [torch.full((5,), float(10 * rank + i)) for i in range(2)]
    res = gather_eval_results(torch.tensor(float(rank + 1)), preds)
    if rank == 0:
        loss, y = res
        assert abs(loss - sum(range(1, world + 1)) / world) < 1e-6
        # reference shapes (models_gnn.py:449,465,482): output[1] is 1-D [C], torch.cat(y_preds) is 1-D
        # [T * C], the rank-ordered cat is 1-D [world * T * C]
        want = torch.cat([torch.full((5,), float(10 * r + i)) for r in range(world) for i in range(2)])
        assert y.shape == (world * 2 * 5,) and torch.equal(y, want)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_eval_results_gloo(world):
    """Collective tail of the reference eval loop (models_gnn.py:452-489): losses averaged over the
    ranks, predictions concatenated in rank order, only rank 0 gets the result."""
    port = 29700 + (os.getpid() % 1500) + world
    mp.spawn(_eval_gather_worker, args=(world, port), nprocs=world, join=True)
