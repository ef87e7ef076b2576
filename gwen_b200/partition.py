"""Row-band mesh partition (multi-GPU path): one process per GPU, ensemble members in the outer batch.

No reference counterpart: GWEN's GNN trainer spawns identical replicas (``models_gnn.py:341`` has
DistributedDataParallel commented out).  BASELINE.json's north_star asks for the mesh to be
spatially partitioned across the GPUs of one NVSwitch box; SURVEY.md section 8(e) fixes the scheme:
nodes are owned in contiguous id ranges (row bands of the H x W grid: node id = r*W + c) and a GCN
layer needs one halo row from each neighbouring band.  Three implementations, fastest first:

* ``PeerMeshBand`` -- plain mesh graphs.  Feature buffers live in CUDA symmetric memory; the halo rows
  are pulled from the neighbours' buffers over NVLink INSIDE the stencil aggregation kernel
  (``gwen_grid_stencil_peer_fwd``: device-side epoch flags, boundary tile rows last).  One launch per
  aggregation, no NCCL call.  ``BandGNNModel`` runs the six-layer model on it, forward and backward,
  with the weight gradients all-reduced (``allreduce_grads``) and the reference's masked L1 loss split
  into per-rank shares (``loss``).
* ``MeshBand`` -- same layout, halos by grouped NCCL ``isend/irecv`` on a side stream under the
  interior rows (gloo in the CPU tests), first / last owned rows in two extra launches.
* ``partition_graph`` + ``HaloExchange`` + ``BandAggregator`` -- arbitrary graphs: a rank's local graph
  is the slice of the GLOBAL destination-sorted CSR for its owned rows, so message order and weights
  (global degrees) are exactly the single-GPU ones; local source numbering is ``[owned | halo]`` with
  the halo sorted by global id and fetched with grouped send/recv before an aggregation.

All three give results bitwise equal to the un-partitioned kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import ctypes as C

import os

import torch
import torch.distributed as dist

from . import ops
from .graph import GraphCSR, bordered_dis

__all__ = ["band_ranges", "partition_graph", "HaloExchange", "LocalGraph", "BandAggregator", "MeshBand",
           "PeerMeshBand", "BandGNNModel"]


def band_ranges(height: int, width: int, world_size: int) -> List[range]:
    """Owned node-id range of every rank: grid rows split as evenly as possible, in rank order."""
    base, extra = divmod(height, world_size)
    out, r = [], 0
    for p in range(world_size):
        rows = base + (1 if p < extra else 0)
        out.append(range(r * width, (r + rows) * width))
        r += rows
    return out


class LocalGraph:
    """A rank's share of a partitioned graph: local CSR + who owns which halo row."""

    def __init__(self, graph: GraphCSR, n_own: int, halo_ids: torch.Tensor, own: range):
        self.graph, self.n_own, self.halo_ids, self.own = graph, n_own, halo_ids, own
        self.n_halo = int(halo_ids.numel())

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_halo


def partition_graph(global_graph: GraphCSR, own: range, grid_rows_width=None) -> LocalGraph:
    """Slice the global CSR to the destination rows in ``own`` and renumber sources locally.

    ``grid_rows_width=(rows, W)`` marks the owned band as a rows x W grid so 2-D tile plans apply.
    One-time setup on the device (index arithmetic only, no feature data).
    """
    g = global_graph
    lo, hi = own.start, own.stop
    rp = g.rowptr[lo:hi + 1]
    e0, e1 = int(rp[0].item()), int(rp[-1].item())
    src_g = g.src[e0:e1].to(torch.int64)
    remote = (src_g < lo) | (src_g >= hi)
    halo_ids = torch.unique(src_g[remote])  # sorted ascending
    local = torch.where(remote, (hi - lo) + torch.searchsorted(halo_ids, src_g), src_g - lo)
    n_own = hi - lo
    lg = GraphCSR((rp - rp[0]).contiguous(), local.to(torch.int32).contiguous(),
                  g.w[e0:e1].contiguous(), None, None, n_own, n_own + int(halo_ids.numel()),
                  e1 - e0, g.flags, edge_index=None, grid_shape=grid_rows_width)
    return LocalGraph(lg, n_own, halo_ids, own)


class HaloExchange:
    """Fetches halo rows from their owners: ``exchange(x_local)`` fills ``x_local[..., n_own:, :]``.

    Setup (collective): every rank announces the global ids it needs; each owner derives its send
    lists.  Per call: one ``rows_gather`` (pack) per peer, one grouped batch of isend/irecv, and
    for batched inputs one ``rows_scatter_`` (unpack) per peer.
    """

    def __init__(self, local: LocalGraph, ranges: Sequence[range], group=None):
        self.local, self.group = local, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        dev = local.halo_ids.device
        need = local.halo_ids.cpu().tolist()
        all_need: List[Optional[list]] = [None] * self.world
        dist.all_gather_object(all_need, need, group=group)
        own = ranges[self.rank]
        # what each peer wants from me (local row numbers), and where what I get from a peer lands
        self.send_idx, self.recv_slice = {}, {}
        for p in range(self.world):
            if p == self.rank:
                continue
            mine = [i - own.start for i in all_need[p] if i in own]
            if mine:
                self.send_idx[p] = torch.tensor(mine, dtype=torch.int32, device=dev)
            theirs = [k for k, i in enumerate(need) if i in ranges[p]]
            if theirs:
                assert theirs == list(range(theirs[0], theirs[-1] + 1))  # halo sorted by owner
                self.recv_slice[p] = (local.n_own + theirs[0], len(theirs))

    def exchange(self, x_local: torch.Tensor) -> torch.Tensor:
        """x_local: [B, n_local, F] (or [n_local, F]) contiguous; rows [:n_own] valid on entry."""
        x3 = x_local if x_local.dim() == 3 else x_local.unsqueeze(0)
        assert x3.is_contiguous() and x3.shape[1] == self.local.n_local
        b = x3.shape[0]
        p2p, recv_tmp = [], {}
        for p, idx in sorted(self.send_idx.items()):
            buf = ops.rows_gather(x3, idx)
            p2p.append(dist.P2POp(dist.isend, buf, p, self.group))
        for p, (start, cnt) in sorted(self.recv_slice.items()):
            if b == 1:
                dst = x3[0, start:start + cnt]  # contiguous: receive straight into place
            else:
                dst = recv_tmp[p] = torch.empty((b, cnt, x3.shape[2]), dtype=x3.dtype, device=x3.device)
            p2p.append(dist.P2POp(dist.irecv, dst, p, self.group))
        if p2p:
            for req in dist.batch_isend_irecv(p2p):
                req.wait()
        for p, tmp in recv_tmp.items():
            start, cnt = self.recv_slice[p]
            idx = torch.arange(start, start + cnt, dtype=torch.int32, device=x3.device)
            ops.rows_scatter_(x3, idx, tmp)
        return x_local


class BandAggregator:
    """Aggregation over a rank's row band with the halo exchange hidden behind interior work.

    Only the first and last tile rows of the band read halo rows.  The band's tile plan lists
    those boundary tiles first, so a step is: (1) halo exchange issued on a side stream,
    (2) ONE launch over the interior tiles on the caller's stream -- a few SMs are left free
    (``sm_reserve``) so the pack and NCCL kernels can run beside the persistent kernel --
    (3) the caller's stream waits for the exchange and runs ONE launch over the boundary tiles.
    """

    def __init__(self, local: LocalGraph, hx: HaloExchange, tile=(8, 16), sm_reserve: int = 8,
                 strip_width: int = 32):
        if local.graph.grid_shape is None:
            raise RuntimeError("BandAggregator needs a grid-shaped band (grid_rows_width)")
        self.local, self.hx, self.tile, self.sm_reserve = local, hx, tuple(tile), sm_reserve
        g = local.graph
        rows, width = g.grid_shape
        dev = g.device
        th, tw = self.tile

        def region(r0, r1, rth, rtw):
            """(order, tile sizes) of the rth x rtw tiles covering grid rows [r0, r1), row-major."""
            if r1 <= r0:
                return (torch.empty(0, dtype=torch.int64, device=dev),) * 2
            r = torch.arange(r0, r1, device=dev).view(-1, 1).expand(-1, width).reshape(-1)
            c = torch.arange(width, device=dev).repeat(r1 - r0)
            tcols = -(-width // rtw)
            tid = ((r - r0) // rth) * tcols + c // rtw
            key = tid * (rth * rtw) + ((r - r0) % rth) * rtw + c % rtw
            perm = torch.argsort(key)
            return (r * width + c)[perm], torch.bincount(tid)

        # boundary: the first and last grid row of the band as 1 x strip_width strips (the only
        # destinations whose sources include halo rows); interior: everything between them.
        top, top_sz = region(0, min(1, rows), 1, strip_width)
        bot, bot_sz = region(max(rows - 1, 1), rows, 1, strip_width)
        mid, mid_sz = region(1, rows - 1, th, tw)
        order = torch.cat([top, bot, mid]).to(torch.int32).contiguous()
        sizes = torch.cat([top_sz, bot_sz, mid_sz])
        tile_ptr = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=dev)
        tile_ptr[1:] = torch.cumsum(sizes, 0)
        assert order.numel() == g.n_dst and int(tile_ptr[-1]) == g.n_dst
        self.n_boundary = int(top_sz.numel() + bot_sz.numel())
        self.n_interior = int(mid_sz.numel())
        self.plan = g.tile_plan_from(order, tile_ptr.to(torch.int32), tw + 2,
                                     ("band", th, tw, strip_width))
        self.side = torch.cuda.Stream(device=dev)

    def __call__(self, x_local: torch.Tensor, bias=None, relu: bool = False, out=None) -> torch.Tensor:
        from ._lib import lib
        g = self.local.graph
        cur = torch.cuda.current_stream(x_local.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            self.hx.exchange(x_local)
            done = torch.cuda.Event()
            done.record(self.side)
        if out is None:
            shape = tuple(x_local.shape[:-2]) + (g.n_dst, x_local.shape[-1])
            out = torch.empty(shape, dtype=x_local.dtype, device=x_local.device)
        kw = dict(bias=bias, relu=relu, kernel="tiled", out=out, plan=self.plan)
        if self.n_interior:
            old = lib().gwen_set_sm_reserve(self.sm_reserve)
            try:
                ops.aggregate(g, x_local, tile_range=(self.n_boundary, self.n_interior), **kw)
            finally:
                lib().gwen_set_sm_reserve(old)
        cur.wait_event(done)
        ops.aggregate(g, x_local, tile_range=(0, self.n_boundary), **kw)
        return out


class MeshBand:
    """A rank's row band of a plain H x W mesh for the stencil fast path.

    Local feature layout: ``[B, (rows + 2) * W, F]`` = ``[top halo row | owned rows | bottom halo
    row]`` in mesh order, so every halo is one contiguous row and (for B = 1) is sent and received
    in place, without pack kernels.  Halo rows outside the mesh stay zero and carry ``dis = 0``.
    ``aggregate`` overlaps the exchange with the interior rows: exchange on a side stream, interior
    destination rows [1, rows-1) on the caller's stream (``sm_reserve`` SMs left free for the NCCL
    kernels), then the first and last owned rows once the halos have arrived.
    Results are bitwise equal to the single-GPU stencil kernel on the whole mesh.
    """

    def __init__(self, height: int, width: int, dis_global: torch.Tensor, group=None, sm_reserve: int = 8,
                 rank: Optional[int] = None, world: Optional[int] = None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.H, self.W, self.sm_reserve = height, width, sm_reserve
        ranges = band_ranges(height, width, self.world)
        own = ranges[self.rank]
        self.r0, self.rows = own.start // width, len(own) // width
        dev = dis_global.device
        d2 = dis_global.view(height, width)
        local = torch.zeros((self.rows + 2, width), dtype=torch.float32, device=dev)
        lo, hi = max(self.r0 - 1, 0), min(self.r0 + self.rows + 1, height)
        local[lo - (self.r0 - 1):hi - (self.r0 - 1)] = d2[lo:hi]
        self.dis = bordered_dis(local)
        self.up = self.rank - 1 if self.rank > 0 else None
        self.down = self.rank + 1 if self.rank < self.world - 1 else None
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    @property
    def n_own(self) -> int:
        return self.rows * self.W

    @property
    def n_local(self) -> int:
        return (self.rows + 2) * self.W

    def alloc(self, batch: int, feat: int, dtype, device) -> torch.Tensor:
        """Zero-initialised local feature buffer (halo rows outside the mesh must stay zero)."""
        return torch.zeros((batch, self.n_local, feat), dtype=dtype, device=device)

    def owned(self, x_local: torch.Tensor) -> torch.Tensor:
        return x_local[..., self.W:self.W + self.n_own, :]

    def exchange(self, x_local: torch.Tensor) -> None:
        """Fill the halo rows from the neighbouring ranks (x_local [B, n_local, F] contiguous)."""
        w, rows = self.W, self.rows
        x3 = x_local if x_local.dim() == 3 else x_local.unsqueeze(0)
        first, last = slice(w, 2 * w), slice(rows * w, (rows + 1) * w)
        top, bot = slice(0, w), slice((rows + 1) * w, (rows + 2) * w)
        p2p, post = [], []
        for peer, snd, rcv in ((self.up, first, top), (self.down, last, bot)):
            if peer is None:
                continue
            if x3.shape[0] == 1:
                sbuf, rbuf = x3[0, snd], x3[0, rcv]       # contiguous rows: in place
            else:
                sbuf = x3[:, snd].contiguous()
                rbuf = torch.empty_like(sbuf)
                post.append((rcv, rbuf))
            p2p.append(dist.P2POp(dist.isend, sbuf, peer, self.group))
            p2p.append(dist.P2POp(dist.irecv, rbuf, peer, self.group))
        if p2p:
            for req in dist.batch_isend_irecv(p2p):
                req.wait()
        for rcv, rbuf in post:
            x3[:, rcv] = rbuf

    def aggregate(self, x_local: torch.Tensor, bias=None, relu: bool = False, out=None,
                  overlap: bool = True) -> torch.Tensor:
        from ._lib import lib
        w, rows = self.W, self.rows
        x3 = x_local if x_local.dim() == 3 else x_local.unsqueeze(0)
        b, _, f = x3.shape
        if out is None:
            out = torch.empty((b, self.n_own, f), dtype=x3.dtype, device=x3.device)
        o3 = out if out.dim() == 3 else out.unsqueeze(0)
        kw = dict(bias=bias, relu=relu)
        if not overlap or rows < 3:
            self.exchange(x3)
            ops.mesh_stencil(x3, self.dis, rows + 2, rows, w, 1, out=o3, **kw)
            return out
        cur = torch.cuda.current_stream(x3.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            self.exchange(x3)
            done = torch.cuda.Event()
            done.record(self.side)
        old = lib().gwen_set_sm_reserve(self.sm_reserve)
        try:   # interior destination rows 1 .. rows-2 read owned rows only
            ops.mesh_stencil(x3, self.dis, rows + 2, rows - 2, w, 2, out=o3[:, w:(rows - 1) * w], **kw)
        finally:
            lib().gwen_set_sm_reserve(old)
        cur.wait_event(done)
        ops.mesh_stencil(x3, self.dis, rows + 2, 1, w, 1, out=o3[:, :w], **kw)
        ops.mesh_stencil(x3, self.dis, rows + 2, 1, w, rows, out=o3[:, (rows - 1) * w:], **kw)
        return out


class PeerMeshBand(MeshBand):
    """:class:`MeshBand` whose halo exchange happens INSIDE the aggregation kernel, over NVLink peer
    memory (``gwen_grid_stencil_peer_fwd``): no NCCL call, no pack kernels, one launch per step.

    Feature buffers come from ``alloc`` (CUDA symmetric memory, ``torch.distributed._symmetric_memory``):
    every rank allocates the same shapes in the same order, which gives each rank device pointers
    to its neighbours' buffers.  The neighbours' boundary rows are pulled by one extra warp per CTA
    while the interior tile rows are aggregated; flags in symmetric memory order the exchange
    (see include/gwen_b200.h).  Results are bitwise equal to ``MeshBand.aggregate``.
    """

    def __init__(self, height: int, width: int, dis_global: torch.Tensor, group=None,
                 rank: Optional[int] = None, world: Optional[int] = None):
        super().__init__(height, width, dis_global, group=group, sm_reserve=0, rank=rank, world=world)
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self._group = group if group is not None else dist.group.WORLD
        dev = dis_global.device
        # rows owned by every rank (the neighbours' layouts differ by at most one row)
        self._rows_of = [len(r) // width for r in band_ranges(height, width, self.world)]
        self._ctl = symm.empty(64, dtype=torch.int32, device=dev)
        self._ctl.zero_()
        self._ctl_hdl = symm.rendezvous(self._ctl, self._group)
        self._bufs = {}
        torch.cuda.synchronize(dev)
        dist.barrier(self._group)

    def alloc(self, batch: int, feat: int, dtype, device) -> torch.Tensor:
        """Zero-initialised SYMMETRIC local feature buffer [batch, n_local_max, feat] (collective:
        all ranks call it with the same arguments).  Sized for the tallest band so that every rank's
        allocation has the same shape; use ``[:, :n_local]``."""
        rows_max = max(self._rows_of)
        t = self._symm.empty((batch, (rows_max + 2) * self.W, feat), dtype=dtype, device=device)
        t.zero_()
        hdl = self._symm.rendezvous(t, self._group)
        self._bufs[t.data_ptr()] = (t, hdl)
        torch.cuda.synchronize(device)
        dist.barrier(self._group)
        return t

    def error_word(self) -> int:
        """The protocol's sticky error word (``ctl[5]``, see include/gwen_b200.h): 0 = every halo wait so
        far completed; bit 0 = a neighbour never announced an epoch within GWEN_PEER_TIMEOUT_MS, bit 1 =
        a CTA of a launch never published its halo share.  Synchronises the device."""
        return int(self._ctl[5].item())

    def check(self) -> None:
        """Raise if a peer launch timed out waiting for a neighbour (results are then undefined)."""
        e = self.error_word()
        if e:
            raise RuntimeError("gwen_b200: peer halo exchange timed out (error word %d): the ranks did not "
                               "launch the same sequence of peer aggregations, or a neighbour died" % e)

    def _peers(self, x3: torch.Tensor):
        from . import _lib
        t, hdl = self._bufs[x3.data_ptr()]
        esz, w, f = x3.element_size(), self.W, x3.shape[-1]
        bstride = t.stride(0)
        ctl = self._ctl_hdl.buffer_ptrs
        up_row = down_row = up_flag = down_flag = None
        if self.up is not None:      # its last owned row = local row rows_up
            up_row = hdl.buffer_ptrs[self.up] + self._rows_of[self.up] * w * f * esz
            up_flag = ctl[self.up] + 4 * 1          # I am its DOWN neighbour
        if self.down is not None:    # its first owned row = local row 1
            down_row = hdl.buffer_ptrs[self.down] + w * f * esz
            down_flag = ctl[self.down] + 4 * 0      # I am its UP neighbour
        return _lib.HaloPeersStruct(up_row, down_row, bstride, bstride, up_flag, down_flag,
                                    self._ctl.data_ptr())

    def aggregate(self, x_local: torch.Tensor, bias=None, relu: bool = False, out=None,
                  overlap: bool = True) -> torch.Tensor:
        from . import _lib
        from ._lib import check, lib
        x3 = x_local if x_local.dim() == 3 else x_local.unsqueeze(0)
        if x3.data_ptr() not in self._bufs:
            raise RuntimeError("PeerMeshBand.aggregate needs a buffer from PeerMeshBand.alloc (symmetric memory)")
        b, _, f = x3.shape
        if out is None:
            out = torch.empty((b, self.n_own, f), dtype=x3.dtype, device=x3.device)
        o3 = out if out.dim() == 3 else out.unsqueeze(0)
        peers = self._peers(x3)
        bias32 = None if bias is None else bias.detach().to(torch.float32).contiguous()
        with torch.cuda.device(x3.device):
            check(lib().gwen_grid_stencil_peer_fwd(
                x3.data_ptr(), o3.data_ptr(), self.dis.data_ptr(), self.dis.shape[1], self.dis.shape[0], b, self.rows, self.W, f,
                x3.stride(0), f, o3.stride(0) if b > 1 else self.n_own * f, ops.dtype_code(x3.dtype),
                None if bias32 is None else bias32.data_ptr(), _lib.EPI_RELU if relu else 0, 0, 0,
                C.byref(peers), torch.cuda.current_stream().cuda_stream), "gwen_grid_stencil_peer_fwd")
        return out


# ---------------------------------------------------------------------------------------------
# the six-layer model on a row band (forward + backward), SURVEY.md section 8(e)
# ---------------------------------------------------------------------------------------------
class _BandGCNFn(torch.autograd.Function):
    """One GCN layer on this rank's band: ``_GCNConvFn`` with the aggregation done by the band
    (halo exchange inside the kernel).  The tensor that gets aggregated lives in a symmetric buffer
    of ``net``; producers write into it in place where they can.  Backward: ``A_hat^T = A_hat`` on the
    mesh, so the gradient aggregation is the same exchange + stencil on gradient buffers."""

    @staticmethod
    def forward(ctx, x, weight, bias, net, li, relu, agg_first):
        band = net.band
        b = x.shape[0]
        if agg_first:      # (A_hat x) W^T
            xs = net.stage(("f", li), x)
            h = band.aggregate(xs)[:, :band.n_own]
            y = ops.linear(h, weight, bias, relu, out=net.out_view(li, b, weight.shape[0], x.dtype))
            saved_in = h
        else:              # A_hat (x W^T)
            hs = net.buffer(("f", li), b, weight.shape[0], x.dtype)
            ops.linear(x, weight, out=band.owned(hs))
            y = band.aggregate(hs, bias, relu, out=net.out_view(li, b, weight.shape[0], x.dtype))
            saved_in = x
        ctx.net, ctx.li, ctx.relu, ctx.agg_first, ctx.has_bias = net, li, relu, agg_first, bias is not None
        ctx.save_for_backward(saved_in, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        saved_in, weight, y = ctx.saved_tensors
        net, li, band = ctx.net, ctx.li, ctx.net.band
        b = dy.shape[0]
        gs = None
        if not ctx.agg_first:   # dz is what gets aggregated: write it straight into its band buffer
            gs = net.buffer(("b", li), b, weight.shape[0], dy.dtype)
        # the next layer's dgrad epilogue may already have applied this layer's ReLU mask (see nn.ReluLink): then only
        # the bias-gradient sums (and the copy into the band buffer) remain
        masked = ctx.relu and net._masked.pop(li, False)
        from . import nn as _nn
        # nothing left to mask and dz goes straight into the wgrad: db from the wgrad's own pass (ones-column MMA)
        db_from_wgrad = ctx.agg_first and ctx.has_bias and (masked or not ctx.relu) and _nn.BWD_MASK_FUSION \
            and dy.dtype == torch.bfloat16
        if db_from_wgrad:
            dz, db = dy, None
        else:
            dz, db = ops.relu_bias_bwd(dy, y if (ctx.relu and not masked) else None, ctx.has_bias,
                                       out=None if gs is None else band.owned(gs))
        need_dx = ctx.needs_input_grad[0]
        dx = None
        # weight gradient: written straight into this layer's slice of the flat fp32 gradient buffer and
        # all-reduced from there while the remaining layers' backward runs (net.overlap_grads)
        if net.overlap_grads and li in net._pending:   # an earlier backward's bucket is still pending: deliver it first
            if net.bucket_mode == "layer":
                net._deliver(li)
            else:
                net.allreduce_grads()
        wv = net.grad_views(li)[0] if net.overlap_grads else None
        if ctx.agg_first:
            dw = None
            if db_from_wgrad:
                both = ops.linear_bwd_weight_bias(dz, saved_in, out=wv)
                if both is None:
                    dz, db = ops.relu_bias_bwd(dy, None, True)
                else:
                    dw, db = both
            if dw is None:
                dw = ops.linear_bwd_weight(dz, saved_in, out=wv)      # dW = dz^T (A_hat x)
        else:
            dh = band.aggregate(gs)                                   # A_hat^T dz
            dw = ops.linear_bwd_weight(dh, saved_in, out=wv)          # dW = dh^T x
        if net.overlap_grads:
            net.launch_bucket(li, db)
            dw = db = None              # delivered by allreduce_grads(), already summed over the ranks
        else:
            dw = dw.to(weight.dtype)
            if db is not None:
                db = db.to(weight.dtype)
        if need_dx:
            if ctx.agg_first:
                gs = net.buffer(("b", li), b, weight.shape[1], dz.dtype)
                ops.linear_bwd_data(dz, weight, out=band.owned(gs))
                dx = band.aggregate(gs)
            else:
                if li > 0 and net.layers[li - 1][1] and _nn.BWD_MASK_FUSION:   # x = relu(layer li - 1): its mask here
                    dst = None
                    if not net._agg_first[li - 1]:    # that layer aggregates its gradient: write it where it is staged
                        dst = band.owned(net.buffer(("b", li - 1), b, weight.shape[1], dh.dtype))
                    dx = ops.linear_bwd_data_masked(dh, weight, saved_in, out=dst)
                    if dx is not None:
                        net._masked[li - 1] = True
                if dx is None:
                    dx = ops.linear_bwd_data(dh, weight)
        return dx, dw, db, None, None, None, None


class BandGNNModel(torch.nn.Module):
    """:class:`gwen_b200.models_gnn.GNNModel` evaluated on one row band per rank (mesh partitioned
    over the GPUs, ensemble members in the outer batch).  Shares the wrapped model's parameters, so
    ``state_dict`` and optimizers are the plain model's.  ``forward(x_own [B, n_own, C]) -> [B, n_own, C]``;
    after ``backward`` call :meth:`allreduce_grads` (the weight / bias gradients of the ranks are partial
    sums over their own rows).  Outputs are bitwise equal to the un-partitioned model's rows.

    ``overlap_grads`` (default): every layer's backward writes its fp32 ``dW`` / ``db`` into that layer's
    slice of ONE pre-flattened gradient buffer; ``allreduce_grads`` reduces the buffer in place with one NCCL
    all-reduce and hands the (summed) slices to ``p.grad`` -- no ``torch.cat``, no copy-back.  With
    ``GWEN_GRAD_BUCKETS=layer`` each slice's all-reduce starts inside that layer's backward instead and runs under
    the backward of the earlier layers (``allreduce_grads`` then only waits for the handles).
    With ``overlap_grads=False`` autograd delivers the local gradients and ``allreduce_grads`` reduces
    them in one flat all-reduce afterwards."""

    def __init__(self, model, band: "PeerMeshBand", overlap_grads: bool = True):
        super().__init__()
        self.model, self.band, self.overlap_grads = model, band, overlap_grads
        # "flat" (default): ONE in-place all-reduce of the flat gradient buffer after backward -- nothing shares the SMs
        # with the persistent kernels meanwhile; "layer": a bucket's all-reduce starts inside that layer's backward and
        # runs under the remaining layers (measured: the NCCL kernels then take SMs from the 148-CTA kernels beside
        # them -- cfg 5 shape, 21 members at N = 2: 138.5 ms vs 132.6-133.9 flat; N = 8: cfg 4 8.05 vs 7.77 ms)
        self.bucket_mode = os.environ.get("GWEN_GRAD_BUCKETS", "flat")
        self._flat = None
        self._views = {}
        self._pending = {}
        self._masked = {}       # li -> True: layer li's ReLU mask was applied by layer li + 1's dgrad epilogue
        d, u = model.conv_layers.down_conv_layers, model.conv_layers.up_conv_layers
        self.layers = [(d.conv1, True), (d.conv2, True), (d.conv3, True),
                       (u.upconv3, True), (u.upconv4, True), (u.upconv5, False)]
        self._agg_first = [c.in_channels < c.out_channels for c, _ in self.layers]
        self._sym = {}

    # -- symmetric staging buffers (allocated collectively, in first-use order, once) -------------
    def buffer(self, key, batch, feat, dtype) -> torch.Tensor:
        k = (key, batch, feat, dtype)
        if k not in self._sym:
            self._sym[k] = self.band.alloc(batch, feat, dtype, self.band.dis.device)
        return self._sym[k]

    def stage(self, key, x: torch.Tensor) -> torch.Tensor:
        """The symmetric buffer for ``key`` holding ``x`` in its owned rows (no copy when the producer
        already wrote there through :meth:`out_view`)."""
        t = self.buffer(key, x.shape[0], x.shape[-1], x.dtype)
        own = self.band.owned(t)
        if not (x.data_ptr() == own.data_ptr() and x.stride() == own.stride()):
            ops.copy_rows_(own, x)
        return t

    def out_view(self, li: int, batch: int, feat: int, dtype):
        """Where layer ``li`` should write its output: the owned rows of the next layer's staging
        buffer when that layer aggregates first, else None (a fresh tensor)."""
        if li + 1 < len(self.layers) and self._agg_first[li + 1]:
            return self.band.owned(self.buffer(("f", li + 1), batch, feat, dtype))
        return None

    def forward(self, x_own: torch.Tensor) -> torch.Tensor:
        x = x_own if x_own.dim() == 3 else x_own.unsqueeze(0)
        for li, (conv, relu) in enumerate(self.layers):
            x = _BandGCNFn.apply(x, conv.lin.weight, conv.bias, self, li, relu, self._agg_first[li])
        return x if x_own.dim() == 3 else x[0]

    def loss(self, y_own: torch.Tensor, target_own: torch.Tensor, mask_own: torch.Tensor) -> torch.Tensor:
        """This rank's share of the reference ``loss_func`` (masked L1, ``models_gnn.py:261-265``) over the
        WHOLE mesh: local masked sum / (B x GLOBAL count x C), so the shares of all ranks add up to the
        global masked mean and ``backward`` + :meth:`allreduce_grads` gives its gradient.  A band without
        masked nodes contributes 0 (not 0/0).  The global count is all-reduced once per mask tensor and
        cached (the cache holds the tensor, so a recycled address cannot alias a stale count)."""
        from .train import masked_l1_loss
        cached = getattr(self, "_mask_cache", None)
        if cached is None or cached[0] is not mask_own or cached[1] != mask_own._version:
            tot = mask_own.sum().to(torch.float64).reshape(1)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.band.group)
            self._mask_cache = (mask_own, mask_own._version, tot.to(torch.float32))
        return masked_l1_loss(y_own, target_own, mask_own, count=self._mask_cache[2])

    # -- flat gradient buffer: [dW_0 | db_0 | dW_1 | db_1 | ...] in fp32, one bucket per live layer ------
    def grad_views(self, li: int):
        """(dW view [out, in], db view [out] or None, bucket view) of layer ``li`` in the flat buffer."""
        if self._flat is None:
            dev = self.band.dis.device
            sizes = [c.lin.weight.numel() + (c.bias.numel() if c.bias is not None else 0) for c, _ in self.layers]
            self._flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            off = 0
            for i, (c, _) in enumerate(self.layers):
                nw = c.lin.weight.numel()
                wv = self._flat[off:off + nw].view_as(c.lin.weight)
                bv = self._flat[off + nw:off + sizes[i]] if c.bias is not None else None
                self._views[i] = (wv, bv, self._flat[off:off + sizes[i]])
                off += sizes[i]
        return self._views[li]

    def launch_bucket(self, li: int, db) -> None:
        """Called from layer ``li``'s backward once its dW sits in the flat buffer: add db, start the
        asynchronous all-reduce of the bucket (NCCL's stream waits for the current stream's work so far;
        the current stream goes on with the next layer's backward)."""
        wv, bv, bucket = self.grad_views(li)
        if bv is not None:
            if db is not None:
                bv.copy_(db)
            else:
                bv.zero_()
        if self.bucket_mode == "layer":
            self._pending[li] = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.band.group, async_op=True)
        else:               # "flat": one all-reduce of the whole flat buffer in allreduce_grads()
            self._pending[li] = None

    def _deliver(self, li: int) -> None:
        """Wait for layer ``li``'s bucket and add it to the parameters' ``.grad``."""
        if li not in self._pending:
            return
        work = self._pending.pop(li)
        if work is not None:
            work.wait()
        conv = self.layers[li][0]
        wv, bv, _ = self._views[li]
        for p, v in ((conv.lin.weight, wv), (conv.bias, bv)):
            if p is None or v is None or not p.requires_grad:
                continue
            g = v.to(p.dtype, copy=True)
            p.grad = g if p.grad is None else p.grad.add_(g)

    def allreduce_grads(self) -> None:
        """Make ``p.grad`` the gradient summed over the ranks.  ``overlap_grads``: the per-layer
        all-reduces were started inside backward; wait for them and deliver the slices.  Otherwise: one
        flat NCCL all-reduce of the local gradients autograd delivered."""
        if self.overlap_grads:
            if self.bucket_mode != "layer" and self._pending:
                dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.band.group)
            for li in sorted(self._pending, reverse=True):   # completion order: last layer first
                self._deliver(li)
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1).float() for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.band.group)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
