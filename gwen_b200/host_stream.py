"""Message passing on HOST-resident features: chunked, double-buffered H2D -> aggregate -> D2H.

The reference keeps node features on the host and copies them to the GPU every iteration
(``batch.to(device)``, ``src/gwen/models_gnn.py:359-362``).  With the aggregation itself at ~85 us
for a COSMO-2E mesh, an end-to-end call is PCIe time; this module makes it ONE PCIe transfer time
instead of two: the mesh is cut into row chunks, chunk k is aggregated (``gwen_grid_stencil_fwd`` on
the rows it owns, reading one extra source row above and below) as soon as its source rows have
landed, and its result goes back while chunk k+1 is still arriving (PCIe is full duplex).  Two
device buffer sets let consecutive calls overlap as well.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .graph import GraphCSR

__all__ = ["HostPropagator", "HostBandPropagator", "pinned_near_gpu", "gpu_numa_cpus"]


class HostPropagator:
    """``out_host = A_hat x_host (+ bias, relu)`` for a plain H x W mesh graph, ``x_host`` /
    ``out_host`` pinned ``[N, F]`` host tensors.  Asynchronous: the caller's current stream waits for
    the result (synchronise it, or the returned event, before reading ``out_host``)."""

    def __init__(self, graph: GraphCSR, feat: int, dtype=torch.float32, chunks: int = 8):
        if not graph.is_plain_mesh:
            raise RuntimeError("HostPropagator needs a plain H x W mesh graph")
        self.graph, self.feat, self.dtype = graph, feat, dtype
        self.h, self.w = graph.grid_shape
        dev = graph.device
        self.dev = dev
        chunks = max(1, min(chunks, self.h))
        edges = [self.h * k // chunks for k in range(chunks + 1)]
        self.bounds = [(edges[k], edges[k + 1]) for k in range(chunks) if edges[k + 1] > edges[k]]
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
            self.x = [torch.empty((self.h * self.w, feat), dtype=dtype, device=dev) for _ in range(2)]
            self.y = [torch.empty((self.h * self.w, feat), dtype=dtype, device=dev) for _ in range(2)]
        self.dis = graph.dis_padded()
        self.free = [None, None]   # event: the buffer set's previous call has finished (its D2H is done)
        self.turn = 0

    def __call__(self, x_host: torch.Tensor, out_host: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 relu: bool = False) -> torch.cuda.Event:
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("HostPropagator needs pinned host tensors")
        if x_host.shape != (self.h * self.w, self.feat) or out_host.shape != x_host.shape or x_host.dtype != self.dtype:
            raise ValueError("x_host / out_host must be [%d, %d] %s" % (self.h * self.w, self.feat, self.dtype))
        p = self.turn
        self.turn ^= 1
        x, y, w = self.x[p], self.y[p], self.w
        with torch.cuda.device(self.dev):
            if self.free[p] is not None:          # WAR on this buffer set (two calls ago)
                self.s_in.wait_event(self.free[p])
            top = 0                                # source rows [0, top) are on the device
            for (a, b) in self.bounds:
                need = min(self.h, b + 1)          # chunk rows [a, b) read source rows [a - 1, b + 1)
                with torch.cuda.stream(self.s_in):
                    if need > top:
                        x[top * w:need * w].copy_(x_host[top * w:need * w], non_blocking=True)
                        top = need
                    landed = torch.cuda.Event()
                    landed.record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(landed)
                    ops.mesh_stencil(x, self.dis, self.h, b - a, w, a, bias=bias, relu=relu,
                                     out=y[a * w:b * w].unsqueeze(0))
                    ran = torch.cuda.Event()
                    ran.record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ran)
                    out_host[a * w:b * w].copy_(y[a * w:b * w], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.s_out)
            self.free[p] = done
            torch.cuda.current_stream(self.dev).wait_event(done)
        return done


# ---------------------------------------------------------------------------------------------
# the same pipeline on one rank's row band of a partitioned mesh (multi-GPU end to end)
# ---------------------------------------------------------------------------------------------
def gpu_numa_cpus(device_index: int):
    """CPU ids of the NUMA node the GPU's PCIe root hangs off (None if sysfs does not say)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        return node, cpus
    except Exception:  # noqa: BLE001  (no sysfs / no NUMA: nothing to bind)
        return None


def pinned_near_gpu(shape, dtype, device_index: int) -> torch.Tensor:
    """Pinned host tensor whose pages sit on the GPU's NUMA node: the calling thread is bound to that
    node's cores while the buffer is allocated and first touched (Linux places pages on the toucher's
    node), then the previous affinity is restored.  Without sysfs NUMA information this is a plain
    ``pin_memory()``."""
    import os
    info = None if os.environ.get("GWEN_NO_NUMA") else gpu_numa_cpus(device_index)
    old = None
    if info is not None and hasattr(os, "sched_setaffinity"):
        try:
            old = os.sched_getaffinity(0)
            allowed = sorted(set(info[1]) & old) or sorted(old)
            os.sched_setaffinity(0, allowed)
        except OSError:
            old = None
    try:
        t = torch.empty(shape, dtype=dtype).pin_memory()
        t.zero_()
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
    return t


class HostBandPropagator:
    """``out_host = (A_hat x)[own rows]`` for one rank's row band of a partitioned plain mesh, with
    ``x_host`` / ``out_host`` pinned ``[n_own, F]`` host tensors (BASELINE's multi-GPU configs measured end
    to end).  Same pipeline as :class:`HostPropagator`: the band is cut into row chunks, chunk k is
    aggregated with a sub-range launch of ``gwen_grid_stencil_fwd`` as soon as its source rows have
    landed and goes back while chunk k + 1 is still arriving.  Only the band's FIRST and LAST owned row
    need the neighbours' rows: they come out of one peer launch (``PeerMeshBand.aggregate``, the halo
    rows pulled over NVLink inside the kernel) once the whole band is on the device -- ~80 us of kernel
    for two rows, against milliseconds of PCIe.  Symmetric device buffers rotate over ``sets`` calls, so
    a neighbour still reading this rank's boundary rows never sees the next call's H2D."""

    def __init__(self, band, feat: int, dtype=torch.float32, chunks: int = 8, sets: int = 3):
        self.band, self.feat, self.dtype = band, feat, dtype
        dev = band.dis.device
        self.dev = dev
        rows, w = band.rows, band.W
        chunks = max(1, min(chunks, max(1, rows - 2)))
        # interior destination rows [1, rows - 1) (local owned-row numbering) in `chunks` pieces
        edges = [1 + (rows - 2) * k // chunks for k in range(chunks + 1)] if rows > 2 else [1, 1]
        self.bounds = [(edges[k], edges[k + 1]) for k in range(len(edges) - 1) if edges[k + 1] > edges[k]]
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
            self.x = [band.alloc(1, feat, dtype, dev) for _ in range(sets)]          # collective
            self.y = [torch.empty((band.n_own, feat), dtype=dtype, device=dev) for _ in range(sets)]
            self.edge = [torch.empty((1, band.n_own, feat), dtype=dtype, device=dev) for _ in range(sets)]
        self.free = [None] * sets
        self.turn = 0

    def __call__(self, x_host: torch.Tensor, out_host: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 relu: bool = False) -> torch.cuda.Event:
        band, w, rows = self.band, self.band.W, self.band.rows
        n_own = band.n_own
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("HostBandPropagator needs pinned host tensors")
        if x_host.shape != (n_own, self.feat) or out_host.shape != x_host.shape or x_host.dtype != self.dtype:
            raise ValueError("x_host / out_host must be [%d, %d] %s" % (n_own, self.feat, self.dtype))
        p = self.turn
        self.turn = (self.turn + 1) % len(self.x)
        xb, y, edge = self.x[p], self.y[p], self.edge[p]
        xo = band.owned(xb[0])                    # [n_own, F] view: local rows 1 .. rows of the band buffer
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            start = torch.cuda.Event()
            start.record(cur)
            self.s_in.wait_event(start)
            if self.free[p] is not None:
                self.s_in.wait_event(self.free[p])
            top = 0                                # owned rows [0, top) are on the device
            for (a, b) in self.bounds:             # destination owned rows [a, b) read owned rows [a - 1, b + 1)
                need = min(rows, b + 1)
                with torch.cuda.stream(self.s_in):
                    if need > top:
                        xo[top * w:need * w].copy_(x_host[top * w:need * w], non_blocking=True)
                        top = need
                    landed = torch.cuda.Event()
                    landed.record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(landed)
                    # band-local source rows: owned row r = local row r + 1, so destination owned row a
                    # reads local rows a .. a + 2: row_off = a + 1
                    ops.mesh_stencil(xb[0, :band.n_local].unsqueeze(0), band.dis, rows + 2, b - a, w, a + 1,
                                     bias=bias, relu=relu, out=y[a * w:b * w].unsqueeze(0))
                    ran = torch.cuda.Event()
                    ran.record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ran)
                    out_host[a * w:b * w].copy_(y[a * w:b * w], non_blocking=True)
            with torch.cuda.stream(self.s_in):
                if top < rows:
                    xo[top * w:].copy_(x_host[top * w:], non_blocking=True)
                all_in = torch.cuda.Event()
                all_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                # first / last owned row: the peer launch (neighbours' rows over NVLink inside the kernel)
                self.s_run.wait_event(all_in)
                band.aggregate(xb, bias, relu, out=edge)
                ran = torch.cuda.Event()
                ran.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ran)
                out_host[:w].copy_(edge[0, :w], non_blocking=True)
                if rows > 1:
                    out_host[(rows - 1) * w:].copy_(edge[0, (rows - 1) * w:], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.s_out)
            self.free[p] = done
            cur.wait_event(done)
        return done
