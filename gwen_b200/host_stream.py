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

__all__ = ["HostPropagator"]


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
