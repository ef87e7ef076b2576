"""``NeighborLoader``: the mini-batch step in front of the layer stack, on the GPU.

Reference call sites: ``NeighborLoader(data, num_neighbors=[-1] * 2, batch_size=configs.batch_size,
shuffle=False)`` at ``src/gwen/models_gnn.py:351-356`` (training) and ``:434-439`` (evaluation); every
batch's ``.x``, ``.edge_index`` and ``.target_mask`` feed one model call (``:359-365``).  The reference
runs torch_geometric 2.3.1's loader -> torch_sparse 0.6.17 ``neighbor_sample`` on the CPU for every
batch of every time step; on GWEN's complete member graph each batch re-emits the WHOLE graph
renumbered, so with ``batch_size = 1`` (``config.json:2``) that is ``N`` whole-graph extractions per
time step before any layer runs.

Here the extraction runs in ``libgwen_b200.so`` (``gwen_neighbor_sample_full`` /
``gwen_neighbor_complete``, csrc/neighbor.cu) on data that already lives on the device, in the output
order of the reference sampler (contract restated in ``oracle/neighbor_oracle.py``; "parity unpinned":
the third-party source is not in the reference tree):

* ``n_id``: the batch's seeds first, then newly reached nodes in discovery order, hop by hop;
* ``edge_index``: relabelled, one edge per (frontier node, in-neighbour) in visiting order;
* ``x = data.x[n_id]``, ``target_mask = data.target_mask[n_id]`` (PyG ``filter_data``).

Only full-neighbour fan-outs (``-1``) are served -- the only form the reference uses; random fan-outs
would have to reproduce torch_sparse's RNG stream.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib
from .graph import _build, _ptr, _require_cuda, _stream

__all__ = ["NeighborLoader"]


class NeighborLoader:
    """``NeighborLoader(data, num_neighbors, batch_size=1, shuffle=False, input_nodes=None)``.

    ``data`` needs ``x [N, ...]``, ``edge_index int64 [2, E]`` (CUDA tensors) and optionally
    ``target_mask [N]`` -- what ``GraphDataset.get`` returns.  Iterating yields objects with ``x``,
    ``edge_index``, ``target_mask``, ``n_id``, ``e_id``, ``input_id``, ``batch_size``, ``num_nodes``.
    ``complete="auto"`` uses the closed-form one-launch kernel when ``edge_index`` is the complete
    directed graph sorted by (row, col) (checked once), ``False`` forces the general kernels.
    """

    def __init__(self, data, num_neighbors: Sequence[int], batch_size: int = 1, shuffle: bool = False,
                 input_nodes: Optional[torch.Tensor] = None, drop_last: bool = False, complete="auto"):
        if any(int(k) >= 0 for k in num_neighbors):
            raise NotImplementedError("gwen_b200.NeighborLoader serves full-neighbour fan-outs (-1) only "
                                      "(the reference uses num_neighbors=[-1] * 2)")
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        self.data, self.hops, self.batch_size = data, len(num_neighbors), int(batch_size)
        self.shuffle, self.drop_last = shuffle, drop_last
        x, ei = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(ei, "data.edge_index")
        self.n = int(x.shape[0])
        self.e = int(ei.shape[1])
        self.dev = x.device
        self.input_nodes = None if input_nodes is None else torch.as_tensor(input_nodes, dtype=torch.int64).to(self.dev)
        # CSC of the RAW edge list (no self-loop normalisation): K0 with flags = 0 is destination-sorted
        # and stable in edge_index order = PyG's to_csc
        self._csc = _build(ei, self.n, 0)
        self._complete = False
        if complete and self.n >= 2 and self.e == self.n * (self.n - 1) and self.hops >= 2 and self.input_nodes is None:
            from .graph import complete_graph
            self._complete = bool(torch.equal(complete_graph(self.n, self.dev), ei))   # one-time check
        with torch.cuda.device(self.dev):
            need = C.c_size_t()
            check(lib().gwen_neighbor_workspace_bytes(self.n, self.e, C.byref(need)), "neighbor ws")
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.dev)
            self._counts = torch.zeros(4, dtype=torch.int32, device=self.dev)

    def __len__(self) -> int:
        m = self.n if self.input_nodes is None else int(self.input_nodes.numel())
        return m // self.batch_size if self.drop_last else -(-m // self.batch_size)

    # -- one batch ------------------------------------------------------------------------------
    def _gather(self, src: torch.Tensor, n_id: torch.Tensor) -> torch.Tensor:
        src = src.contiguous()
        rows = int(n_id.numel())
        out = torch.empty((rows,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        row_bytes = src.element_size() * (src.numel() // max(1, src.shape[0]))
        check(lib().gwen_gather_rows_bytes(_ptr(src), _ptr(n_id), _ptr(out), rows, row_bytes, src.shape[0],
                                           _stream()), "gwen_gather_rows_bytes")
        return out

    def extract(self, seeds: torch.Tensor, input_id: Optional[torch.Tensor] = None,
                seed_start: Optional[int] = None):
        """The batch for the given seed nodes (int64 CUDA tensor, distinct).  ``seed_start``: the caller
        knows the seeds are ``seed_start .. seed_start + len - 1`` (the un-shuffled loader), which lets the
        complete-graph path run without looking at the device."""
        seeds = seeds.to(device=self.dev, dtype=torch.int64).contiguous()
        bs = int(seeds.numel())
        n, e = self.n, self.e
        with torch.cuda.device(self.dev):
            node = torch.empty(n, dtype=torch.int64, device=self.dev)
            lo = -1
            if self._complete and bs:
                if seed_start is not None:
                    lo = int(seed_start)
                else:
                    lo = int(seeds[0].item())
                    if not (lo + bs <= n and (bs == 1 or torch.equal(seeds, torch.arange(lo, lo + bs, device=self.dev)))):
                        lo = -1
            if lo >= 0 and lo + bs <= n:
                ei = torch.empty((2, e), dtype=torch.int64, device=self.dev)
                eid = torch.empty(e, dtype=torch.int64, device=self.dev)
                check(lib().gwen_neighbor_complete(n, lo, bs, _ptr(node), _ptr(ei), _ptr(eid), _stream()),
                      "gwen_neighbor_complete")
                n_id, edge_index, e_id = node, ei, eid
            else:
                row = torch.empty(max(e, 1), dtype=torch.int64, device=self.dev)
                col = torch.empty(max(e, 1), dtype=torch.int64, device=self.dev)
                edge = torch.empty(max(e, 1), dtype=torch.int64, device=self.dev)
                g = self._csc
                check(lib().gwen_neighbor_sample_full(_ptr(g.rowptr), _ptr(g.src), _ptr(g.perm), n, e, _ptr(seeds),
                                                      bs, self.hops, _ptr(node), _ptr(row), _ptr(col), _ptr(edge),
                                                      _ptr(self._counts), _ptr(self._ws), self._ws.numel(),
                                                      _stream()), "gwen_neighbor_sample_full")
                nn, ne, bad = self._counts[:3].tolist()          # the one host round trip of a batch
                if bad:
                    raise IndexError("NeighborLoader: %d seed nodes are out of range or repeated" % bad)
                n_id, e_id = node[:nn], edge[:ne]
                edge_index = torch.stack([row[:ne], col[:ne]])
            data = self.data
            out = SimpleNamespace(x=self._gather(data.x, n_id), edge_index=edge_index, n_id=n_id, e_id=e_id,
                                  batch_size=bs, num_nodes=int(n_id.numel()),
                                  input_id=input_id if input_id is not None else seeds)
            if getattr(data, "target_mask", None) is not None:
                out.target_mask = self._gather(data.target_mask.to(self.dev), n_id)
        return out

    def __iter__(self):
        m = self.n if self.input_nodes is None else int(self.input_nodes.numel())
        if self.shuffle:
            # torch.utils.data.RandomSampler: a seed drawn from the global RNG feeds a private generator
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            order = torch.randperm(m, generator=torch.Generator().manual_seed(seed)).to(self.dev)
        else:
            order = torch.arange(m, device=self.dev)
        for s in range(0, m, self.batch_size):
            idx = order[s:s + self.batch_size]
            if self.drop_last and idx.numel() < self.batch_size:
                break
            seeds = idx if self.input_nodes is None else self.input_nodes[idx]
            known = s if (not self.shuffle and self.input_nodes is None) else None
            yield self.extract(seeds, input_id=idx, seed_start=known)
