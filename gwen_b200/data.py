"""Graph + feature ingestion around the layer stack.

Mirror of ``GraphDataset`` (reference ``src/gwen/utils.py:164-211``) for data that is already in
memory: the reference wraps an ``xarray`` dataset ``theta_v(time, member, height, ncells)`` and, per
time step, stacks ``(height, ncells)`` into the feature axis (``x [members, H * ncells]`` float32),
builds the complete graph over the ensemble members and a boolean ``target_mask`` from a shuffled
member split.  Here the array is a numpy / torch array with the same axis order (``xarray``, ``zarr``
and ``torch_geometric.data.Data`` are not part of this path); the whole time series is tensorised
ONCE into a pinned host buffer (or onto the device), so ``get(idx)`` is a view, not a load + stack +
``torch.tensor`` copy per sample.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from .graph import erdos_renyi_graph

__all__ = ["GraphDataset"]


class GraphDataset:
    """``GraphDataset(data, split)``; ``data`` is ``[time, member, height, ncells]``.

    Same attributes as the reference (``nodes``, ``edge_index``, ``input_indices``,
    ``target_indices``, ``channels``), same consumption of the torch and numpy global RNGs
    (``erdos_renyi_graph`` draws ``N(N-1)/2`` uniforms; the member permutation comes from
    ``np.random.shuffle``), so a seeded script splits the members identically.
    ``get(idx)`` returns an object with ``x [members, height * ncells]`` float32, ``edge_index`` and
    ``target_mask [members]`` like the reference's ``Data``.
    """

    def __init__(self, data, split: int, device="cuda", resident: str = "device"):
        arr = torch.as_tensor(np.asarray(data) if not torch.is_tensor(data) else data)
        if arr.dim() != 4:
            raise ValueError("data must be [time, member, height, ncells]")
        self.split = split
        self.nodes = int(arr.shape[1])
        self.edge_index = erdos_renyi_graph(self.nodes, edge_prob=1, device=device)   # utils.py:176
        member_indices = np.arange(self.nodes)
        np.random.shuffle(member_indices)                                               # utils.py:181
        self.input_indices = member_indices[: self.split]
        self.target_indices = member_indices[self.split:]
        self.channels = int(arr.shape[2] * arr.shape[3])
        # .stack(features=["height", "ncells"]) of every time step at once: [T, members, H * ncells]
        x = arr.to(torch.float32).reshape(arr.shape[0], self.nodes, self.channels).contiguous()
        if resident == "device":
            self._x = x.to(device)
        elif resident == "pinned":
            self._x = x.pin_memory()
        else:
            raise ValueError("resident must be 'device' or 'pinned'")
        mask = torch.zeros(self.nodes, dtype=torch.bool)
        mask[torch.as_tensor(self.target_indices, dtype=torch.long)] = True
        self._mask = mask.to(self._x.device) if resident == "device" else mask.pin_memory()

    def len(self) -> int:
        return int(self._x.shape[0])

    __len__ = len

    def get(self, idx: int):
        return SimpleNamespace(x=self._x[idx], edge_index=self.edge_index, target_mask=self._mask)

    __getitem__ = get
