"""CPU oracle for the NeighborLoader step of the GWEN loops.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this file; nothing under ``gwen_b200/`` does.

What it restates
----------------
The reference builds its mini-batches with

    NeighborLoader(data, num_neighbors=[-1] * 2, batch_size=configs.batch_size, shuffle=False)

(``src/gwen/models_gnn.py:351-356`` training, ``:434-439`` evaluation) and consumes ``.x``,
``.edge_index`` and ``.target_mask`` of every batch (``:359-361``, ``:441-443``).  The code lives in
third-party packages that are NOT in the reference tree and not installable here:

* ``torch-geometric==2.3.1`` (``requirements/environment.yml:552``):
  ``loader/neighbor_loader.py`` (NeighborLoader -> NodeLoader), ``sampler/neighbor_sampler.py``
  (``NeighborSampler._sample``: homogeneous graph -> ``torch.ops.torch_sparse.neighbor_sample(colptr, row,
  seed, num_neighbors, replace=False, directed=True)`` -- pyg-lib is not in the reference environment),
  ``sampler/utils.py::to_csc`` (CSC by a stable sort of the destination column) and
  ``loader/utils.py::filter_data`` (node attributes indexed by the node list, ``edge_index`` rebuilt as
  ``stack([row, col])``).
* ``torch-sparse==0.6.17`` (``requirements/environment.yml:554``): ``csrc/cpu/neighbor_sample_cpu.cpp``,
  ``sample<replace=false, directed=true>`` -- the sequential loop restated in ``neighbor_sample`` below.

PARITY PINNING: the reference's tests never construct a NeighborLoader and hold no batch fixture, and
neither package can be imported here: this oracle is **"parity unpinned"**.  It is a line-by-line
restatement of the published algorithm (hash map + BFS loop) pinned by hand-derived known-answer tests
(``tests/test_neighbor_oracle.py``: K_2, K_5, K_125 with batch sizes 1 and 21 as in ``config.json:2`` /
``models_gnn.py:54``, a 3 x 4 grid, a directed path).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def to_csc(edge_index: np.ndarray, num_nodes: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """PyG ``to_csc`` for a ``Data`` with ``edge_index``: stable sort by destination (``col``).
    Returns (colptr int64[N+1], row int64[E] = sources in CSC order, perm int64[E])."""
    row, col = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    perm = np.argsort(col, kind="stable")
    colptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(col, minlength=num_nodes), out=colptr[1:])
    return colptr, row[perm].astype(np.int64), perm.astype(np.int64)


def neighbor_sample(colptr: np.ndarray, row: np.ndarray, input_node: Sequence[int],
                    num_neighbors: Sequence[int]) -> Tuple[List[int], List[int], List[int], List[int]]:
    """torch-sparse 0.6.17 ``sample<replace=false, directed=true>`` for ``num_neighbors`` entries that take
    every neighbour (``-1``, or a fan-out >= the in-degree).  Returns (samples, rows, cols, edges)."""
    samples: List[int] = []
    to_local = {}
    for i, v in enumerate(input_node):
        samples.append(int(v))
        to_local.setdefault(int(v), i)
    rows: List[int] = []
    cols: List[int] = []
    edges: List[int] = []
    begin, end = 0, len(samples)
    for num in num_neighbors:
        for i in range(begin, end):
            w = samples[i]
            c0, c1 = int(colptr[w]), int(colptr[w + 1])
            if c1 == c0:
                continue
            if not (num < 0 or num >= c1 - c0):
                raise NotImplementedError("random fan-out sampling is not part of the GWEN path")
            for off in range(c0, c1):
                v = int(row[off])
                if v not in to_local:
                    to_local[v] = len(samples)
                    samples.append(v)
                cols.append(i)
                rows.append(to_local[v])
                edges.append(off)
        begin, end = end, len(samples)
    return samples, rows, cols, edges


def neighbor_loader_batches(x: np.ndarray, edge_index: np.ndarray, target_mask: np.ndarray,
                            num_neighbors: Sequence[int], batch_size: int, input_nodes=None):
    """The batches ``NeighborLoader(data, num_neighbors, batch_size, shuffle=False)`` yields, as dicts with
    ``x``, ``edge_index`` (int64 [2, E_b]), ``target_mask``, ``n_id``, ``e_id``, ``input_id``, ``batch_size``."""
    n = x.shape[0]
    colptr, row, perm = to_csc(edge_index, n)
    ids = np.arange(n) if input_nodes is None else np.asarray(input_nodes)
    for s in range(0, len(ids), batch_size):
        seed = ids[s:s + batch_size]
        node, r, c, e = neighbor_sample(colptr, row, seed.tolist(), num_neighbors)
        node = np.asarray(node, dtype=np.int64)
        yield {
            "x": x[node], "target_mask": target_mask[node],
            "edge_index": np.stack([np.asarray(r, dtype=np.int64), np.asarray(c, dtype=np.int64)]).reshape(2, -1),
            "n_id": node, "e_id": perm[np.asarray(e, dtype=np.int64)] if len(e) else np.zeros(0, dtype=np.int64),
            "input_id": np.arange(s, s + len(seed)), "batch_size": len(seed),
        }
