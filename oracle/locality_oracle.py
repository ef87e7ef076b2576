"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the locality-tile preprocessor ``gwen_locality_tiles``
(``gwen_b200/csrc/locality.cu``).  Only ``tests/`` may import this module.

There is no reference counterpart: torch_geometric's ``scatter_add`` path (reference call sites
``src/gwen/models_gnn.py:147-149,204-206``) has no locality pass, and the aggregation result does not depend on the
tiles at all (the staged kernel sums every destination's messages in CSR order whatever the plan).  This file pins
the INTEGER outputs of the CUDA pass (order, tile boundaries, cells, depths: bit-exact) against an independent
statement of the same algorithm, so that a change of the kernels cannot silently change the plans:

1. seeds: Luby rounds for a maximal independent set of the ``radius``-th power of the graph; priority of node v is
   ``mix32(v + 1)`` (murmur3 finaliser, a bijection); an undecided node becomes a seed when it holds the largest
   priority among the undecided nodes that reach it within ``radius`` hops; nodes a new seed reaches within
   ``radius`` hops are covered.  After ``rounds`` rounds every still undecided node becomes a seed.
2. cells: multi-source BFS over ``radius`` levels, an unassigned node joins the smallest cell index among the
   in-neighbours assigned one level earlier (cell index = rank of the seed's node id); nodes never reached share one
   extra cell at depth 0.
3. order: nodes sorted by (cell, min(depth, 255), id); tiles: cells above ``merge_rows`` rows stand alone, cut into
   ``ceil(size / cap_rows)`` equal chunks; smaller cells are packed in index order while a tile stays within
   ``merge_rows``.
4. ``deal > 0``: tiles renumbered by descending row count and dealt in snake order over ``deal`` CTAs (the staged kernel
   gives tile t to CTA ``t mod grid``).
"""
from __future__ import annotations

import numpy as np


def mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def locality_tiles(rowptr: np.ndarray, src: np.ndarray, n: int, radius: int, rounds: int, merge_rows: int,
                   cap_rows: int, deal: int = 0):
    """-> (order int32[n], tile_ptr int32[tiles + 1], cell int32[n], depth int32[n], status [cells, tiles,
    largest cell, unreached])."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    src = np.asarray(src, dtype=np.int64)
    dst = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr[:n + 1]))
    keep = src < n
    s_e, d_e = src[keep], dst[keep]
    prio = mix32(np.arange(n, dtype=np.uint64) + 1)

    def sweep(a):
        out = a.copy()
        np.maximum.at(out, d_e, a[s_e])
        return out

    UND, SEED, COV = 0, 1, 2
    state = np.zeros(n, dtype=np.uint8)
    for _ in range(rounds):
        m = np.where(state == UND, prio, 0).astype(np.uint32)
        for _h in range(radius):
            m = sweep(m)
        new = (state == UND) & (m == prio)
        state[new] = SEED
        f = new.astype(np.uint32)
        for _h in range(radius):
            f = sweep(f)
        state[(state == UND) & (f > 0)] = COV
    is_seed = state != COV
    seed_idx = np.cumsum(is_seed) - is_seed          # exclusive scan
    n_seeds = int(is_seed.sum())
    cell = np.where(is_seed, seed_idx, -1).astype(np.int64)
    depth = np.where(is_seed, 0, -1).astype(np.int64)
    big = np.iinfo(np.int64).max
    for d in range(1, radius + 1):
        sel = (depth[s_e] == d - 1) & (depth[d_e] < 0)
        best = np.full(n, big, dtype=np.int64)
        np.minimum.at(best, d_e[sel], cell[s_e[sel]])
        hit = best != big
        cell[hit] = best[hit]
        depth[hit] = d
    unreached = int((cell < 0).sum())
    depth[cell < 0] = 0
    cell[cell < 0] = n_seeds
    cells = n_seeds + (1 if unreached else 0)
    order = np.lexsort((np.arange(n), np.minimum(depth, 255), cell)).astype(np.int32)
    size = np.bincount(cell, minlength=cells)
    tile_ptr, pos, cur = [0], 0, 0
    for c in range(cells):
        s = int(size[c])
        if s == 0:
            continue
        if s > merge_rows:
            if cur > 0:
                tile_ptr.append(pos)
                cur = 0
            chunks = -(-s // cap_rows)
            for j in range(1, chunks + 1):
                tile_ptr.append(pos + s * j // chunks)
            pos += s
        else:
            if cur + s > merge_rows:
                tile_ptr.append(pos)
                cur = 0
            cur += s
            pos += s
    if cur > 0:
        tile_ptr.append(pos)
    status = [cells, len(tile_ptr) - 1, int(size.max()) if cells else 0, unreached]
    tile_ptr = np.asarray(tile_ptr, dtype=np.int64)
    if deal > 0:
        # 4. tiles renumbered by size: descending row count (ties: old index), dealt in snake order over `deal` CTAs
        nt = len(tile_ptr) - 1
        sizes = np.diff(tile_ptr)
        by_size = np.lexsort((np.arange(nt), -sizes))
        new_of_old = np.empty(nt, dtype=np.int64)
        for k, old in enumerate(by_size):
            r, p = divmod(k, deal)
            cnt = min(deal, nt - r * deal)
            new_of_old[old] = r * deal + (cnt - 1 - p if r & 1 else p)
        old_of_new = np.argsort(new_of_old)
        new_ptr = np.concatenate([[0], np.cumsum(sizes[old_of_new])])
        order = np.concatenate([order[tile_ptr[o]:tile_ptr[o + 1]] for o in old_of_new]).astype(np.int32)
        tile_ptr = new_ptr
    return order, tile_ptr.astype(np.int32), cell.astype(np.int32), depth.astype(np.int32), status
