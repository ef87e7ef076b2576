"""CPU tests of the locality-tile restatement (oracle/locality_oracle.py): the structural guarantees the CUDA pass
(gwen_b200/csrc/locality.cu) is then held to bit for bit in tests/test_gpu_locality.py."""
import numpy as np
import torch

from oracle import gcn_oracle as orc
from oracle import locality_oracle as lo


def permuted_mesh_csr(h, w, seed, permute=True):
    """dst-sorted CSR (self loops included) of the h x w 8-neighbour mesh with node ids permuted."""
    n = h * w
    ei = orc.grid(h, w)
    if permute:
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
        ei = perm[ei]
    src, dst = ei[0].numpy(), ei[1].numpy()
    keep = src != dst
    src = np.concatenate([src[keep], np.arange(n)])
    dst = np.concatenate([dst[keep], np.arange(n)])
    o = np.lexsort((src, dst))
    src, dst = src[o], dst[o]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    return np.cumsum(rowptr), src, n


def hop_distances(rowptr, src, n, start, limit):
    """BFS depth (<= limit) from ``start`` along CSR edges source -> destination; -1 beyond."""
    dst = np.repeat(np.arange(n), np.diff(rowptr))
    depth = np.full(n, -1)
    depth[start] = 0
    for d in range(1, limit + 1):
        reach = np.zeros(n, dtype=bool)
        reach[dst[depth[src] == d - 1]] = True
        new = reach & (depth < 0)
        depth[new] = d
    return depth


def test_mix32_is_injective_and_nonzero():
    p = lo.mix32(np.arange(1, 200001, dtype=np.uint64))
    assert p.min() > 0 and np.unique(p).size == p.size


def test_structure_on_a_permuted_mesh():
    rowptr, src, n = permuted_mesh_csr(40, 56, 3)
    radius = 5
    order, tile_ptr, cell, depth, status = lo.locality_tiles(rowptr, src, n, radius, 12, 64, 96)
    cells, tiles, biggest, unreached = status
    assert unreached == 0
    assert np.array_equal(np.sort(order), np.arange(n))                  # a permutation
    assert tile_ptr[0] == 0 and tile_ptr[-1] == n and np.all(np.diff(tile_ptr) > 0) and tiles == len(tile_ptr) - 1
    assert np.diff(tile_ptr).max() <= 96
    assert depth.min() == 0 and depth.max() <= radius
    seeds = np.nonzero(depth == 0)[0]
    assert seeds.size == cells
    # independent set of the radius-th power: no other seed within `radius` hops; and maximal: everyone is covered
    for s in seeds[:12]:
        d = hop_distances(rowptr, src, n, s, radius)
        near = np.nonzero((d >= 0) & (depth == 0))[0]
        assert near.tolist() == [s]
    # positions of a cell are contiguous in the order, depth ascending inside a cell
    pos_cell = cell[order]
    assert np.all(np.diff(pos_cell) >= 0)
    same = np.diff(pos_cell) == 0
    assert np.all(np.diff(depth[order])[same] >= 0)
    # compact tiles: far fewer distinct sources than messages (a random 96-row tile would need ~9 per row)
    ts = np.diff(tile_ptr)
    tile_of_node = np.empty(n, dtype=np.int64)
    tile_of_node[order] = np.repeat(np.arange(tiles), ts)
    dst = np.repeat(np.arange(n), np.diff(rowptr))
    distinct = np.unique(tile_of_node[dst] * n + src).size
    assert distinct / n < 2.0


def test_asymmetric_edges_and_isolated_nodes():
    # a directed path 0 -> 1 -> ... -> 9 (CSR by destination) plus two isolated nodes without any edge
    n = 12
    src = np.arange(0, 9)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[2:11] = np.arange(1, 10)
    rowptr[11:] = 9
    order, tile_ptr, cell, depth, status = lo.locality_tiles(rowptr, src, n, 2, 8, 4, 8)
    assert np.array_equal(np.sort(order), np.arange(n))
    assert tile_ptr[0] == 0 and tile_ptr[-1] == n and np.all(np.diff(tile_ptr) > 0)
    assert np.diff(tile_ptr).max() <= 8
    assert cell.min() == 0 and cell.max() == status[0] - 1


def test_rounds_exhausted_leaves_valid_tiles():
    rowptr, src, n = permuted_mesh_csr(16, 16, 1)
    order, tile_ptr, cell, depth, status = lo.locality_tiles(rowptr, src, n, 3, 1, 32, 48)    # one Luby round only
    assert np.array_equal(np.sort(order), np.arange(n)) and tile_ptr[-1] == n
    assert status[3] == 0 and depth.max() <= 3


def test_snake_deal_balances_and_keeps_tiles_intact():
    rowptr, src, n = permuted_mesh_csr(60, 80, 2)
    o0, t0, c0, d0, s0 = lo.locality_tiles(rowptr, src, n, 4, 12, 40, 64, 0)
    o1, t1, c1, d1, s1 = lo.locality_tiles(rowptr, src, n, 4, 12, 40, 64, 7)
    assert s0 == s1 and np.array_equal(c0, c1) and np.array_equal(np.sort(o1), np.arange(n))
    # the same tiles (as node sets), renumbered
    sets0 = sorted(tuple(sorted(o0[t0[i]:t0[i + 1]])) for i in range(len(t0) - 1))
    sets1 = sorted(tuple(sorted(o1[t1[i]:t1[i + 1]])) for i in range(len(t1) - 1))
    assert sets0 == sets1
    # CTA c gets tiles c, c + 7, ...: the dealt numbering is better balanced than cell order
    def heaviest(tp):
        sz = np.diff(tp)
        load = np.zeros(7)
        for t, s in enumerate(sz):
            load[t % 7] += s
        return load.max() / load.mean()
    assert heaviest(t1) <= heaviest(t0) and heaviest(t1) < 1.05
