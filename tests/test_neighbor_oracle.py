"""Known-answer tests that pin oracle/neighbor_oracle.py (the restatement of torch_sparse 0.6.17
``neighbor_sample`` + torch_geometric 2.3.1 ``to_csc`` / ``filter_data`` behind the reference's
``NeighborLoader(data, [-1, -1], batch_size)``, src/gwen/models_gnn.py:351-356).  Hand-derived answers:
the reference tree holds no fixture for this step ("parity unpinned")."""
import numpy as np
import pytest

from oracle import gcn_oracle as orc
from oracle import neighbor_oracle as nbo


def _k(n):
    return orc.complete_graph(n).numpy()


def test_k2_batch1():
    ei = _k(2)                                      # [[0, 1], [1, 0]]
    x = np.arange(2 * 3, dtype=np.float32).reshape(2, 3)
    mask = np.array([False, True])
    b0, b1 = list(nbo.neighbor_loader_batches(x, ei, mask, [-1, -1], 1))
    assert b0["n_id"].tolist() == [0, 1] and b1["n_id"].tolist() == [1, 0]
    # hop 1: seed's in-neighbour; hop 2: that node's in-neighbour (the seed)
    assert b0["edge_index"].tolist() == [[1, 0], [0, 1]]
    assert b0["e_id"].tolist() == [1, 0]            # (1 -> 0) is edge 1, (0 -> 1) is edge 0
    assert b1["e_id"].tolist() == [0, 1]
    assert np.array_equal(b1["x"], x[[1, 0]]) and b1["target_mask"].tolist() == [True, False]
    assert b0["batch_size"] == 1 and b0["input_id"].tolist() == [0]


@pytest.mark.parametrize("n,bs", [(5, 1), (5, 2), (125, 1), (125, 21)])
def test_complete_graph_closed_form(n, bs):
    """On K_N every batch is the whole graph renumbered [seeds | the others ascending]; edges are visited
    frontier node by frontier node (seeds, then everybody else), in-neighbours ascending."""
    ei = _k(n)
    x = np.arange(n, dtype=np.float32).reshape(n, 1)
    mask = np.arange(n) % 3 == 0
    batches = list(nbo.neighbor_loader_batches(x, ei, mask, [-1, -1], bs))
    assert len(batches) == -(-n // bs)
    for bi, b in enumerate(batches):
        s0, k = bi * bs, min(bs, n - bi * bs)
        want = list(range(s0, s0 + k)) + [v for v in range(n) if not s0 <= v < s0 + k]
        assert b["n_id"].tolist() == want
        loc = {v: i for i, v in enumerate(want)}
        rows, cols, eids = [], [], []
        for i, w in enumerate(want):
            for v in range(n):
                if v != w:
                    rows.append(loc[v])
                    cols.append(i)
                    eids.append(v * (n - 1) + (w - (w > v)))
        assert b["edge_index"].tolist() == [rows, cols]
        assert b["e_id"].tolist() == eids
        assert b["edge_index"].shape[1] == n * (n - 1)
        # the relabelled graph is the same graph: e_id maps back to the original edges
        assert np.array_equal(b["n_id"][b["edge_index"][0]], ei[0][b["e_id"]])
        assert np.array_equal(b["n_id"][b["edge_index"][1]], ei[1][b["e_id"]])


def test_grid_3x4_seed0_hand_derived():
    ei = orc.grid(3, 4).numpy()                     # 8-neighbour mesh with self loops, node id = r * 4 + c
    x = np.arange(12, dtype=np.float32).reshape(12, 1)
    mask = np.zeros(12, dtype=bool)
    b = next(iter(nbo.neighbor_loader_batches(x, ei, mask, [-1, -1], 1)))
    # hop 1: in-neighbours of 0 = {0, 1, 4, 5}; hop 2: of 1 = {0,1,2,4,5,6}, of 4 = {0,1,4,5,8,9},
    # of 5 = {0,1,2,4,5,6,8,9,10}
    assert b["n_id"].tolist() == [0, 1, 4, 5, 2, 6, 8, 9, 10]
    rows = [0, 1, 2, 3] + [0, 1, 4, 2, 3, 5] + [0, 1, 2, 3, 6, 7] + [0, 1, 4, 2, 3, 5, 6, 7, 8]
    cols = [0] * 4 + [1] * 6 + [2] * 6 + [3] * 9
    assert b["edge_index"].tolist() == [rows, cols]


def test_directed_path_two_hops_only():
    ei = np.array([[0, 1, 2], [1, 2, 3]])           # 0 -> 1 -> 2 -> 3
    x = np.zeros((4, 1), dtype=np.float32)
    b = list(nbo.neighbor_loader_batches(x, ei, np.zeros(4, bool), [-1, -1], 1, input_nodes=[3]))[0]
    assert b["n_id"].tolist() == [3, 2, 1]          # node 0 is three hops away
    assert b["edge_index"].tolist() == [[1, 2], [0, 1]]
    assert b["e_id"].tolist() == [2, 1]
    # an isolated seed
    b = list(nbo.neighbor_loader_batches(x, ei, np.zeros(4, bool), [-1, -1], 1, input_nodes=[0]))[0]
    assert b["n_id"].tolist() == [0] and b["edge_index"].shape == (2, 0)


def test_unsorted_edge_list_keeps_edge_index_order_within_a_destination():
    """to_csc is a STABLE sort by destination: in-neighbours are visited in edge_index order."""
    ei = np.array([[3, 1, 2, 0], [0, 0, 0, 1]])
    x = np.zeros((4, 1), dtype=np.float32)
    b = list(nbo.neighbor_loader_batches(x, ei, np.zeros(4, bool), [-1], 1, input_nodes=[0]))[0]
    assert b["n_id"].tolist() == [0, 3, 1, 2]
    assert b["e_id"].tolist() == [0, 1, 2]
