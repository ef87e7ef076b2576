"""CPU tests that pin the oracle: analytic known-answer tests, the
independent dense fp64 restatement, and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import gcn_oracle as orc
from tests.golden import weights as wts


def nmax(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def test_complete_graph_matches_erdos_renyi_and_consumes_rng():
    for n in (2, 5, 17):
        torch.manual_seed(1)
        ei = orc.erdos_renyi_graph(n, 1)
        after = torch.rand(1)
        torch.manual_seed(1)
        torch.rand(n * (n - 1) // 2)
        assert torch.equal(after, torch.rand(1))          # N(N-1)/2 draws, like PyG
        assert torch.equal(ei, orc.complete_graph(n))
        assert ei.shape == (2, n * (n - 1))
    assert orc.complete_graph(5)[:, :5].tolist() == [[0, 0, 0, 0, 1], [1, 2, 3, 4, 0]]


def test_grid_3x4_known_answer(golden_dir):
    ei = orc.grid(3, 4)
    assert ei.shape == (2, 70)
    assert ei[1][ei[0] == 0].tolist() == [0, 1, 4, 5]
    assert ei[1][ei[0] == 1].tolist() == [0, 1, 2, 4, 5, 6]
    assert ei[1][ei[0] == 5].tolist() == [0, 1, 2, 4, 5, 6, 8, 9, 10]
    key = ei[0] * 12 + ei[1]
    assert torch.all(key[1:] > key[:-1])                   # sorted by (row, col), no duplicates
    assert torch.equal(ei, torch.from_numpy(np.load(os.path.join(golden_dir, "grid_3x4_edges.npy"))))
    deg = torch.bincount(orc.add_remaining_self_loops(ei, 12)[1], minlength=12)
    assert sorted(set(deg.tolist())) == [4, 6, 9]
    # symmetric edge set
    back = torch.stack([ei[1], ei[0]])
    kb = torch.sort(back[0] * 12 + back[1]).values
    assert torch.equal(kb, key)


@pytest.mark.parametrize("h,w", [(1, 1), (1, 5), (5, 1), (2, 2), (4, 7)])
def test_grid_closed_form(h, w):
    ei = orc.grid(h, w)
    want = []
    for r in range(h):
        for c in range(w):
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    rr, cc = r + dr, c + dc
                    if 0 <= rr < h and 0 <= cc < w:
                        want.append((r * w + c, rr * w + cc))
    assert ei.t().tolist() == [list(p) for p in want]


def test_self_loop_normalisation_and_duplicates():
    # C.4: existing loops are replaced by one loop of weight 1; duplicates count twice
    ei = torch.tensor([[0, 1, 1, 2, 2, 0], [1, 1, 2, 2, 0, 1]])   # (1,1),(2,2) loops; (0,1) twice
    e2 = orc.add_remaining_self_loops(ei, 3)
    assert e2.tolist() == [[0, 1, 2, 0, 0, 1, 2], [1, 2, 0, 1, 0, 1, 2]]
    _, ew, dis = orc.gcn_norm(ei, 3)
    deg = torch.tensor([2.0, 3.0, 2.0])
    assert torch.allclose(dis, deg.pow(-0.5))
    x = wts.features((3, 4), 1)
    w = wts.glorot(5, 4, 2)
    b = wts.small_bias(5, 3)
    assert nmax(orc.gcn_conv_forward(x, ei, w, b), orc.dense_gcn_forward(x, ei, w, b)) < 1e-6


def test_k2_and_kn_known_answer():
    # C.1 / C.2: complete graph -> A_hat = 1/N, every output row = mean(x) W^T + b
    for n in (2, 125):
        x = wts.features((n, 10), n)
        w, b = wts.glorot(7, 10, 3), wts.small_bias(7, 4)
        out = orc.gcn_conv_forward(x, orc.complete_graph(n), w, b)
        want = (x.double().mean(0, keepdim=True) @ w.double().t() + b.double()).expand(n, -1)
        assert nmax(out, want) < 1e-6
        assert (out - out[0:1]).abs().max() < 1e-6


def test_isolated_nodes_and_empty_edge_list():
    x = wts.features((4, 6), 9)
    w, b = wts.glorot(3, 6, 1), wts.small_bias(3, 2)
    ei = torch.empty((2, 0), dtype=torch.long)
    out = orc.gcn_conv_forward(x, ei, w, b)
    assert nmax(out, x.double() @ w.double().t() + b.double()) < 1e-6


def test_dis_modes_within_one_ulp():
    deg = torch.arange(1, 5001, dtype=torch.float32)
    a = deg.pow(-0.5)
    e = orc.exact_dis(deg)
    ulp = torch.from_numpy(np.spacing(e.numpy()))
    assert torch.all((a - e).abs() <= ulp)
    assert orc.exact_dis(torch.tensor([0.0, 4.0])).tolist() == [0.0, 0.5]


def test_dense_fp64_oracle_random_graphs():
    g = torch.Generator().manual_seed(0)
    for n, e in ((1, 0), (9, 30), (64, 500), (300, 2000)):
        ei = torch.randint(0, n, (2, e), generator=g)
        x = wts.features((n, 16), n)
        w, b = wts.glorot(24, 16, 5), wts.small_bias(24, 6)
        assert nmax(orc.gcn_conv_forward(x, ei, w, b), orc.dense_gcn_forward(x, ei, w, b)) < 1e-5


def test_batched_equals_loop():
    ei = orc.grid(4, 5)
    x = wts.features((3, 20, 8), 11)
    w, b = wts.glorot(6, 8, 1), wts.small_bias(6, 2)
    out = orc.gcn_conv_forward(x, ei, w, b)
    for i in range(3):
        assert torch.equal(out[i], orc.gcn_conv_forward(x[i], ei, w, b))


def test_dst_sorted_csr(golden_dir):
    ei = orc.grid(5, 7)
    rowptr, src, perm, dis = orc.dst_sorted_csr(ei, 35)
    gold = np.load(os.path.join(golden_dir, "grid_5x7_csr.npz"))
    for k, v in (("rowptr", rowptr), ("src", src), ("perm", perm), ("dis", dis)):
        assert np.array_equal(gold[k], v)
    e2 = orc.add_remaining_self_loops(ei, 35).numpy()
    for i in range(35):
        seg = slice(rowptr[i], rowptr[i + 1])
        assert np.all(e2[1][perm[seg]] == i)
        assert np.all(np.diff(perm[seg]) > 0)              # stable: list order inside a segment
        assert src[seg][-1] == i                           # the self loop is last (B.3)
        assert np.all(np.diff(src[seg][:-1]) > 0)
    # CSR-order accumulation reproduces scatter_add_ bitwise
    x = wts.features((35, 8), 3)
    _, ew, _ = orc.gcn_norm(ei, 35, dis_mode="exact")
    ref = orc.propagate(x, torch.from_numpy(e2), ew, 35)
    out = torch.zeros_like(x)
    wcsr = (dis[src] * np.float32(1.0)) * dis[np.repeat(np.arange(35), np.diff(rowptr))]
    for i in range(35):
        acc = np.zeros(8, dtype=np.float32)
        for s in range(rowptr[i], rowptr[i + 1]):
            acc = acc + wcsr[s] * x[src[s]].numpy()
        out[i] = torch.from_numpy(acc)
    assert torch.equal(out, ref)


def test_golden_layer_and_cfg1(golden_dir):
    g = np.load(os.path.join(golden_dir, "layer_grid_6x5.npz"))
    x, w, b = wts.features((30, 12), 5), wts.glorot(20, 12, 6), wts.small_bias(20, 7)
    assert np.array_equal(g["x"], x.numpy()) and np.array_equal(g["w"], w.numpy())
    out = orc.gcn_conv_forward(x, orc.grid(6, 5), w, b)
    assert nmax(out, torch.from_numpy(g["out"])) < 1e-6
    assert nmax(out, torch.from_numpy(g["dense"])) < 1e-6
    # config 1: the reference's own sample data through the full six-layer model on K_2
    xs = torch.from_numpy(np.load(os.path.join(golden_dir, "cfg1_x.npy")))
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, "cfg1_out.npy")))
    model = orc.GNNModelOracle(100, 100, 1024)
    wts.fill_model_(model, 23)
    ei = orc.complete_graph(2)
    with torch.no_grad():
        for t in range(2):
            assert nmax(model(xs[t], ei), gold[t]) < 1e-6


def test_model_wiring_and_state_dict_keys():
    model = orc.GNNModelOracle(8, 8, 64)
    keys = set(model.state_dict())
    want = {"conv_layers.down_conv_layers.conv%d.%s" % (i, s) for i in range(1, 6) for s in ("bias", "lin.weight")}
    want |= {"conv_layers.up_conv_layers.upconv%d.%s" % (i, s) for i in range(1, 6) for s in ("bias", "lin.weight")}
    assert keys == want
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.endswith("weight")}
    assert shapes["conv_layers.down_conv_layers.conv1.lin.weight"] == (64, 8)
    assert shapes["conv_layers.down_conv_layers.conv3.lin.weight"] == (16, 32)
    assert shapes["conv_layers.up_conv_layers.upconv3.lin.weight"] == (32, 16)
    assert shapes["conv_layers.up_conv_layers.upconv5.lin.weight"] == (8, 64)
    # ReLU after the first five live layers only: the output can be negative
    wts.fill_model_(model, 1)
    x = wts.features((6, 8), 2)
    y = model(x, orc.complete_graph(6))
    assert (y < 0).any()
    # manual chain == model
    d, u = model.conv_layers.down_conv_layers, model.conv_layers.up_conv_layers
    ei = orc.complete_graph(6)
    h = x
    for m in (d.conv1, d.conv2, d.conv3, u.upconv3, u.upconv4):
        h = torch.relu(m(h, ei))
    assert torch.equal(u.upconv5(h, ei), y)


def test_gradcheck_fp64():
    ei = orc.grid(3, 3)
    x = torch.randn(9, 4, dtype=torch.double, requires_grad=True)
    w = torch.randn(5, 4, dtype=torch.double, requires_grad=True)
    b = torch.randn(5, dtype=torch.double, requires_grad=True)
    assert torch.autograd.gradcheck(lambda x, w, b: orc.gcn_conv_forward(x, ei, w, b), (x, w, b))
