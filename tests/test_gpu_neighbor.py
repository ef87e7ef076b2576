"""GPU parity of gwen_b200.NeighborLoader (csrc/neighbor.cu, through the C ABI) against the restated
torch_sparse / torch_geometric sampler contract (oracle/neighbor_oracle.py): node order, relabelled
edge_index, e_id, gathered x / target_mask -- all bit-exact (integer / byte work)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import gwen_b200 as gw
from oracle import gcn_oracle as orc
from oracle import neighbor_oracle as nbo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _data(ei, n, f, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, f, generator=g)
    mask = torch.rand(n, generator=g) < 0.4
    return SimpleNamespace(x=x.to(dev), edge_index=ei.to(dev), target_mask=mask.to(dev)), x.numpy(), mask.numpy()


def _check(loader, x, ei, mask, hops, bs, input_nodes=None):
    want = list(nbo.neighbor_loader_batches(x, ei.numpy(), mask, [-1] * hops, bs, input_nodes=input_nodes))
    got = list(loader)
    assert len(got) == len(want) == len(loader)
    for g, w in zip(got, want):
        assert g.n_id.cpu().tolist() == w["n_id"].tolist()
        assert g.edge_index.dtype == torch.int64 and g.edge_index.shape == w["edge_index"].shape
        assert np.array_equal(g.edge_index.cpu().numpy(), w["edge_index"])
        assert np.array_equal(g.e_id.cpu().numpy(), w["e_id"])
        assert np.array_equal(g.x.cpu().numpy(), w["x"])                      # bit-exact row gather
        assert g.target_mask.dtype == torch.bool and np.array_equal(g.target_mask.cpu().numpy(), w["target_mask"])
        assert g.batch_size == w["batch_size"] and g.num_nodes == len(w["n_id"])
        assert g.input_id.cpu().tolist() == w["input_id"].tolist()


@pytest.mark.parametrize("n,bs", [(2, 1), (5, 1), (5, 2), (125, 1), (125, 21)])
@pytest.mark.parametrize("complete", ["auto", False])
def test_complete_graph_batches(dev, n, bs, complete):
    """GWEN's graph (utils.py:176) with the reference batch sizes (config.json:2 / models_gnn.py:54): the
    closed-form kernel and the general kernels both reproduce the sampler's output."""
    ei = orc.complete_graph(n)
    data, x, mask = _data(ei, n, 12, dev, seed=n)
    loader = gw.NeighborLoader(data, [-1, -1], batch_size=bs, complete=complete)
    assert loader._complete == (complete == "auto")
    if n == 125 and bs == 1:     # 125 whole-graph batches: check a few, iterate all
        batches = list(loader)
        want = list(nbo.neighbor_loader_batches(x, ei.numpy(), mask, [-1, -1], bs))
        for i in (0, 1, 63, 124):
            assert batches[i].n_id.cpu().tolist() == want[i]["n_id"].tolist()
            assert np.array_equal(batches[i].edge_index.cpu().numpy(), want[i]["edge_index"])
            assert np.array_equal(batches[i].e_id.cpu().numpy(), want[i]["e_id"])
            assert np.array_equal(batches[i].x.cpu().numpy(), want[i]["x"])
        return
    _check(loader, x, ei, mask, 2, bs)


@pytest.mark.parametrize("name,hops,bs", [("grid3x4", 2, 1), ("grid17x23", 2, 21), ("grid17x23", 3, 5),
                                         ("random", 2, 7), ("random", 1, 64), ("random_sparse", 2, 33),
                                         ("path", 2, 1), ("unsorted", 2, 2), ("empty", 2, 4)])
def test_general_graph_batches(dev, name, hops, bs):
    g = torch.Generator().manual_seed(3)
    if name == "grid3x4":
        ei, n = orc.grid(3, 4), 12
    elif name == "grid17x23":
        ei, n = orc.grid(17, 23), 17 * 23
    elif name == "random":
        ei, n = torch.randint(0, 300, (2, 4000), generator=g), 300       # duplicates and self loops included
    elif name == "random_sparse":
        ei, n = torch.randint(0, 1000, (2, 700), generator=g), 1000      # many isolated nodes
    elif name == "path":
        ei, n = torch.tensor([[0, 1, 2], [1, 2, 3]]), 4
    elif name == "unsorted":
        ei, n = torch.tensor([[3, 1, 2, 0, 2], [0, 0, 0, 1, 3]]), 4
    else:
        ei, n = torch.empty((2, 0), dtype=torch.long), 6
    data, x, mask = _data(ei, n, 5, dev, seed=1)
    _check(gw.NeighborLoader(data, [-1] * hops, batch_size=bs), x, ei, mask, hops, bs)


def test_input_nodes_shuffle_and_errors(dev):
    ei, n = orc.grid(9, 8), 72
    data, x, mask = _data(ei, n, 3, dev)
    sel = torch.tensor([70, 3, 41, 8, 9])
    _check(gw.NeighborLoader(data, [-1, -1], batch_size=2, input_nodes=sel), x, ei, mask, 2, 2,
           input_nodes=sel.numpy())
    # shuffle=True consumes the global RNG like torch's RandomSampler and visits every node once
    torch.manual_seed(5)
    seen = torch.cat([b.n_id[:b.batch_size] for b in gw.NeighborLoader(data, [-1, -1], batch_size=16, shuffle=True)])
    assert sorted(seen.cpu().tolist()) == list(range(n))
    torch.manual_seed(5)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    assert seen.cpu().tolist() == torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist()
    with pytest.raises(NotImplementedError):
        gw.NeighborLoader(data, [10, 10], batch_size=2)
    with pytest.raises(IndexError):
        gw.NeighborLoader(data, [-1, -1]).extract(torch.tensor([3, 3]))
    with pytest.raises(RuntimeError):
        gw.NeighborLoader(SimpleNamespace(x=data.x.cpu(), edge_index=data.edge_index), [-1, -1])


def test_reference_loop_shape_dataset_loader_model(dev):
    """The reference inner loop (models_gnn.py:350-370) on the reference's own sample shape: GraphDataset ->
    NeighborLoader -> GNNModel -> masked L1 on every batch, against the oracle model on the oracle batches."""
    torch.manual_seed(23)
    np.random.seed(23)
    t, members, hgt, cells = 2, 7, 3, 4
    arr = np.random.rand(t, members, hgt, cells).astype(np.float32)
    ds = gw.GraphDataset(arr, split=5, device=dev)
    c = hgt * cells
    cfg = gw.GNNConfig(nodes_in=members, nodes_out=members, channels_in=c, channels_out=c, hidden_feats=64)
    model = gw.GNNModel(cfg)
    ref = orc.GNNModelOracle(c, c, 64)
    ref.load_state_dict(model.state_dict())
    model = model.to(dev)
    ei_cpu = ds.edge_index.cpu()
    for idx in range(len(ds)):
        data = ds[idx]
        want = list(nbo.neighbor_loader_batches(data.x.cpu().numpy(), ei_cpu.numpy(), data.target_mask.cpu().numpy(),
                                                [-1, -1], 3))
        for flow, w in zip(gw.NeighborLoader(data, num_neighbors=[-1] * 2, batch_size=3, shuffle=False), want):
            with torch.no_grad():
                out = model(flow.x, flow.edge_index)
                loss = gw.masked_l1_loss(out, flow.x, flow.target_mask)
                xr = torch.from_numpy(w["x"])
                outr = ref(xr, torch.from_numpy(w["edge_index"]))
                lossr = orc.loss_func(outr, xr, torch.from_numpy(w["target_mask"]))
            assert ((out.cpu() - outr).abs().max() / outr.abs().max()) <= 1e-5
            assert abs(loss.item() - lossr.item()) <= 1e-5 * abs(lossr.item())
