"""GPU tests of the locality tiles (csrc/locality.cu) and the staged kernel's row-gather producer
(csrc/aggregate.cu, GWEN_PLAN_GATHER): the integer outputs of the pass bit-exact against the numpy restatement,
the aggregation over locality tiles BITWISE equal to the row kernel (which is bit-exact vs the CPU scatter_add_
order, tests/test_gpu_parity.py) and within tolerance of the CPU oracle."""
import numpy as np
import pytest
import torch

import gwen_b200 as gw
from gwen_b200 import ops
from oracle import gcn_oracle as orc
from oracle import locality_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def permuted_mesh(h, w, seed, dev):
    n = h * w
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
    return perm[orc.grid(h, w)].contiguous().to(dev), n


def _check_tiles_vs_oracle(g, radius, rounds, merge_rows, cap_rows):
    for deal in (0, 5, 148):
        status = _check_tiles_vs_oracle_deal(g, radius, rounds, merge_rows, cap_rows, deal)
    return status


def _check_tiles_vs_oracle_deal(g, radius, rounds, merge_rows, cap_rows, deal):
    order, tile_ptr, cell, depth, status = g.locality_tiles(radius, rounds, merge_rows, cap_rows, deal)
    ro, rt, rc, rd, rs = lo.locality_tiles(g.rowptr.cpu().numpy(), g.src.cpu().numpy(), g.n_dst, radius, rounds,
                                           merge_rows, cap_rows, deal)
    assert status == rs
    assert np.array_equal(cell.cpu().numpy(), rc)
    assert np.array_equal(depth.cpu().numpy(), rd)
    assert np.array_equal(order.cpu().numpy(), ro)
    assert np.array_equal(tile_ptr.cpu().numpy(), rt)
    return status


@pytest.mark.parametrize("h,w,radius,rounds,merge,cap", [(40, 56, 5, 12, 64, 96), (33, 47, 8, 12, 160, 256),
                                                        (16, 16, 3, 1, 32, 48), (64, 64, 2, 6, 8, 8)])
def test_locality_tiles_bit_exact_vs_oracle(dev, h, w, radius, rounds, merge, cap):
    ei, n = permuted_mesh(h, w, 7, dev)
    g = gw.build_graph(ei, n)
    assert g.grid_shape is None
    _check_tiles_vs_oracle(g, radius, rounds, merge, cap)
    # deterministic: a second run gives the same arrays
    a = g.locality_tiles(radius, rounds, merge, cap)
    b = g.locality_tiles(radius, rounds, merge, cap)
    assert all(torch.equal(p, q) for p, q in zip(a[:4], b[:4])) and a[4] == b[4]


def test_locality_tiles_asymmetric_and_isolated(dev):
    # directed path + isolated nodes (no self loops added: rows without any message), and an Erdos-Renyi graph
    n = 40
    ei = torch.stack([torch.arange(0, 30), torch.arange(1, 31)]).to(dev)
    g = gw.build_graph(ei, n, add_self_loops=False, grid_shape=None)
    st = _check_tiles_vs_oracle(g, 2, 8, 4, 8)
    assert st[3] >= 0
    torch.manual_seed(5)
    er = orc.erdos_renyi_graph(300, 0.02).to(dev)
    g2 = gw.build_graph(er, 300, grid_shape=None)
    _check_tiles_vs_oracle(g2, 2, 10, 32, 48)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f,b,relu", [(256, 1, False), (64, 3, True), (40, 2, False), (520, 1, True)])
def test_aggregate_over_locality_tiles_bitwise_equals_rows(dev, dtype, f, b, relu):
    if dtype == torch.bfloat16 and f % 8:
        f = (f // 8) * 8
    ei, n = permuted_mesh(70, 90, 11, dev)
    g = gw.build_graph(ei, n)
    plan = g.locality_plan()
    assert plan is not None and plan.flags == 1 and plan.run_len == 1
    assert plan.amplification < 2.0
    torch.manual_seed(2)
    x = torch.randn(b, n, f, device=dev).to(dtype)
    bias = torch.randn(f, device=dev) * 0.1
    y = ops.aggregate(g, x, bias, relu, kernel="locality")
    yr = ops.aggregate(g, x, bias, relu, kernel="rows")
    assert torch.equal(y, yr)
    # and against the CPU oracle (fp32: the row kernel is bit-exact vs scatter_add_ in CSR order; here a tolerance,
    # the oracle adds in edge_index order)
    ei2, ew, _ = orc.gcn_norm(ei.cpu(), n)
    ref = orc.propagate(x.float().cpu(), ei2, ew, n) + bias.cpu()
    if relu:
        ref = torch.relu(ref)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    err = ((y.float().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert err <= tol, err


def test_auto_picks_locality_tiles_and_falls_back(dev, monkeypatch):
    ei, n = permuted_mesh(64, 80, 3, dev)
    g = gw.build_graph(ei, n)
    x = torch.randn(n, 128, device=dev)
    monkeypatch.setattr(ops, "LOCALITY_MIN_NODES", 1000)
    y0 = ops.aggregate(g, x)                      # first call on a graph handle: row kernel, no plan is built
    assert ("locality", None) not in g._plans
    y = ops.aggregate(g, x)                       # second call: auto -> locality tiles
    assert ("locality", None) in g._plans and g._plans[("locality", None)] is not None
    assert torch.equal(y0, y)
    assert torch.equal(y, ops.aggregate(g, x, kernel="rows"))
    # a graph without locality (random edges): the plan is refused, auto falls back to the row kernel
    torch.manual_seed(9)
    er = orc.erdos_renyi_graph(2000, 0.01).to(dev)
    g2 = gw.build_graph(er, 2000)
    x2 = torch.randn(2000, 64, device=dev)
    ops.aggregate(g2, x2)
    y2 = ops.aggregate(g2, x2)
    assert g2._plans[("locality", None)] is None
    assert torch.equal(y2, ops.aggregate(g2, x2, kernel="rows"))
    with pytest.raises(RuntimeError):
        ops.aggregate(g2, x2, kernel="locality")


def test_locality_tiles_red_zone(dev):
    """out= destination with guard bands: the gather producer / consumers write nothing outside the rows."""
    ei, n = permuted_mesh(48, 64, 4, dev)
    g = gw.build_graph(ei, n)
    f = 96
    x = torch.randn(n, f, device=dev)
    buf = torch.full((n + 64, f), 7.5, device=dev)
    out = buf[32:32 + n]
    ops.aggregate(g, x, kernel="locality", out=out.unsqueeze(0))
    assert torch.all(buf[:32] == 7.5) and torch.all(buf[32 + n:] == 7.5)
    assert torch.equal(out, ops.aggregate(g, x, kernel="rows"))


def test_model_on_a_permuted_mesh_matches_oracle(dev, monkeypatch):
    """The six-layer model (fp32) on a mesh whose node ids are permuted: every layer's aggregation goes through
    kernel="auto" -> locality tiles; against the CPU oracle on the same permuted edge list, and equal to the model
    on the natural numbering with the rows permuted back (the layers are permutation-equivariant) within 1e-5."""
    from tests.golden import weights as wts
    monkeypatch.setattr(ops, "LOCALITY_MIN_NODES", 1000)
    h, w, c, hid = 48, 64, 16, 64
    n = h * w
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))
    ei_c = perm[orc.grid(h, w)].contiguous()
    ref = orc.GNNModelOracle(c, c, hid)
    wts.fill_model_(ref, 3)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = torch.randn(2, n, c, generator=torch.Generator().manual_seed(6))
    ei = ei_c.to(dev)
    with torch.no_grad():
        y = model(x.to(dev), ei)
        yr = ref(x, ei_c)
    g = gw.get_graph(ei, n)
    assert g.grid_shape is None and g._plans.get(("locality", None)) is not None     # the path under test ran
    err = ((y.cpu() - yr).abs().max() / yr.abs().max()).item()
    assert err <= 1e-5, err
    # backward through the same path (transposed graph -> its own locality plan)
    xg = x.to(dev).requires_grad_(True)
    model(xg, ei).square().sum().backward()
    xr = x.clone().requires_grad_(True)
    ref(xr, ei_c).square().sum().backward()
    gerr = ((xg.grad.cpu() - xr.grad).norm() / xr.grad.norm()).item()
    assert gerr <= 1e-4, gerr


def test_triangular_cells_unstructured_numbering(dev):
    """An ICON-like graph (triangular cells, three neighbours each) in a random numbering: the radius search finds
    tiles with few staged rows per row, the tiles are bit-exact vs the restatement, the aggregation is bitwise equal
    to the row kernel."""
    from tests.graphs import tri_mesh_edges
    h, w = 60, 80
    n = 2 * h * w
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(8))
    ei = perm[tri_mesh_edges(h, w)].contiguous().to(dev)
    g = gw.build_graph(ei, n)
    assert g.grid_shape is None
    _check_tiles_vs_oracle_deal(g, 12, 12, 200, 256, 148)
    plan = g.locality_plan()
    assert plan is not None and g.locality_radius > 8 and plan.amplification < 1.5
    x = torch.randn(2, n, 128, device=dev)
    assert torch.equal(ops.aggregate(g, x, kernel="locality"), ops.aggregate(g, x, kernel="rows"))
