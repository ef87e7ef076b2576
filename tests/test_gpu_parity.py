"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star / SURVEY.md 7.3):
  * graph construction, edge order, CSR, perm: bit-exact; dis: bit-exact vs the fp64-rounded
    oracle (<= 1 ulp from torch's pow(-0.5), pinned in test_oracle.py)
  * K1 aggregation fp32: BIT-EXACT vs CPU scatter_add_ order (same weights)
  * layer / model outputs fp32: max|y - y_ref| / max|y_ref| <= 1e-5
  * bf16: <= 2e-2 normalised vs the fp32 oracle on the same (bf16-rounded) inputs
"""
import os

import numpy as np
import pytest
import torch

import gwen_b200 as gw
from gwen_b200 import ops
from oracle import gcn_oracle as orc
from tests.golden import weights as wts

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def nmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def random_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g)


GRAPHS = {
    "k2": lambda: (orc.complete_graph(2), 2),
    "k125": lambda: (orc.complete_graph(125), 125),
    "grid3x4": lambda: (orc.grid(3, 4), 12),
    "grid17x23": lambda: (orc.grid(17, 23), 17 * 23),
    "grid_noloops": lambda: ((lambda e: e[:, e[0] != e[1]])(orc.grid(9, 8)), 72),
    "empty": lambda: (torch.empty((2, 0), dtype=torch.long), 6),
    "loops_dups": lambda: (torch.tensor([[0, 1, 1, 2, 2, 0, 3, 3], [1, 1, 2, 2, 0, 1, 3, 0]]), 5),
    "random": lambda: (random_graph(300, 4000, 1), 300),
    "random_sparse": lambda: (random_graph(1000, 700, 2), 1000),
}


# ---------------------------------------------------------------------------------------------
# builders + K0
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w", [(1, 1), (1, 7), (7, 1), (2, 2), (3, 4), (31, 18), (64, 100)])
def test_grid_builder_bit_exact(dev, h, w):
    assert torch.equal(gw.grid(h, w, dev).cpu(), orc.grid(h, w))


@pytest.mark.parametrize("n", [1, 2, 3, 125])
def test_complete_graph_bit_exact(dev, n):
    torch.manual_seed(7)
    a = gw.erdos_renyi_graph(n, 1, device=dev)
    nxt = torch.rand(1)
    torch.manual_seed(7)
    b = orc.erdos_renyi_graph(n, 1)
    assert torch.equal(a.cpu(), b)
    assert torch.equal(nxt, torch.rand(1))          # same RNG consumption as the reference builder


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_graph_build_bit_exact(dev, name):
    ei, n = GRAPHS[name]()
    g = gw.build_graph(ei.to(dev), n)
    rowptr, src, perm, dis = orc.dst_sorted_csr(ei, n)
    assert g.num_messages == len(src)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.src.cpu().numpy(), src)
    assert np.array_equal(g.perm.cpu().numpy(), perm)
    assert np.array_equal(g.dis.cpu().numpy(), dis)                       # bit-exact
    ei2, ew, _ = orc.gcn_norm(ei, n, dis_mode="exact")
    assert np.array_equal(g.w.cpu().numpy(), ew.numpy()[perm])            # bit-exact weights
    # transposed graph: same weights, segments by source
    gt = g.transposed()
    rt, st, pt, _ = orc.dst_sorted_csr(torch.stack([ei[1], ei[0]]), n)
    assert np.array_equal(gt.rowptr.cpu().numpy(), rt)
    assert np.array_equal(gt.src.cpu().numpy(), st)
    assert np.array_equal(gt.w.cpu().numpy(), ew.numpy()[pt])


def test_graph_build_rejects_out_of_range(dev):
    ei = torch.tensor([[0, 1, 9], [1, 0, 0]], device=dev)
    with pytest.raises(IndexError):
        gw.build_graph(ei, 3)
    with pytest.raises(ValueError):
        gw.build_graph(torch.zeros((3, 4), dtype=torch.long, device=dev), 3)


def test_grid_detection_and_cache(dev):
    ei = gw.grid(12, 9, dev)
    g = gw.get_graph(ei, 108)
    assert g.grid_shape == (12, 9)
    assert gw.get_graph(ei, 108) is g                                     # cache hit
    ei2 = ei.clone()
    assert gw.get_graph(ei2, 108) is not g                                # different memory
    ei2[0, 0] = 1                                                         # in-place edit bumps version
    g3 = gw.get_graph(ei2, 108)
    assert g3.grid_shape is None
    assert gw.build_graph(orc.complete_graph(12).to(dev), 12).grid_shape is None
    nl = ei[:, ei[0] != ei[1]]
    assert gw.build_graph(nl, 108).grid_shape == (12, 9)
    gw.clear_graph_cache()


# ---------------------------------------------------------------------------------------------
# K1 aggregation
# ---------------------------------------------------------------------------------------------
def oracle_aggregate(x, ei, n, bias=None, relu=False):
    ei2, ew, _ = orc.gcn_norm(ei, n, dis_mode="exact")
    out = orc.propagate(x, ei2, ew, n)
    if bias is not None:
        out = out + bias
    return torch.relu(out) if relu else out


@pytest.mark.parametrize("name", sorted(GRAPHS))
@pytest.mark.parametrize("feat", [4, 100, 256, 1028])
def test_aggregate_rows_fp32_bit_exact(dev, name, feat):
    ei, n = GRAPHS[name]()
    g = gw.build_graph(ei.to(dev), n)
    x = wts.features((n, feat), 3)
    b = wts.small_bias(feat, 4)
    out = ops.aggregate(g, x.to(dev), b.to(dev), relu=True, kernel="rows")
    assert torch.equal(out.cpu(), oracle_aggregate(x, ei, n, b, True))


@pytest.mark.parametrize("feat", [3, 7, 33])
def test_aggregate_scalar_path_bit_exact(dev, feat):
    ei, n = GRAPHS["random"]()
    g = gw.build_graph(ei.to(dev), n)
    x = wts.features((n, feat), 5)
    out = ops.aggregate(g, x.to(dev), kernel="rows")
    assert torch.equal(out.cpu(), oracle_aggregate(x, ei, n))


@pytest.mark.parametrize("hw,tile", [((17, 23), (8, 32)), ((40, 70), (8, 32)), ((40, 70), (4, 16)),
                                     ((33, 65), (16, 16)), ((9, 300), (2, 64))])
@pytest.mark.parametrize("feat,slab", [(64, 0), (256, 0), (256, 64), (384, 128), (1000, 32), (24, 0)])
def test_aggregate_tiled_fp32_bit_exact(dev, hw, tile, feat, slab):
    h, w = hw
    ei = orc.grid(h, w)
    g = gw.build_graph(ei.to(dev), h * w)
    assert g.grid_shape == (h, w)
    x = wts.features((h * w, feat), 7)
    b = wts.small_bias(feat, 8)
    ref = oracle_aggregate(x, ei, h * w, b, False)
    out = ops.aggregate(g, x.to(dev), b.to(dev), kernel="tiled", tile=tile, slab=slab)
    assert torch.equal(out.cpu(), ref)
    out_r = ops.aggregate(g, x.to(dev), b.to(dev), kernel="rows")
    assert torch.equal(out_r.cpu(), ref)


def test_aggregate_tiled_generic_graph(dev):
    ei, n = GRAPHS["random"]()
    g = gw.build_graph(ei.to(dev), n)
    x = wts.features((n, 128), 2)
    out = ops.aggregate(g, x.to(dev), kernel="tiled", tile=(64,), run_len=4)
    assert torch.equal(out.cpu(), oracle_aggregate(x, ei, n))
    plan = g.tile_plan((64,), 4)
    assert plan.num_tiles == 5 and plan.max_tile_runs * 4 <= n + 4


def test_tile_plan_structure(dev):
    h, w = 20, 50
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    plan = g.tile_plan((8, 32))
    assert plan.run_len == 34
    order = plan.order.cpu().numpy()
    assert sorted(order.tolist()) == list(range(h * w))
    tp = plan.tile_ptr.cpu().numpy()
    assert plan.num_tiles == 3 * 2 and tp[0] == 0 and tp[-1] == h * w
    first = order[tp[0]:tp[1]]
    assert set(first.tolist()) == {r * w + c for r in range(8) for c in range(32)}
    rp = plan.run_ptr.cpu().numpy()
    rs = plan.run_start.cpu().numpy()
    assert rs[rp[0]:rp[1]].tolist() == [r * w for r in range(9)]          # tile (0,0): 9 row segments
    assert rs[rp[1]:rp[2]].tolist() == [r * w + 31 for r in range(9)]     # tile (0,1): cols 31..49
    assert plan.max_tile_runs == 10
    assert 0.7 < plan.staging_efficiency <= 1.0          # border tiles stage 34-row boxes for 19-col runs
    # records + messages in processing order point at the right destination / source / weight
    rec = plan.trec.cpu().numpy()
    base = plan.tmsg_base.cpu().numpy()
    tm = plan.tmsg.cpu().numpy().view(np.uint64)
    li = (tm & np.uint64(0xFFFFFFFF)).astype(np.int64)
    wbits = (tm >> np.uint64(32)).astype(np.uint32).view(np.float32)
    rowptr, src, wcsr = g.rowptr.cpu().numpy(), g.src.cpu().numpy(), g.w.cpu().numpy()
    assert base[-1] == g.num_messages
    assert plan.max_tile_rows == 8 * 32 and plan.max_tile_msgs >= 9 * 8 * 32 - 200
    for t in range(plan.num_tiles):
        m0 = base[t] & ~1
        for p in range(tp[t], tp[t + 1]):
            d, moff, deg, _ = rec[p]
            assert d == order[p] and deg == rowptr[d + 1] - rowptr[d]
            for k in range(deg):
                q, s_ = m0 + moff + k, rowptr[d] + k
                run, off = divmod(li[q], plan.run_len)
                assert run < rp[t + 1] - rp[t]
                assert rs[rp[t] + run] + off == src[s_]
                assert wbits[q] == wcsr[s_]


@pytest.mark.parametrize("kernel", ["rows", "tiled"])
def test_aggregate_batched_and_bf16(dev, kernel):
    h, w, f = 24, 40, 128
    ei = orc.grid(h, w)
    g = gw.build_graph(ei.to(dev), h * w)
    x = wts.features((3, h * w, f), 9)
    out = ops.aggregate(g, x.to(dev), kernel=kernel)
    for i in range(3):
        assert torch.equal(out[i].cpu(), oracle_aggregate(x[i], ei, h * w))
    xb = x.to(torch.bfloat16)
    outb = ops.aggregate(g, xb.to(dev), kernel=kernel)
    ref = oracle_aggregate(xb.float(), ei, h * w)          # fp32 accumulate, one rounding at the end
    # bf16 path accumulates in fp32 with fused multiply-add: within 1 bf16 ulp of the rounded oracle
    err = (outb.float().cpu() - ref).abs()
    assert torch.all(err <= ref.abs() * 2.0 ** -7 + 1e-6)
    assert (outb.cpu() == ref.to(torch.bfloat16)).float().mean() > 0.98


def test_aggregate_is_deterministic(dev):
    g = gw.build_graph(gw.grid(64, 64, dev), 4096)
    x = torch.randn(4096, 256, device=dev)
    a = ops.aggregate(g, x, kernel="tiled")
    for _ in range(3):
        assert torch.equal(ops.aggregate(g, x, kernel="tiled"), a)
        assert torch.equal(ops.aggregate(g, x, kernel="rows"), a)


# ---------------------------------------------------------------------------------------------
# K2 linear (vs a plain PyTorch fp32 reference of the same op)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n", [(2, 100, 1024), (300, 64, 1024), (257, 1024, 512), (1000, 512, 256),
                                   (130, 37, 19), (1, 1, 1), (4096, 256, 256), (20000, 512, 1024),
                                   (70001, 1024, 256), (4096, 36, 64), (5000, 64, 192), (9000, 100, 128)])
def test_linear_fp32(dev, m, k, n):
    """fp32 projection: CUDA-core GEMM for small / ragged problems, 3xTF32 tensor-core GEMM (fp32-level
    accuracy) for m >= 4096 with k % 4 == 0 and n % 64 == 0 -- both within the fp32 parity bar."""
    x, w, b = wts.features((m, k), 1), wts.glorot(n, k, 2), wts.small_bias(n, 3)
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    y = ops.linear(x.to(dev), w.to(dev), b.to(dev), relu=True)
    assert nmax(y, ref) <= FP32_TOL
    y2 = ops.linear(x.to(dev), w.to(dev))
    assert nmax(y2, x.double() @ w.double().t()) <= FP32_TOL


@pytest.mark.parametrize("m,k,n", [(300, 64, 1024), (1000, 1024, 512), (640, 512, 256), (130, 40, 24),
                                   (257, 72, 96), (643, 1024, 64), (5, 256, 512), (4096, 256, 256),
                                   (20000, 128, 128), (40000, 192, 1024), (70001, 1024, 256)])
def test_linear_bf16(dev, m, k, n):
    """bf16 projection (tcgen05 path when K % 8 == 0 and N % 32 == 0, CUDA-core path otherwise)
    against a plain fp64 reference of the same bf16-rounded operands."""
    x, w, b = wts.features((m, k), 1).bfloat16(), wts.glorot(n, k, 2).bfloat16(), wts.small_bias(n, 3)
    ref = x.double() @ w.double().t() + b.double()
    y = ops.linear(x.to(dev), w.to(dev), b.to(dev))
    assert y.dtype == torch.bfloat16
    assert nmax(y, ref) <= 1e-2                           # one bf16 rounding of an fp32-accumulated sum
    # elementwise: |y - ref| <= 1 bf16 ulp of |ref| + fp32 accumulation slack
    err = (y.double().cpu() - ref).abs()
    assert torch.all(err <= ref.abs() * 2.0 ** -7 + 1e-3)
    yr = ops.linear(x.to(dev), w.to(dev), b.to(dev), relu=True)
    assert torch.equal(yr, torch.relu(y))


@pytest.mark.parametrize("m,k,n", [(300, 64, 128), (5000, 96, 40), (129, 256, 512), (6000, 512, 1024),
                                   (9000, 64, 100), (4096, 1024, 36)])
def test_linear_backward_pieces(dev, m, k, n):
    x, w, dy = wts.features((m, k), 1), wts.glorot(n, k, 2), wts.features((m, n), 3)
    dx = ops.linear_bwd_data(dy.to(dev), w.to(dev))
    assert nmax(dx, dy.double() @ w.double()) <= FP32_TOL
    dw = ops.linear_bwd_weight(dy.to(dev), x.to(dev))
    assert nmax(dw, dy.double().t() @ x.double()) <= FP32_TOL
    assert torch.equal(dw, ops.linear_bwd_weight(dy.to(dev), x.to(dev)))   # deterministic
    db = ops.bias_grad(dy.to(dev))
    assert nmax(db, dy.double().sum(0)) <= FP32_TOL
    y = wts.features((m, n), 4)
    d2 = ops.relu_bwd_(y.to(dev), dy.to(dev).clone())
    assert torch.equal(d2.cpu(), dy * (y > 0))
    # fused mask + bias gradient (one pass when the width allows, the two kernels otherwise)
    dz, db2 = ops.relu_bias_bwd(dy.to(dev), y.to(dev), True)
    assert torch.equal(dz.cpu(), dy * (y > 0))
    assert nmax(db2, (dy * (y > 0)).double().sum(0)) <= FP32_TOL
    dz0, db0 = ops.relu_bias_bwd(dy.to(dev), None, True)
    assert torch.equal(dz0.cpu(), dy) and nmax(db0, dy.double().sum(0)) <= FP32_TOL
    yb, dyb = y.bfloat16(), dy.bfloat16()
    dzb, dbb = ops.relu_bias_bwd(dyb.to(dev), yb.to(dev), True)
    assert torch.equal(dzb.cpu(), dyb * (yb > 0))
    assert nmax(dbb, (dyb * (yb > 0)).double().sum(0)) <= FP32_TOL
    assert torch.equal(dbb, ops.relu_bias_bwd(dyb.to(dev), yb.to(dev), True)[1])   # deterministic


@pytest.mark.parametrize("m,k,n", [(1000, 512, 1024), (40000, 256, 192), (300, 1024, 64), (70001, 128, 512),
                                   (5000, 64, 1024), (257, 96, 40)])
def test_linear_backward_pieces_bf16(dev, m, k, n):
    """bf16 dgrad / wgrad (tcgen05 kernels with MN-major operands where the shapes allow, CUDA-core
    path otherwise) against fp64 references of the same bf16-rounded operands."""
    x, w, dy = wts.features((m, k), 1).bfloat16(), wts.glorot(n, k, 2).bfloat16(), wts.features((m, n), 3).bfloat16()
    dx = ops.linear_bwd_data(dy.to(dev), w.to(dev))
    ref = dy.double() @ w.double()
    assert dx.dtype == torch.bfloat16
    err = (dx.double().cpu() - ref).abs()
    assert torch.all(err <= ref.abs() * 2.0 ** -7 + 1e-3 * ref.abs().max())
    dw = ops.linear_bwd_weight(dy.to(dev), x.to(dev))
    refw = dy.double().t() @ x.double()
    assert dw.dtype == torch.float32
    assert nmax(dw, refw) <= 1e-5                                           # fp32 accumulation of exact products
    assert torch.equal(dw, ops.linear_bwd_weight(dy.to(dev), x.to(dev)))   # deterministic


# ---------------------------------------------------------------------------------------------
# GCNConv layer and the full model
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GRAPHS))
@pytest.mark.parametrize("fi,fo", [(16, 40), (40, 16), (100, 1024)])
def test_gcnconv_layer_fp32(dev, name, fi, fo):
    ei, n = GRAPHS[name]()
    x, w, b = wts.features((n, fi), 1), wts.glorot(fo, fi, 2), wts.small_bias(fo, 3)
    conv = gw.GCNConv(fi, fo).to(dev)
    with torch.no_grad():
        conv.lin.weight.copy_(w)
        conv.bias.copy_(b)
    y = conv(x.to(dev), ei.to(dev))
    assert nmax(y, orc.gcn_conv_forward(x, ei, w, b)) <= FP32_TOL
    if n <= 400:
        assert nmax(y, orc.dense_gcn_forward(x, ei, w, b)) <= FP32_TOL
    assert nmax(conv(x.to(dev), ei.to(dev), relu=True), torch.relu(orc.gcn_conv_forward(x, ei, w, b))) <= FP32_TOL


def test_golden_layer(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "layer_grid_6x5.npz"))
    conv = gw.GCNConv(12, 20).to(dev)
    with torch.no_grad():
        conv.lin.weight.copy_(torch.from_numpy(g["w"]))
        conv.bias.copy_(torch.from_numpy(g["b"]))
    y = conv(torch.from_numpy(g["x"]).to(dev), gw.grid(6, 5, dev))
    assert nmax(y, torch.from_numpy(g["out"])) <= FP32_TOL
    assert nmax(y, torch.from_numpy(g["dense"])) <= FP32_TOL


def test_config1_reference_sample_full_model(dev, golden_dir):
    """BASELINE config 1: the reference's tests/test_data sample (K_2, x [2, 100], 2 time steps)
    through GNNModel(100 -> 1024 -> 512 -> 256 -> 512 -> 1024 -> 100)."""
    xs = torch.from_numpy(np.load(os.path.join(golden_dir, "cfg1_x.npy")))
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, "cfg1_out.npy")))
    cfg = gw.GNNConfig(nodes_in=2, nodes_out=2, channels_in=100, channels_out=100, hidden_feats=1024)
    model = gw.GNNModel(cfg)
    wts.fill_model_(model, 23)
    model = model.to(dev)
    ei = gw.erdos_renyi_graph(2, 1, device=dev)
    assert ei.tolist() == [[0, 1], [1, 0]]
    with torch.no_grad():
        for t in range(2):
            y = model(xs[t].to(dev), ei)
            assert nmax(y, gold[t]) <= FP32_TOL
            assert nmax(y[0], y[1]) <= FP32_TOL            # K_2: both rows equal (Appendix C.1)


@pytest.mark.parametrize("graph", ["grid", "k125"])
def test_full_model_fp32_and_bf16(dev, graph):
    if graph == "grid":
        ei, n = orc.grid(30, 26), 780
    else:
        ei, n = orc.complete_graph(125), 125
    c, hid = 64, 256
    ref = orc.GNNModelOracle(c, c, hid)
    wts.fill_model_(ref, 3)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = wts.features((2, n, c), 5)                         # ensemble members as the outer batch
    with torch.no_grad():
        yr = ref(x, ei)
        y = model(x.to(dev), ei.to(dev))
        assert nmax(y, yr) <= FP32_TOL
        yb = model.to(torch.bfloat16)(x.to(dev).bfloat16(), ei.to(dev))
        assert yb.dtype == torch.bfloat16
        assert nmax(yb.float(), yr) <= BF16_TOL


def test_backward_matches_oracle_autograd(dev):
    ei, n, c, hid = orc.grid(14, 11), 154, 24, 64
    ref = orc.GNNModelOracle(c, c, hid)
    wts.fill_model_(ref, 4)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = wts.features((n, c), 6)
    mask = torch.arange(n) % 5 == 4
    xr = x.clone().requires_grad_(True)
    lr = orc.loss_func(ref(xr, ei), x, mask)
    lr.backward()
    xd = x.to(dev).requires_grad_(True)
    ld = gw.loss_func(model(xd, ei.to(dev)), x.to(dev), mask.to(dev))
    ld.backward()
    assert abs(ld.item() - lr.item()) <= 1e-5 * abs(lr.item())
    assert nmax(xd.grad, xr.grad) <= 1e-4
    got = dict(model.named_parameters())
    for name, p in ref.named_parameters():
        if p.grad is None:
            assert got[name].grad is None                 # conv4/5, upconv1/2 never run
            continue
        assert nmax(got[name].grad, p.grad) <= 1e-4, name


def test_backward_non_symmetric_graph(dev):
    ei, n = GRAPHS["random"]()
    x, w, b = wts.features((n, 20), 1), wts.glorot(12, 20, 2), wts.small_bias(12, 3)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    orc.gcn_conv_forward(xr, ei, wr, br).square().sum().backward()
    for agg_first in (False, True):
        xd = x.to(dev).requires_grad_(True)
        wd = w.to(dev).requires_grad_(True)
        bd = b.to(dev).requires_grad_(True)
        g = gw.build_graph(ei.to(dev), n)
        gw.gcn_conv(xd, g, wd, bd, agg_first=agg_first).square().sum().backward()
        assert nmax(xd.grad, xr.grad) <= 1e-4
        assert nmax(wd.grad, wr.grad) <= 1e-4
        assert nmax(bd.grad, br.grad) <= 1e-4


def test_large_grid_properties(dev):
    """BASELINE config 2 size (582 x 390, F = 256): size-independent checks -- a row band against
    the oracle, linearity, constant-input invariance and rows == tiled bitwise."""
    h, w, f = 582, 390, 256
    n = h * w
    ei = gw.grid(h, w, dev)
    assert ei.size(1) == 2036992
    g = gw.get_graph(ei, n)
    assert g.grid_shape == (h, w) and g.num_messages == 2036992
    x = torch.randn(n, f, device=dev)
    a = ops.aggregate(g, x, kernel="tiled")
    assert torch.equal(a, ops.aggregate(g, x, kernel="rows"))
    # band check: grid rows 100..103 against the oracle on the 8-row sub-grid 98..105 (the
    # compared rows and all their neighbours have the same degrees in both graphs)
    sub = x[98 * w:106 * w].cpu()
    ref = oracle_aggregate(sub, orc.grid(8, w), 8 * w)
    assert torch.equal(a[100 * w:104 * w].cpu(), ref[2 * w:6 * w])
    # linearity (exact for scaling by powers of two)
    assert torch.equal(ops.aggregate(g, x * 2.0, kernel="tiled"), a * 2.0)
    # A_hat 1 = D^-1/2 (A+I) D^-1/2 1: interior rows sum nine weights of 1/9
    ones = ops.aggregate(g, torch.ones(n, 4, device=dev), kernel="rows")
    assert abs(ones[200 * w + 100, 0].item() - 1.0) < 1e-6
    gw.clear_graph_cache()


def test_tile_range_launches_cover_the_plan(dev):
    h, w, f = 40, 50, 64
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    x = torch.randn(h * w, f, device=dev)
    full = ops.aggregate(g, x, kernel="tiled", tile=(8, 16))
    plan = g.tile_plan((8, 16))
    tc = 4
    assert plan.num_tiles == 5 * tc
    out = torch.full_like(full, float("nan"))
    ops.aggregate(g, x, kernel="tiled", tile=(8, 16), out=out, tile_range=(tc, 3 * tc))
    assert torch.isnan(out[:8 * w]).all() and torch.isnan(out[32 * w:]).all()
    assert torch.equal(out[8 * w:32 * w], full[8 * w:32 * w])
    ops.aggregate(g, x, kernel="tiled", tile=(8, 16), out=out, tile_range=(0, tc))
    ops.aggregate(g, x, kernel="tiled", tile=(8, 16), out=out, tile_range=(4 * tc, tc))
    assert torch.equal(out, full)
    with pytest.raises(RuntimeError):
        ops.aggregate(g, x, kernel="tiled", tile=(8, 16), out=out, tile_range=(18, 5))


def test_band_plan_matches_plain_plan(dev):
    """The partition's boundary-first tile layout (1-row strips + interior tiles) gives the same
    result as the plain plan, also when launched as interior / boundary ranges."""
    from gwen_b200 import partition

    class _NoExchange:
        send_idx = {}

        def exchange(self, x):
            return x
    h, w, f = 37, 50, 128
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    lg = partition.LocalGraph(g, h * w, torch.empty(0, dtype=torch.int64, device=dev), range(0, h * w))
    band = partition.BandAggregator(lg, _NoExchange(), tile=(8, 16))
    assert band.n_boundary == 2 * 2 and band.n_interior == 5 * 4
    x = torch.randn(h * w, f, device=dev)
    b = torch.randn(f, device=dev)
    ref = ops.aggregate(g, x, b, kernel="rows")
    assert torch.equal(band(x, b), ref)
    assert torch.equal(ops.aggregate(g, x, b, kernel="tiled", plan=band.plan), ref)


# ---------------------------------------------------------------------------------------------
# K1s mesh stencil fast path
# ---------------------------------------------------------------------------------------------
STENCIL_TOL = 1e-6   # fp32: same terms, different (separable) summation order


@pytest.mark.parametrize("hw", [(2, 2), (3, 4), (8, 32), (9, 33), (17, 23), (40, 70), (5, 130), (64, 7)])
@pytest.mark.parametrize("feat,slab,tw", [(64, 0, 0), (256, 0, 0), (256, 128, 16), (36, 32, 8), (520, 64, 64)])
def test_stencil_fp32_vs_oracle(dev, hw, feat, slab, tw):
    h, w = hw
    ei = orc.grid(h, w)
    g = gw.build_graph(ei.to(dev), h * w)
    assert g.is_plain_mesh
    x = wts.features((h * w, feat), 11)
    b = wts.small_bias(feat, 12)
    ref = oracle_aggregate(x, ei, h * w, b, True)
    out = ops.aggregate(g, x.to(dev), b.to(dev), relu=True, kernel="stencil", slab=slab,
                        tile=(tw,) if tw else None)
    assert nmax(out, ref) <= STENCIL_TOL
    dense = torch.relu(orc.dense_norm_adj(ei, h * w) @ x.double() + b.double()) if h * w <= 1000 else None
    if dense is not None:
        assert nmax(out, dense) <= STENCIL_TOL
    # the graph without its explicit self loops normalises to the same operator
    g2 = gw.build_graph(ei[:, ei[0] != ei[1]].to(dev), h * w)
    assert g2.is_plain_mesh
    assert torch.equal(ops.aggregate(g2, x.to(dev), b.to(dev), relu=True, kernel="stencil", slab=slab,
                                     tile=(tw,) if tw else None), out)


def test_stencil_batched_bf16_deterministic_and_auto(dev):
    h, w, f = 30, 45, 128
    ei = orc.grid(h, w)
    g = gw.build_graph(ei.to(dev), h * w)
    x = wts.features((3, h * w, f), 9)
    out = ops.aggregate(g, x.to(dev), kernel="stencil")
    for i in range(3):
        assert nmax(out[i], oracle_aggregate(x[i], ei, h * w)) <= STENCIL_TOL
    assert torch.equal(out, ops.aggregate(g, x.to(dev), kernel="stencil"))          # run-to-run
    assert torch.equal(out, ops.aggregate(g, x.to(dev)))                             # auto -> stencil
    assert torch.equal(out, ops.aggregate(g, x.to(dev), kernel="stencil", tile=(16,), slab=32))  # tiling-independent
    assert nmax(out, ops.aggregate(g, x.to(dev), kernel="tiled")) <= STENCIL_TOL
    xb = x.to(torch.bfloat16)
    outb = ops.aggregate(g, xb.to(dev), kernel="stencil")
    ref = oracle_aggregate(xb.float(), ei, h * w)
    assert torch.all((outb.float().cpu() - ref).abs() <= ref.abs() * 2.0 ** -7 + 1e-6)
    # not a plain mesh -> refused / auto falls back to the CSR kernel
    gi = gw.build_graph(ei.to(dev), h * w, improved=True)
    assert not gi.is_plain_mesh
    with pytest.raises(RuntimeError):
        ops.aggregate(gi, x.to(dev), kernel="stencil")
    gr = gw.build_graph(GRAPHS["random"]()[0].to(dev), 300)
    assert not gr.is_plain_mesh


def test_stencil_large_mesh(dev):
    """BASELINE config 2 size: stencil vs the bit-exact tiled kernel, and a band vs the oracle."""
    h, w, f = 582, 390, 256
    g = gw.get_graph(gw.grid(h, w, dev), h * w)
    x = torch.randn(h * w, f, device=dev)
    b = torch.randn(f, device=dev)
    a = ops.aggregate(g, x, b, kernel="stencil")
    assert nmax(a, ops.aggregate(g, x, b, kernel="tiled")) <= STENCIL_TOL
    sub = x[98 * w:106 * w].cpu()
    ref = oracle_aggregate(sub, orc.grid(8, w), 8 * w, b.cpu())
    assert nmax(a[100 * w:104 * w], ref[2 * w:6 * w]) <= STENCIL_TOL
    gw.clear_graph_cache()


@pytest.mark.parametrize("world", [2, 3])
def test_mesh_band_stencil_equals_single_gpu(dev, world):
    """Row-band partition of the stencil path, all ranks emulated on one GPU: every band's output
    (whole-band launch and the interior / first-row / last-row split) is BITWISE equal to the
    single-GPU stencil on the whole mesh."""
    from gwen_b200 import partition
    h, w, f = 41, 50, 64
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    x = torch.randn(2, h * w, f, device=dev)
    b = torch.randn(f, device=dev)
    full = ops.aggregate(g, x, b, relu=True, kernel="stencil")
    for rank in range(world):
        band = partition.MeshBand(h, w, g.dis, rank=rank, world=world)
        xl = band.alloc(2, f, torch.float32, dev)
        lo, hi = max(band.r0 - 1, 0), min(band.r0 + band.rows + 1, h)
        xl[:, (lo - (band.r0 - 1)) * w:(hi - (band.r0 - 1)) * w] = x[:, lo * w:hi * w]   # halos as exchanged
        want = full[:, band.r0 * w:(band.r0 + band.rows) * w]
        rows = band.rows
        out = ops.mesh_stencil(xl, band.dis, rows + 2, rows, w, 1, bias=b, relu=True)
        assert torch.equal(out, want)
        out2 = torch.full_like(out, float("nan"))
        ops.mesh_stencil(xl, band.dis, rows + 2, rows - 2, w, 2, bias=b, relu=True, out=out2[:, w:(rows - 1) * w])
        ops.mesh_stencil(xl, band.dis, rows + 2, 1, w, 1, bias=b, relu=True, out=out2[:, :w])
        ops.mesh_stencil(xl, band.dis, rows + 2, 1, w, rows, bias=b, relu=True, out=out2[:, (rows - 1) * w:])
        assert torch.equal(out2, want)


@pytest.mark.parametrize("chunks", [1, 5, 64])
def test_host_propagator_matches_device_path(dev, chunks):
    """Chunked H2D -> stencil -> D2H pipeline on pinned host tensors == the device-resident
    aggregation, bitwise, over consecutive calls (double-buffered device sets)."""
    h, w, f = 37, 29, 64
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    hp = gw.HostPropagator(g, f, torch.float32, chunks=chunks)
    b = wts.small_bias(f, 3).to(dev)
    outs, refs = [], []
    for i in range(3):
        xh = wts.features((h * w, f), 20 + i).pin_memory()
        oh = torch.empty(h * w, f).pin_memory()
        hp(xh, oh, b, relu=True)
        outs.append(oh)
        refs.append(ops.aggregate(g, xh.to(dev), b, relu=True, kernel="stencil"))
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o, r.cpu())
    with pytest.raises(RuntimeError):
        hp(torch.empty(h * w, f), outs[0])          # not pinned


# ---------------------------------------------------------------------------------------------
# training-step glue: fused masked L1 loss, train_step
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(154, 24), (3, 154, 24), (1000, 100), (2, 70, 1), (2, 333, 64)])
def test_masked_l1_loss_matches_reference_loss_func(dev, shape):
    n = shape[-2]
    y, t = wts.features(shape, 1), wts.features(shape, 2)
    t.view(-1)[::7] = y.view(-1)[::7]                              # exact ties: sign(0) = 0
    mask = torch.arange(n) % 5 == 4
    yr = y.clone().requires_grad_(True)
    lr = orc.loss_func(yr, t, mask) if y.dim() == 2 else torch.nn.L1Loss()(yr[:, mask], t[:, mask])
    (lr * 3.0).backward()
    yd = y.to(dev).requires_grad_(True)
    ld = gw.masked_l1_loss(yd, t.to(dev), mask.to(dev))
    (ld * 3.0).backward()
    assert abs(ld.item() - lr.item()) <= 1e-6 * abs(lr.item())
    assert nmax(yd.grad, yr.grad) <= 1e-6
    assert torch.equal(ld, gw.masked_l1_loss(yd, t.to(dev), mask.to(dev)))   # deterministic
    # bf16 inputs: same value as the fp32 formula on the rounded inputs
    yb, tb = y.bfloat16(), t.bfloat16()
    lb = gw.masked_l1_loss(yb.to(dev), tb.to(dev), mask.to(dev))
    ref_b = (yb.float() - tb.float()).abs()[..., mask, :].mean()
    assert abs(lb.item() - ref_b.item()) <= 1e-5 * abs(ref_b.item())
    # bf16 gradient (16-byte vector kernel when the width allows, scalar otherwise): sign(y - t) * mask * scale
    ybd = yb.to(dev).requires_grad_(True)
    (gw.masked_l1_loss(ybd, tb.to(dev), mask.to(dev)) * 3.0).backward()
    cnt = int(mask.sum()) * (shape[0] if len(shape) == 3 else 1) * shape[-1]
    gref = (torch.sign(yb.float() - tb.float()) * mask.view(-1, 1) * (3.0 / cnt)).bfloat16()
    assert torch.equal(ybd.grad.cpu(), gref)
    # empty mask -> NaN, as the mean of an empty selection
    assert torch.isnan(gw.masked_l1_loss(yd, t.to(dev), torch.zeros(n, dtype=torch.bool, device=dev)))


def test_train_step_matches_oracle_sgd(dev):
    """Two iterations of the reference inner loop (zero_grad / forward / loss vs the input / backward /
    optimizer.step) with our layers + fused loss == the oracle model + reference loss_func on the CPU."""
    ei, n, c, hid = orc.grid(14, 11), 154, 24, 64
    ref = orc.GNNModelOracle(c, c, hid)
    wts.fill_model_(ref, 4)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    x = wts.features((n, c), 6)
    mask = torch.arange(n) % 5 == 4
    opt_r = torch.optim.SGD(ref.parameters(), lr=0.05)
    opt_d = torch.optim.SGD(model.parameters(), lr=0.05)
    for _ in range(2):
        opt_r.zero_grad()
        lr = orc.loss_func(ref(x, ei), x, mask)
        lr.backward()
        opt_r.step()
        ld = gw.train_step(model, x.to(dev), ei.to(dev), mask.to(dev), opt_d)
        assert abs(ld.item() - lr.item()) <= 1e-5 * abs(lr.item())
    got = dict(model.named_parameters())
    for name, p in ref.named_parameters():
        assert nmax(got[name].detach(), p.detach()) <= 1e-5, name


# ---------------------------------------------------------------------------------------------
# K1 + K2 fused layer kernel (mesh graphs, bf16)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(8, 32), (9, 33), (17, 23), (40, 70), (3, 5), (64, 130)])
@pytest.mark.parametrize("k,n,batch", [(64, 1024, 1), (256, 512, 2), (192, 1024, 1), (128, 128, 3), (256, 256, 1)])
def test_fused_layer_matches_two_kernel_path(dev, hw, k, n, batch):
    """gwen_gcn_fused_fwd == stencil aggregation followed by the tcgen05 projection (same bf16
    rounding of the aggregated rows, same MMA order): bitwise on the mesh, incl. tiles that overhang
    it; and within bf16 tolerance of the fp32 oracle."""
    h, w = hw
    ei = orc.grid(h, w)
    g = gw.build_graph(ei.to(dev), h * w)
    x = wts.features((batch, h * w, k), 31).bfloat16()
    wt, b = wts.glorot(n, k, 32).bfloat16(), wts.small_bias(n, 33)
    assert ops.gcn_fused_supported(g, x.to(dev), wt.to(dev))
    y = ops.gcn_fused(g, x.to(dev), wt.to(dev), b.to(dev), relu=True)
    hh = ops.aggregate(g, x.to(dev), kernel="stencil")
    y2 = ops.linear(hh, wt.to(dev), b.to(dev), relu=True)
    assert y.shape == y2.shape == (batch, h * w, n)
    assert torch.equal(y, y2)
    assert torch.equal(y, ops.gcn_fused(g, x.to(dev), wt.to(dev), b.to(dev), relu=True))      # deterministic
    if h * w <= 1500:
        ref = torch.relu(oracle_aggregate(x[0].float(), ei, h * w) @ wt.float().t() + b)
        assert nmax(y[0].float(), ref) <= BF16_TOL
    # no bias, no relu
    assert torch.equal(ops.gcn_fused(g, x.to(dev), wt.to(dev)), ops.linear(hh, wt.to(dev)))


def test_fused_layer_autograd_matches_unfused(dev, monkeypatch):
    from gwen_b200 import nn as gnn
    monkeypatch.setattr(ops, "FUSED_MIN_ITEMS", 0)        # force the fused choice on a small mesh
    monkeypatch.setattr(gnn, "FUSED_IN_TRAINING", True)   # ... also when gradients are wanted (default: two kernels)
    h, w, k, n = 21, 19, 64, 256
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    x = wts.features((2, h * w, k), 41).bfloat16().to(dev)
    conv = gw.GCNConv(k, n).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        conv.bias.copy_(wts.small_bias(n, 42))
    t = wts.features((2, h * w, n), 43).bfloat16().to(dev)

    def run():
        conv.zero_grad()
        xx = x.clone().requires_grad_(True)
        y = conv(xx, g, relu=True)
        (y.float() * t.float()).sum().backward()
        return y.detach(), xx.grad, conv.lin.weight.grad.clone(), conv.bias.grad.clone()
    a = run()
    monkeypatch.setenv("GWEN_NO_FUSED", "1")
    bref = run()
    for u, v in zip(a, bref):
        assert torch.equal(u, v)


def test_graph_dataset_matches_reference_tensorisation(dev):
    """GraphDataset (utils.py:164-211 mirror) on the reference's own sample data: x of every time step
    equals the golden stacking, the member split follows np.random.shuffle, the graph is K_members."""
    import numpy as np
    gx = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_x.npy"))     # [2, 2, 100]
    theta = gx.reshape(2, 2, 10, 10)                                                  # t, member, h, cell
    np.random.seed(23)
    want = np.arange(2)
    np.random.shuffle(want)
    np.random.seed(23)
    torch.manual_seed(23)
    ds = gw.GraphDataset(theta, split=1, device=dev)
    assert ds.len() == 2 and ds.nodes == 2 and ds.channels == 100
    assert list(ds.input_indices) == list(want[:1]) and list(ds.target_indices) == list(want[1:])
    assert ds.edge_index.cpu().tolist() == [[0, 1], [1, 0]]
    for t in range(2):
        d = ds.get(t)
        assert torch.equal(d.x.cpu(), torch.from_numpy(gx[t]))
        assert d.target_mask.cpu().tolist() == [i in set(ds.target_indices) for i in range(2)]
    # and the sample flows through the model + train step
    cfg = gw.GNNConfig(nodes_in=2, nodes_out=2, channels_in=100, channels_out=100, hidden_feats=64)
    model = gw.GNNModel(cfg).to(dev)
    d = ds.get(0)
    loss = gw.train_step(model, d.x, d.edge_index, d.target_mask)
    assert torch.isfinite(loss)


# ---------------------------------------------------------------------------------------------
# BASELINE full sizes through size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(1158, 774), (2048, 2048)])
def test_full_size_mesh_operator_properties(dev, hw):
    """Config 3 / config 4 meshes (too large for the CPU oracle in seconds): the normalised mesh
    operator must (a) give row sums dis[d] * sum_{s in N(d)} dis[s] on a constant input -- checked
    against degrees known in closed form, (b) be self-adjoint <A x, y> = <x, A y>, (c) be linear, and
    (d) agree between the stencil kernel and the generic CSR kernel on the same graph handle."""
    h, w = hw
    n, f = h * w, 8
    g = gw.get_graph(gw.grid(h, w, dev), n)
    assert g.is_plain_mesh and g.num_messages == gw.grid_edge_count(h, w)
    # (a) closed-form degrees of the 8-neighbour mesh incl. self loop: 4 corners, 6 edges, 9 interior
    r = torch.arange(h, device=dev).view(-1, 1)
    c = torch.arange(w, device=dev).view(1, -1)
    deg = ((1 + (r > 0).long() + (r < h - 1).long()) * (1 + (c > 0).long() + (c < w - 1).long())).double()
    dis = deg.pow(-0.5)
    assert torch.equal(g.dis.view(h, w), dis.float())                      # K0: fp64 1/sqrt(deg) rounded once
    ones = torch.ones(n, f, device=dev)
    pad = torch.nn.functional.pad(dis, (1, 1, 1, 1))
    box = sum(pad[1 + dr:h + 1 + dr, 1 + dc:w + 1 + dc] for dr in (-1, 0, 1) for dc in (-1, 0, 1))
    want = (dis * box).reshape(n, 1)
    got = ops.aggregate(g, ones, kernel="stencil")
    assert nmax(got[:, :1], want) <= STENCIL_TOL
    # (b) adjointness and (c) linearity on random vectors, fp64 inner products
    gen = torch.Generator(dev).manual_seed(3)
    x = torch.randn(n, f, device=dev, generator=gen)
    y = torch.randn(n, f, device=dev, generator=gen)
    ax, ay = ops.aggregate(g, x, kernel="stencil"), ops.aggregate(g, y, kernel="stencil")
    lhs, rhs = (ax.double() * y.double()).sum(), (x.double() * ay.double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)
    axy = ops.aggregate(g, 2.0 * x - 0.5 * y, kernel="stencil")
    assert nmax(axy, 2.0 * ax.double() - 0.5 * ay.double()) <= 2e-6
    # (d) stencil vs the bit-exact CSR kernel
    assert nmax(ax, ops.aggregate(g, x, kernel="rows")) <= STENCIL_TOL
    gw.clear_graph_cache()


def test_linear_into_strided_rows(dev):
    """ops.linear / linear_bwd_data writing into the owned rows of a larger batch-strided buffer (how the
    band model stages activations without a copy) == the plain call."""
    b, n, k, m = 3, 700, 64, 128
    x = wts.features((b, n, k), 51).bfloat16().to(dev)
    w = wts.glorot(m, k, 52).bfloat16().to(dev)
    bias = wts.small_bias(m, 53).to(dev)
    big = torch.full((b, n + 40, m), 7.0, dtype=torch.bfloat16, device=dev)
    view = big[:, 20:20 + n, :]
    y = ops.linear(x, w, bias, relu=True, out=view)
    assert y.data_ptr() == view.data_ptr()
    assert torch.equal(view, ops.linear(x, w, bias, relu=True))
    assert torch.all(big[:, :20] == 7.0) and torch.all(big[:, 20 + n:] == 7.0)       # nothing else touched
    # strided INPUT as well
    y2 = ops.linear(view, wts.glorot(k, m, 54).bfloat16().to(dev))
    assert torch.equal(y2, ops.linear(view.contiguous(), wts.glorot(k, m, 54).bfloat16().to(dev)))
    dy = wts.features((b, n, m), 55).bfloat16().to(dev)
    gbig = torch.zeros((b, n + 8, k), dtype=torch.bfloat16, device=dev)
    ops.linear_bwd_data(dy, w, out=gbig[:, 8:, :])
    assert torch.equal(gbig[:, 8:, :], ops.linear_bwd_data(dy, w))
    assert torch.all(gbig[:, :8] == 0)


def test_mixed_precision_fp32_master_weights(dev):
    """BASELINE config 5 runs bf16 activations with fp32 master weights: an fp32-parameter model fed
    bf16 features computes exactly what the bf16-cast model computes (weights are rounded to bf16 per
    call), and its parameter gradients come back in fp32."""
    h, w, c, hid = 12, 11, 64, 128
    ei = gw.grid(h, w, dev)
    cfg = gw.GNNConfig(nodes_in=h * w, nodes_out=h * w, channels_in=c, channels_out=c, hidden_feats=hid)
    torch.manual_seed(5)
    master = gw.GNNModel(cfg).to(dev)                       # fp32 parameters
    import copy
    cast = copy.deepcopy(master).to(torch.bfloat16)
    x = wts.features((2, h * w, c), 61).bfloat16().to(dev)
    mask = (torch.arange(h * w, device=dev) % 3) == 1
    y = master(x, ei)
    assert y.dtype == torch.bfloat16 and torch.equal(y, cast(x, ei))
    loss = gw.masked_l1_loss(y, x, mask)
    loss.backward()
    lc = gw.masked_l1_loss(cast(x, ei), x, mask)
    lc.backward()
    for (name, p), (_, q) in zip(master.named_parameters(), cast.named_parameters()):
        if p.grad is None:
            assert q.grad is None
            continue
        assert p.grad.dtype == torch.float32 and q.grad.dtype == torch.bfloat16
        # same fp32 gradient, only the final cast differs
        assert torch.equal(p.grad.to(torch.bfloat16), q.grad), name


def test_relu_bias_bwd_strided(dev):
    """The fused mask + bias pass on batch-strided y / dy / out (band buffers) == the dense call."""
    b, n, f = 3, 500, 128
    y = wts.features((b, n, f), 71).bfloat16().to(dev)
    dy = wts.features((b, n, f), 72).bfloat16().to(dev)
    dz_ref, db_ref = ops.relu_bias_bwd(dy, y, True)
    ybig = torch.zeros(b, n + 30, f, dtype=torch.bfloat16, device=dev)
    ybig[:, 10:10 + n] = y
    obig = torch.full((b, n + 16, f), 3.0, dtype=torch.bfloat16, device=dev)
    dz, db = ops.relu_bias_bwd(dy, ybig[:, 10:10 + n], True, out=obig[:, 16:])
    assert dz.data_ptr() == obig[:, 16:].data_ptr()
    assert torch.equal(obig[:, 16:], dz_ref) and torch.all(obig[:, :16] == 3.0)
    assert nmax(db, db_ref.double()) <= 1e-6
    dz2, db2 = ops.relu_bias_bwd(dy, None, True, out=obig[:, 16:])          # no mask: copy + column sums
    assert torch.equal(obig[:, 16:], dy) and nmax(db2, dy.double().sum((0, 1))) <= 1e-5
    dz3, db3 = ops.relu_bias_bwd(dy, None, False)
    assert dz3.data_ptr() == dy.data_ptr() and db3 is None


# ---------------------------------------------------------------------------------------------
# red zones: kernels write exactly their output (compute-sanitizer is closed on this pool)
# ---------------------------------------------------------------------------------------------
def _guarded(shape, dtype, dev, pad=64):
    """(view, check): an output view of `shape` inside a larger sentinel-filled buffer."""
    n = 1
    for s in shape:
        n *= s
    big = torch.full((n + 2 * pad,), 1234.0, dtype=dtype, device=dev)
    view = big[pad:pad + n].view(shape)

    def check():
        assert torch.all(big[:pad] == 1234.0) and torch.all(big[pad + n:] == 1234.0), "wrote outside its output"
    return view, check


@pytest.mark.parametrize("hw", [(3, 5), (9, 33), (17, 23), (40, 70)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_red_zones_aggregation_and_fused(dev, hw, dt):
    h, w = hw
    g = gw.build_graph(gw.grid(h, w, dev), h * w)
    for f in (64, 264):
        x = wts.features((2, h * w, f), 81).to(dt).to(dev)
        b = wts.small_bias(f, 82).to(dev)
        for kern in ("rows", "tiled", "stencil"):
            out, check = _guarded((2, h * w, f), dt, dev)
            ops.aggregate(g, x, b, relu=True, kernel=kern, out=out)
            torch.cuda.synchronize()
            check()
            assert nmax(out.float(), ops.aggregate(g, x, b, relu=True, kernel="rows").float()) <= (1e-6 if dt == torch.float32 else 2e-2)
    if dt == torch.bfloat16:
        x = wts.features((2, h * w, 128), 83).bfloat16().to(dev)
        wt = wts.glorot(384, 128, 84).bfloat16().to(dev)
        out, check = _guarded((2, h * w, 384), dt, dev)
        ops.gcn_fused(g, x, wt, wts.small_bias(384, 85).to(dev), relu=True, out=out)
        torch.cuda.synchronize()
        check()


@pytest.mark.parametrize("m,k,n", [(300, 64, 128), (257, 72, 96), (643, 1024, 64), (4100, 36, 64), (4500, 100, 128)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_red_zones_projections(dev, m, k, n, dt):
    x = wts.features((m, k), 91).to(dt).to(dev)
    wt = wts.glorot(n, k, 92).to(dt).to(dev)
    dy = wts.features((m, n), 93).to(dt).to(dev)
    out, check = _guarded((m, n), dt, dev)
    ops.linear(x, wt, wts.small_bias(n, 94).to(dev), relu=True, out=out)
    gout, gcheck = _guarded((m, k), dt, dev)
    ops.linear_bwd_data(dy, wt, out=gout)
    zout, zcheck = _guarded((m, n), dt, dev)
    ops.relu_bias_bwd(dy, out.clone(), True, out=zout)
    torch.cuda.synchronize()
    check(), gcheck(), zcheck()
    assert torch.equal(out, ops.linear(x, wt, wts.small_bias(n, 94).to(dev), relu=True))
    assert torch.equal(gout, ops.linear_bwd_data(dy, wt))


def test_eval_step_matches_oracle(dev):
    """eval_step (reference eval loop body, models_gnn.py:440-450): loss vs the oracle model + loss_func on
    the CPU, and the kept prediction output[1]."""
    ei, n, c, hid = orc.complete_graph(7), 7, 12, 64
    ref = orc.GNNModelOracle(c, c, hid)
    wts.fill_model_(ref, 9)
    cfg = gw.GNNConfig(nodes_in=n, nodes_out=n, channels_in=c, channels_out=c, hidden_feats=hid)
    model = gw.GNNModel(cfg)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev).eval()
    x = wts.features((n, c), 10)
    mask = torch.tensor([0, 1, 0, 1, 1, 0, 0], dtype=torch.bool)
    with torch.no_grad():
        yr = ref(x, ei)
        lr = orc.loss_func(yr, x, mask)
    loss, kept = gw.eval_step(model, x.to(dev), ei.to(dev), mask.to(dev))
    assert abs(loss.item() - lr.item()) <= 1e-5 * abs(lr.item())
    assert nmax(kept, yr[1]) <= FP32_TOL and not loss.requires_grad
