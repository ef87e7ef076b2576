"""CPU oracle for the GWEN GCN message-passing hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``gwen_b200/`` imports it, and the product path
raises when the CUDA library is missing instead of falling back to this code.

What it restates
----------------
The reference's hot-path arithmetic is not in the reference tree: it lives in the
un-vendored dependency ``torch-geometric==2.3.1`` (reference
``requirements/environment.yml:552``), which is absent from this image.  This module
restates the published PyG-2.3.1 algorithm with plain ``torch`` CPU ops
(``index_select`` / ``scatter_add_`` / ``F.linear``) so the CPU baseline executes the
same ATen kernels PyG dispatches to, anchored on the reference's own call sites:

* ``GCNConv(in, out)(x, edge_index)``   <- reference ``src/gwen/models_gnn.py:118-130,172-184``
                                          (constructors) and ``:147-149,204-206`` (calls)
* ``erdos_renyi_graph(N, edge_prob=1)`` <- reference ``src/gwen/utils.py:176``
* model wiring (6 live layers, ReLU x5)  <- reference ``src/gwen/models_gnn.py:106-303``
* masked L1 loss                         <- reference ``src/gwen/models_gnn.py:261-265``

PARITY PINNING: the reference's tests hold **no** golden vector or known-answer test
for this path (``tests/test_gwen/test_models.py:19,36`` mock the conv stack), and the
third-party module cannot be imported here, so against the *reference's own fixtures*
this oracle is "parity unpinned".  It is pinned instead by (a) analytic known-answer
tests (K_2, K_N, isolated nodes, duplicate/self-loop edges,
grid(3,4)), and (b) an independent dense fp64 restatement ``dense_norm_adj`` /
``dense_gcn_forward`` written from the GCN formula rather than from the PyG op
sequence; see ``tests/test_oracle.py``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "erdos_renyi_graph", "complete_graph", "grid", "add_remaining_self_loops", "gcn_norm",
    "gcn_conv_forward", "propagate", "GCNConvOracle", "GNNModelOracle", "loss_func",
    "dense_norm_adj", "dense_gcn_forward", "dst_sorted_csr", "exact_dis",
]


# --------------------------------------------------------------------------------------
# Graph builders (edge ORDER is part of the parity contract: SURVEY.md Appendix B)
# --------------------------------------------------------------------------------------
def _coalesce(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG ``coalesce`` without edge attributes: sort by (row, col), drop duplicates."""
    key = edge_index[0] * num_nodes + edge_index[1]
    key = torch.unique(key, sorted=True)
    return torch.stack([key // num_nodes, key % num_nodes], dim=0)


def erdos_renyi_graph(num_nodes: int, edge_prob: float, directed: bool = False) -> torch.Tensor:
    """PyG-2.3.1 ``erdos_renyi_graph`` as called at reference ``src/gwen/utils.py:176``.

    Pairs (i<j) from ``torch.combinations`` are kept where ``torch.rand(...) < edge_prob``
    (this DRAWS ``N(N-1)/2`` numbers from the global torch RNG exactly as PyG does),
    then symmetrised and coalesced.  With ``edge_prob=1`` the result is the complete
    directed graph sorted by (row, col) without self loops.
    """
    if directed:
        idx = torch.arange((num_nodes - 1) * num_nodes)
        idx = idx.view(num_nodes - 1, num_nodes)
        idx = idx + torch.arange(1, num_nodes).view(-1, 1)
        idx = idx.view(-1)
    else:
        idx = torch.combinations(torch.arange(num_nodes), r=2)
    mask = torch.rand(idx.size(0)) < edge_prob
    idx = idx[mask]
    if directed:
        row = idx.div(num_nodes - 1, rounding_mode="floor")
        col = idx % num_nodes
        return torch.stack([row, col], dim=0)
    ei = idx.t()
    row = torch.cat([ei[0], ei[1]])
    col = torch.cat([ei[1], ei[0]])
    return _coalesce(torch.stack([row, col], dim=0), num_nodes)


def complete_graph(num_nodes: int) -> torch.Tensor:
    """Closed form of ``erdos_renyi_graph(N, 1)``: all (i, j), i != j, sorted by (row, col)."""
    r = torch.arange(num_nodes).repeat_interleave(num_nodes)
    c = torch.arange(num_nodes).repeat(num_nodes)
    m = r != c
    return torch.stack([r[m], c[m]], dim=0)


def grid(height: int, width: int) -> torch.Tensor:
    """PyG-2.3.1 ``torch_geometric.utils.grid`` edge index (8-neighbour mesh + self loops).

    Restates PyG's construction literally (9-offset kernel, first/last three kernel
    entries of each grid row dropped, out-of-range dropped, coalesced) so that the edge
    ORDER, sorted by (row, col), is the reference's.  Node id = r * width + c.
    """
    w = width
    kernel = torch.tensor([-w - 1, -1, w - 1, -w, 0, w, -w + 1, 1, w + 1])
    row = torch.arange(height * width, dtype=torch.long)
    row = row.view(-1, 1).repeat(1, kernel.size(0))
    col = row + kernel.view(1, -1)
    row, col = row.view(height, -1), col.view(height, -1)
    index = torch.arange(3, row.size(1) - 3, dtype=torch.long)
    row, col = row[:, index].reshape(-1), col[:, index].reshape(-1)
    mask = (col >= 0) & (col < height * width)
    row, col = row[mask], col[mask]
    return _coalesce(torch.stack([row, col], dim=0), height * width)


# --------------------------------------------------------------------------------------
# gcn_norm
# --------------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """A.2: drop existing self loops (keep order), append (i, i) for i in 0..N-1."""
    mask = edge_index[0] != edge_index[1]
    loop = torch.arange(num_nodes, dtype=edge_index.dtype)
    loop = loop.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index[:, mask], loop], dim=1)


def exact_dis(deg: torch.Tensor) -> torch.Tensor:
    """deg^-1/2 computed in fp64 and rounded once to fp32 (0 where deg == 0)."""
    # fp64 1/sqrt is within 1 ulp(fp64) of the true value; for integer degrees that can
    # only move the fp32 rounding at an exact tie, which 1/sqrt(int) never hits.
    d = deg.double()
    return torch.where(d > 0, 1.0 / d.sqrt(), torch.zeros_like(d)).float()


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32,
             dis_mode: str = "torch", improved: bool = False
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """A.3: returns (edge_index', edge_weight, dis).

    ``dis_mode="torch"`` is the literal PyG sequence ``deg.pow(-0.5)`` in ``dtype``;
    ``"exact"`` rounds the fp64 value once to fp32 (what the CUDA preprocessor emits;
    the two differ by at most 1 ulp, asserted in tests/test_oracle.py).
    ``improved=True`` (``GCNConv(improved=True)``, not used by GWEN): the appended self loops
    weigh ``fill_value = 2``, except at nodes whose self loop was already in the input --
    PyG's ``add_remaining_self_loops`` keeps an existing loop's weight
    (``loop_attr[edge_index[0][inv_mask]] = edge_attr[inv_mask]``), which is 1 here.
    """
    ei = add_remaining_self_loops(edge_index, num_nodes)
    row, col = ei[0], ei[1]
    ew = torch.ones(ei.size(1), dtype=dtype)
    if improved:
        loop_w = torch.full((num_nodes,), 2.0, dtype=dtype)
        inv = edge_index[0] == edge_index[1]
        loop_w[edge_index[0][inv]] = 1.0
        ew[ei.size(1) - num_nodes:] = loop_w
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, col, ew)
    if dis_mode == "torch":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
    elif dis_mode == "exact":
        dis = exact_dis(deg).to(dtype)
    else:
        raise ValueError(dis_mode)
    ew = dis[row] * ew * dis[col]
    return ei, ew, dis


# --------------------------------------------------------------------------------------
# GCNConv forward (A.5, A.6)
# --------------------------------------------------------------------------------------
def propagate(x: torch.Tensor, ei: torch.Tensor, ew: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """gather(row) -> scale -> scatter_add_(col) on dim -2 (PyG ``propagate`` with aggr='add')."""
    row, col = ei[0], ei[1]
    x_j = x.index_select(-2, row)
    msg = ew.view(-1, 1) * x_j
    size = list(x.shape)
    size[-2] = num_nodes
    idx = col.view(*([1] * (x.dim() - 2)), -1, 1).expand_as(msg)
    return x.new_zeros(size).scatter_add_(-2, idx, msg)


def gcn_conv_forward(x: torch.Tensor, edge_index: torch.Tensor, weight: torch.Tensor,
                     bias: Optional[torch.Tensor], dis_mode: str = "torch") -> torch.Tensor:
    """One ``GCNConv.forward`` with ``cached=False`` (norm recomputed, as GWEN uses it)."""
    n = x.size(-2)
    ei, ew, _ = gcn_norm(edge_index, n, x.dtype, dis_mode)
    xp = F.linear(x, weight)
    out = propagate(xp, ei, ew, n)
    if bias is not None:
        out = out + bias
    return out


class _LinOracle(torch.nn.Module):
    """PyG ``Linear(in, out, bias=False)``: an uninitialised [out, in] weight (no RNG draw at
    construction, unlike torch.nn.Linear) applied with F.linear."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))

    def forward(self, x):
        return F.linear(x, self.weight)


class GCNConvOracle(torch.nn.Module):
    """``GCNConv(in, out)`` with PyG's parameter names (``lin.weight``, ``bias``) and init (A.1)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _LinOracle(in_channels, out_channels)
        self.bias = torch.nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x, edge_index):
        return gcn_conv_forward(x, edge_index, self.lin.weight, self.bias)


class _Down(torch.nn.Module):
    def __init__(self, c_in, h):  # reference models_gnn.py:109-133
        super().__init__()
        self.conv1 = GCNConvOracle(c_in, h)
        self.conv2 = GCNConvOracle(h, h // 2)
        self.conv3 = GCNConvOracle(h // 2, h // 4)
        self.conv4 = GCNConvOracle(h // 4, h // 8)
        self.conv5 = GCNConvOracle(h // 8, h // 16)

    def forward(self, x, ei):  # reference models_gnn.py:147-149
        x = torch.relu(self.conv1(x, ei))
        x = torch.relu(self.conv2(x, ei))
        return torch.relu(self.conv3(x, ei))


class _Up(torch.nn.Module):
    def __init__(self, h, c_out):  # reference models_gnn.py:163-187
        super().__init__()
        self.upconv1 = GCNConvOracle(h // 16, h // 8)
        self.upconv2 = GCNConvOracle(h // 8, h // 4)
        self.upconv3 = GCNConvOracle(h // 4, h // 2)
        self.upconv4 = GCNConvOracle(h // 2, h)
        self.upconv5 = GCNConvOracle(h, c_out)

    def forward(self, x, ei):  # reference models_gnn.py:204-206
        x = torch.relu(self.upconv3(x, ei))
        x = torch.relu(self.upconv4(x, ei))
        return self.upconv5(x, ei)


class _Layers(torch.nn.Module):
    def __init__(self, c_in, c_out, h):  # reference models_gnn.py:227-239
        super().__init__()
        self.down_conv_layers = _Down(c_in, h)
        self.up_conv_layers = _Up(h, c_out)

    def forward(self, x, ei):
        return self.up_conv_layers(self.down_conv_layers(x, ei), ei)


class GNNModelOracle(torch.nn.Module):
    """Same module tree (hence state_dict keys) as reference ``GNNModel`` (models_gnn.py:268-303)."""

    def __init__(self, channels_in: int, channels_out: int, hidden_feats: int):
        super().__init__()
        self.conv_layers = _Layers(channels_in, channels_out, hidden_feats)
        self.activation = torch.nn.ReLU()  # unused, as in the reference (:290)

    def forward(self, x, edge_index):
        return self.conv_layers(x, edge_index)


def loss_func(output, target, target_mask):
    """reference ``src/gwen/models_gnn.py:261-265``: L1 over the masked node rows."""
    return torch.nn.L1Loss()(output[target_mask], target[target_mask])


# --------------------------------------------------------------------------------------
# Independent dense fp64 restatement (from the GCN formula, not the PyG op sequence)
# --------------------------------------------------------------------------------------
def dense_norm_adj(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """A_hat = D^-1/2 (A + I) D^-1/2 in fp64; A[dst, src] counts duplicate edges, existing
    self loops are replaced by exactly one, D = row sums of (A + I) (in-degree)."""
    a = torch.zeros(num_nodes, num_nodes, dtype=torch.float64)
    src, dst = edge_index[0].tolist(), edge_index[1].tolist()
    for s, d in zip(src, dst):
        if s != d:
            a[d, s] += 1.0
    a += torch.eye(num_nodes, dtype=torch.float64)
    deg = a.sum(dim=1)
    dis = deg.pow(-0.5)
    return dis.view(-1, 1) * a * dis.view(1, -1)


def dense_gcn_forward(x, edge_index, weight, bias):
    a = dense_norm_adj(edge_index, x.size(-2))
    out = a @ (x.double() @ weight.double().t())
    return out + bias.double() if bias is not None else out


# --------------------------------------------------------------------------------------
# dst-sorted CSR (what the CUDA graph preprocessor must reproduce bit-exactly; App. B.3)
# --------------------------------------------------------------------------------------
def dst_sorted_csr(edge_index: torch.Tensor, num_nodes: int):
    """Stable sort of edge_index' by destination.

    Returns numpy (rowptr int32[N+1], src int32[E'], perm int64[E'], dis fp32[N] exact).
    ``perm[slot]`` is the position of that message in ``edge_index'`` (after A.2).
    """
    ei = add_remaining_self_loops(edge_index, num_nodes).numpy()
    row, col = ei[0], ei[1]
    perm = np.argsort(col, kind="stable").astype(np.int64)
    counts = np.bincount(col, minlength=num_nodes)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    dis = (1.0 / np.sqrt(counts.astype(np.float64))).astype(np.float32)
    return rowptr.astype(np.int32), row[perm].astype(np.int32), perm, dis
