"""``GCNConv``: drop-in for ``torch_geometric.nn.GCNConv`` as GWEN uses it.

Reference boundary: imported at ``src/gwen/models_gnn.py:19``, constructed as
``GCNConv(in_channels, out_channels)`` at ``:118-130`` / ``:172-184`` and called as
``conv(x, edge_index)`` at ``:147-149`` / ``:204-206``.  Same constructor arguments, same
parameter names (``lin.weight [out, in]``, ``bias [out]``) and glorot/zeros initialisation
(SURVEY.md Appendix A.1), so ``state_dict``s are interchangeable with the PyG module.

The arithmetic is ``D^-1/2 (A + I) D^-1/2 (x W^T) + b`` computed by libgwen_b200's kernels:
graph preprocessing once per ``edge_index`` (K0), then per call one projection (K2) and one
deterministic segment-reduce aggregation (K1) -- aggregating on whichever side of the
projection is narrower -- with bias and an optional fused ReLU in the last kernel's epilogue.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch import Tensor

from . import ops
from .graph import GraphCSR, get_graph

__all__ = ["GCNConv", "ReluLink", "gcn_conv", "gcn_conv_pair", "pair_fusable", "b2b_fusable", "gcn_conv_b2b_project"]


class ReluLink:
    """Hand-over between two consecutive layers of a CHAIN  y = conv_b(relu(conv_a(x)))  in training: when conv_b's
    backward produces dx with its dgrad GEMM it folds conv_a's ReLU backward into that kernel's epilogue
    (``ops.linear_bwd_data_masked``, mask = conv_b's input = conv_a's ReLU output) and sets ``masked``; conv_a's
    backward then skips its own mask pass (read dy, read y, write dz -> only the bias-gradient sums of dy remain).
    Only for chains: conv_a's output must have no other consumer (the model code wires it, reference
    models_gnn.py:147-149,204-206 is such a chain).  Same bits as the unfused backward."""
    __slots__ = ("masked",)

    def __init__(self):
        self.masked = False


BWD_MASK_FUSION = os.environ.get("GWEN_BWD_MASK_FUSION", "1") != "0"   # fold ReLU masks into the next layer's dgrad
FUSED_IN_TRAINING = os.environ.get("GWEN_FUSED_TRAIN", "0") == "1"      # fused layer kernel also when gradients are wanted


class _GCNConvFn(torch.autograd.Function):
    """y = epi(A_hat (x W^T) + b); backward per SURVEY.md Appendix A.7."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], graph: GraphCSR,
                relu: bool, agg_first: bool, link_in: Optional[ReluLink] = None,
                link_out: Optional[ReluLink] = None):
        # the fused kernel keeps A_hat x on chip -- which the backward needs (dW = dy^T (A_hat x)): in training the
        # two-kernel form, whose aggregated tensor is saved, beats fused + recomputing the stencil in backward
        # (measured per member step at the cfg 5 shape); GWEN_FUSED_TRAIN=1 restores the fused forward
        training = any(ctx.needs_input_grad[:3])
        fused = agg_first and (FUSED_IN_TRAINING or not training) and ops.gcn_fused_preferred(graph, x, weight)
        if fused:          # (A_hat x) W^T in ONE kernel: the aggregated rows never leave the SM
            y = ops.gcn_fused(graph, x, weight, bias, relu)
            saved_in = x   # backward recomputes A_hat x (one narrow stencil) instead of storing it
        elif agg_first:    # (A_hat x) W^T: aggregate at the narrower input width
            h = ops.aggregate(graph, x)
            y = ops.linear(h, weight, bias, relu)
            saved_in = h
        else:              # A_hat (x W^T): aggregate at the narrower output width
            h = ops.linear(x, weight)
            y = ops.aggregate(graph, h, bias, relu)
            saved_in = x
        ctx.graph, ctx.relu, ctx.agg_first, ctx.fused = graph, relu, agg_first, fused
        ctx.link_in, ctx.link_out = link_in, (link_out if relu else None)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(saved_in, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        saved_in, weight, y = ctx.saved_tensors
        graph_t = ctx.graph.transposed()
        # ReLU mask and bias gradient in one pass over dy (out of place: autograd owns dy); when the next layer's
        # dgrad epilogue has already applied this layer's mask (ReluLink) only the bias-gradient sums remain
        masked = ctx.link_out is not None and ctx.link_out.masked
        if masked:
            ctx.link_out.masked = False
        # no mask left to apply and the gradient goes straight into the wgrad: the bias gradient comes out of the
        # wgrad's own pass over dy (ones-column MMA), no separate pass at all
        db_from_wgrad = ctx.agg_first and ctx.has_bias and (masked or not ctx.relu) and BWD_MASK_FUSION \
            and dy.dtype == torch.bfloat16
        db = None
        if not db_from_wgrad:
            dy, db = ops.relu_bias_bwd(dy.contiguous(), y if (ctx.relu and not masked) else None, ctx.has_bias)
        need_dx = ctx.needs_input_grad[0]
        dx = None
        if ctx.agg_first:
            if ctx.fused:
                saved_in = ops.aggregate(ctx.graph, saved_in)     # recompute A_hat x
            dw = None
            if db_from_wgrad:
                both = ops.linear_bwd_weight_bias(dy, saved_in)
                if both is None:                                  # outside the tensor-core wgrad: the two passes
                    dy, db = ops.relu_bias_bwd(dy.contiguous(), None, True)
                else:
                    dw, db = both
            if dw is None:
                dw = ops.linear_bwd_weight(dy, saved_in)          # dW = dy^T (A_hat x)
            if need_dx:
                dx = ops.aggregate(graph_t, ops.linear_bwd_data(dy, weight))
        else:
            dh = ops.aggregate(graph_t, dy)                       # A_hat^T dy
            dw = ops.linear_bwd_weight(dh, saved_in)              # dW = dh^T x
            if need_dx:
                if ctx.link_in is not None and BWD_MASK_FUSION:   # x = relu(previous layer): its mask in our epilogue
                    dx = ops.linear_bwd_data_masked(dh, weight, saved_in)
                    ctx.link_in.masked = dx is not None
                if dx is None:
                    dx = ops.linear_bwd_data(dh, weight)
        if db is not None:
            db = db.to(weight.dtype)
        return dx, dw.to(weight.dtype), db, None, None, None, None, None


def gcn_conv(x: Tensor, graph: GraphCSR, weight: Tensor, bias: Optional[Tensor] = None,
             relu: bool = False, agg_first: Optional[bool] = None, link_in: Optional[ReluLink] = None,
             link_out: Optional[ReluLink] = None) -> Tensor:
    """Functional form on a prebuilt graph handle.  ``link_in`` / ``link_out``: see :class:`ReluLink`."""
    if agg_first is None:
        agg_first = weight.shape[1] < weight.shape[0]
    return _GCNConvFn.apply(x, weight, bias, graph, relu, agg_first, link_in, link_out)


# Cross-layer fusion (inference): two consecutive layers  y = epi_b(conv_b(relu(conv_a(x))))  where conv_a
# aggregates LAST (out <= in):  conv_a(x) = A_hat (x Wa^T) + ba.  Its aggregation, bias and ReLU move into the
# stencil warps of conv_b's fused kernel, so conv_a is only a GEMM and its aggregated output (the widest tensor
# of the pair) never exists in HBM:
#     p  = x Wa^T                                           (tcgen05 GEMM, no epilogue)
#     q  = relu(A_hat p + ba) Wb^T                          (gwen_gcn_fused_fwd with pre_bias / pre_relu)
#     y  = epi_b(A_hat q + bb)   if conv_b aggregates last  (mesh stencil)   -- GWEN: conv2 -> conv3
#     y  = epi_b(q + bb) comes out of the fused kernel's epilogue otherwise.
# The A operand is rounded to bf16 exactly where the layer-by-layer path stores conv_a's output, so the result
# is bitwise equal to it.
PAIR_FUSION = os.environ.get("GWEN_PAIR_FUSION", "1") != "0"


def pair_fusable(graph: GraphCSR, x: Tensor, conv_a: "GCNConv", conv_b: "GCNConv") -> bool:
    if not PAIR_FUSION or torch.is_grad_enabled() and (x.requires_grad or conv_a.lin.weight.requires_grad):
        return False
    if conv_a.in_channels < conv_a.out_channels or conv_b.in_channels != conv_a.out_channels:
        return False
    if conv_a.bias is None:
        return False
    return ops.gcn_fused_preferred(graph, x[..., :1].expand(x.shape[:-1] + (conv_a.out_channels,)), conv_b.lin.weight)


@torch.no_grad()
def gcn_conv_pair(x: Optional[Tensor], graph: GraphCSR, conv_a: "GCNConv", conv_b: "GCNConv", relu_b: bool,
                  p: Optional[Tensor] = None) -> Tensor:
    """``epi_b(conv_b(relu(conv_a(x))))`` with conv_a's aggregation fused into conv_b's kernel (inference).
    ``p``: conv_a's un-aggregated projection ``x Wa^T`` when the caller already has it (gcn_conv_b2b_project)."""
    if p is None:
        p = ops.linear(x, conv_a.lin.weight)
    b_last = conv_b.in_channels >= conv_b.out_channels
    q = ops.gcn_fused(graph, p, conv_b.lin.weight, None if b_last else conv_b.bias, False if b_last else relu_b,
                      pre_bias=conv_a.bias, pre_relu=True)
    return ops.aggregate(graph, q, conv_b.bias, relu_b) if b_last else q


# Back-to-back projections (inference): conv_a aggregates FIRST (in < out) and conv_b projects FIRST (in > out):
#     conv_b(relu(conv_a(x))) = A_hat ( relu((A_hat x) Wa^T + ba) Wb^T ) + bb
# The hidden tensor relu(...) -- the widest of the stack (GWEN: conv1 -> conv2 and upconv4 -> upconv5, 1024 wide) --
# is produced and consumed inside ONE kernel (gwen_linear_b2b_fwd) and never reaches HBM:
#     a = A_hat x                           (mesh stencil / CSR kernel at conv_a's narrow input width)
#     p = relu(a Wa^T + ba) Wb^T            (two tcgen05 products back to back)
# and p, conv_b's un-aggregated projection, goes on to conv_b's aggregation (or into the pair fusion above).
# The hidden block is rounded to bf16 exactly where the layer-by-layer path stores conv_a's output.
B2B_FUSION = os.environ.get("GWEN_B2B_FUSION", "1") != "0"
B2B_PAIRS = os.environ.get("GWEN_B2B_PAIRS", "down,up").split(",")   # which pairs may fuse: conv1->conv2 (down), upconv4->upconv5 (up)
B2B_MIN_ROWS = 500_000     # rows (members x nodes) from which the one kernel beats the two layers' kernels (measured)


def b2b_fusable(x: Tensor, conv_a: "GCNConv", conv_b: "GCNConv", which: Optional[str] = None) -> bool:
    if which is not None and which not in B2B_PAIRS:
        return False
    if not B2B_FUSION or torch.is_grad_enabled() and (x.requires_grad or conv_a.lin.weight.requires_grad
                                                      or conv_b.lin.weight.requires_grad):
        return False
    if conv_a.in_channels >= conv_a.out_channels or conv_b.in_channels <= conv_b.out_channels or \
            conv_b.in_channels != conv_a.out_channels:
        return False
    if x.numel() // max(1, x.shape[-1]) < B2B_MIN_ROWS:
        return False
    return ops.linear_b2b_supported(x, conv_a.lin.weight, conv_b.lin.weight)


@torch.no_grad()
def gcn_conv_b2b_project(x: Tensor, graph: GraphCSR, conv_a: "GCNConv", conv_b: "GCNConv") -> Tensor:
    """``relu(conv_a(x)) Wb^T``: conv_a with its ReLU, then conv_b's projection, the hidden tensor kept on chip."""
    a = ops.aggregate(graph, x)
    return ops.linear_b2b(a, conv_a.lin.weight, conv_a.bias, True, conv_b.lin.weight)


class _Linear(torch.nn.Module):
    """Weight holder named like PyG's ``Linear(in, out, bias=False)`` (``lin.weight``)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))  # glorot uniform
        with torch.no_grad():
            self.weight.uniform_(-a, a)

    def forward(self, x: Tensor) -> Tensor:
        return ops.linear(x, self.weight)

    def extra_repr(self):
        return "%d, %d, bias=False" % (self.in_channels, self.out_channels)


class GCNConv(torch.nn.Module):
    """``GCNConv(in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
    normalize=True, bias=True)``; ``forward(x [..., N, F_in], edge_index int64 [2, E]) ->
    [..., N, F_out]``.  ``relu=True`` (extension) fuses the ``torch.relu`` the reference model
    applies after the layer into the kernel epilogue.
    """

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False,
                 cached: bool = False, add_self_loops: bool = True, normalize: bool = True,
                 bias: bool = True, **kwargs):
        super().__init__()
        if kwargs:
            raise TypeError("unsupported GCNConv arguments: %s" % sorted(kwargs))
        if not normalize:
            raise NotImplementedError("GCNConv(normalize=False) is not part of the GWEN path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.add_self_loops, self.normalize = improved, cached, add_self_loops, normalize
        self.lin = _Linear(in_channels, out_channels)
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._cached_graph: Optional[GraphCSR] = None
        self.reset_parameters()

    def reset_parameters(self):
        self.lin.reset_parameters()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()
        self._cached_graph = None

    def forward(self, x: Tensor, edge_index, edge_weight: Optional[Tensor] = None,
                relu: bool = False, link_in: Optional[ReluLink] = None,
                link_out: Optional[ReluLink] = None) -> Tensor:
        if edge_weight is not None:
            raise NotImplementedError("edge_weight is not used on the GWEN path")
        if not x.is_cuda:
            raise RuntimeError("gwen_b200.GCNConv runs on CUDA tensors only (no CPU fallback)")
        if isinstance(edge_index, GraphCSR):
            graph = edge_index
        elif self.cached and self._cached_graph is not None:
            graph = self._cached_graph
        else:
            graph = get_graph(edge_index, x.size(-2), self.add_self_loops, self.improved)
            if self.cached:
                self._cached_graph = graph
        return gcn_conv(x, graph, self.lin.weight, self.bias, relu, link_in=link_in, link_out=link_out)

    def propagate(self, edge_index, x: Tensor, add_bias: bool = True, relu: bool = False) -> Tensor:
        """Message + aggregate only (PyG ``MessagePassing.propagate`` with GCN's message): gathers
        ``x`` rows along the edges, scales by the symmetric norm and sums into destinations;
        ``add_bias`` applies the layer bias in the same kernel's epilogue.  Inference-only helper
        (no autograd) used by bench.py for BASELINE config 2."""
        if not x.is_cuda:
            raise RuntimeError("gwen_b200.GCNConv runs on CUDA tensors only (no CPU fallback)")
        graph = edge_index if isinstance(edge_index, GraphCSR) else \
            get_graph(edge_index, x.size(-2), self.add_self_loops, self.improved)
        return ops.aggregate(graph, x.detach(), self.bias if add_bias else None, relu)

    def __repr__(self):
        return "%s(%d, %d)" % (self.__class__.__name__, self.in_channels, self.out_channels)
