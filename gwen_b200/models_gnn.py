"""The GCN layer stack of GWEN on top of :class:`gwen_b200.nn.GCNConv`.

Mirror of the caller side of the boundary in the reference (``src/gwen/models_gnn.py:86-303``):
same class names, constructor arguments, attribute names (hence ``state_dict`` keys
``conv_layers.down_conv_layers.conv{1..5}.{bias,lin.weight}`` /
``conv_layers.up_conv_layers.upconv{1..5}.{bias,lin.weight}``) and layer wiring: six live layers
``C -> h -> h/2 -> h/4 -> h/2 -> h -> C`` with ReLU after the first five; ``conv4/5`` and
``upconv1/2`` own parameters but never run (commented out at reference ``:150-151, :202-203``).
The only differences from the reference file are the import of ``GCNConv`` and that the ReLU is
requested from the layer's fused epilogue (``relu=True``) instead of a separate ``torch.relu``.
Training/eval drivers, MLflow and data loading are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from .graph import GraphCSR, get_graph
from . import ops
from .nn import GCNConv, ReluLink, b2b_fusable, gcn_conv_b2b_project, gcn_conv_pair, pair_fusable

__all__ = ["GNNConfig", "DownConvLayers", "UpConvLayers", "GCNConvLayers", "GNNModel", "loss_func"]


@dataclass
class GNNConfig(dict):
    """reference ``models_gnn.py:86-103``; only the last three fields are used by the layers."""

    nodes_in: int
    nodes_out: int
    channels_in: int
    channels_out: int
    hidden_feats: int


class DownConvLayers(nn.Module):
    def __init__(self, gnn_configs: GNNConfig):
        super().__init__()
        h = gnn_configs.hidden_feats
        self.conv1 = GCNConv(gnn_configs.channels_in, h)
        self.conv2 = GCNConv(h, h // 2)
        self.conv3 = GCNConv(h // 2, h // 4)
        self.conv4 = GCNConv(h // 4, h // 8)
        self.conv5 = GCNConv(h // 8, h // 16)

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        # training: the ReLU backward of conv1 and conv2 rides in the dgrad epilogue of the layer after it
        # (nn.ReluLink; bf16 tcgen05 path only, same bits as the unfused backward)
        l1, l2 = ReluLink(), ReluLink()
        if x.is_cuda and x.dtype == torch.bfloat16:
            # inference on a large mesh (both bitwise equal to the layers run one by one):
            #  * conv1 and conv2's projection back to back in one kernel (gwen_b200.nn.gcn_conv_b2b_project): conv1's
            #    1024-wide output never reaches HBM;
            #  * conv2's aggregation + bias + ReLU inside conv3's fused kernel (gwen_b200.nn.gcn_conv_pair)
            graph = edge_index if isinstance(edge_index, GraphCSR) else get_graph(edge_index, x.size(-2))
            p2 = None
            if b2b_fusable(x, self.conv1, self.conv2, "down"):
                p2 = gcn_conv_b2b_project(x, graph, self.conv1, self.conv2)
                x2_like = p2
            else:
                x = self.conv1(x, edge_index, relu=True, link_out=l1)
                x2_like = x
            if pair_fusable(graph, x2_like, self.conv2, self.conv3):
                return gcn_conv_pair(x if p2 is None else None, graph, self.conv2, self.conv3, relu_b=True, p=p2)
            if p2 is not None:
                x = ops.aggregate(graph, p2, self.conv2.bias, True)
                return self.conv3(x, edge_index, relu=True)
        else:
            x = self.conv1(x, edge_index, relu=True, link_out=l1)
        x = self.conv2(x, edge_index, relu=True, link_in=l1, link_out=l2)
        x = self.conv3(x, edge_index, relu=True, link_in=l2)
        return x


class UpConvLayers(nn.Module):
    def __init__(self, gnn_configs: GNNConfig):
        super().__init__()
        h = gnn_configs.hidden_feats
        self.upconv1 = GCNConv(h // 16, h // 8)
        self.upconv2 = GCNConv(h // 8, h // 4)
        self.upconv3 = GCNConv(h // 4, h // 2)
        self.upconv4 = GCNConv(h // 2, h)
        self.upconv5 = GCNConv(h, gnn_configs.channels_out)

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        x = self.upconv3(x, edge_index, relu=True)
        if x.is_cuda and x.dtype == torch.bfloat16 and b2b_fusable(x, self.upconv4, self.upconv5, "up"):
            # inference on a large mesh: upconv4 and upconv5's projection back to back, upconv4's 1024-wide output
            # never reaches HBM (bitwise equal to the two layers)
            graph = edge_index if isinstance(edge_index, GraphCSR) else get_graph(edge_index, x.size(-2))
            p5 = gcn_conv_b2b_project(x, graph, self.upconv4, self.upconv5)
            return ops.aggregate(graph, p5, self.upconv5.bias, False)
        l4 = ReluLink()
        x = self.upconv4(x, edge_index, relu=True, link_out=l4)
        x = self.upconv5(x, edge_index, link_in=l4)
        return x


class GCNConvLayers(nn.Module):
    def __init__(self, gnn_configs: GNNConfig):
        super().__init__()
        self.down_conv_layers = DownConvLayers(gnn_configs)
        self.up_conv_layers = UpConvLayers(gnn_configs)

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        x = self.down_conv_layers(x, edge_index)
        x = self.up_conv_layers(x, edge_index)
        return x


def loss_func(output, target, target_mask):
    """reference ``models_gnn.py:261-265`` (caller-side; plain torch, not on the kernel path)."""
    return nn.L1Loss()(output[target_mask], target[target_mask])


class GNNModel(nn.Module):
    """reference ``models_gnn.py:268-303`` without the train/eval drivers."""

    def __init__(self, gnn_configs: GNNConfig) -> None:
        super().__init__()
        self.conv_layers = GCNConvLayers(gnn_configs)
        self.activation = nn.ReLU()  # unused, as in the reference

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        return self.conv_layers(x, edge_index)
