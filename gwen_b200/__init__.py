"""gwen_b200 -- B200-native (sm_100a) implementation of GWEN's GCN message-passing hot path.

``from gwen_b200 import GCNConv`` replaces ``from torch_geometric.nn import GCNConv`` at
reference ``src/gwen/models_gnn.py:19``; everything numeric runs in ``libgwen_b200.so``
(C ABI in ``include/gwen_b200.h``).  See DESIGN.md and INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from .graph import (GraphCSR, build_graph, clear_graph_cache, complete_graph, erdos_renyi_graph,
                    get_graph, grid, grid_edge_count)
from .nn import GCNConv, gcn_conv
from .models_gnn import (DownConvLayers, GCNConvLayers, GNNConfig, GNNModel, UpConvLayers,
                         loss_func)
from . import ops  # noqa: F401
from . import partition  # noqa: F401
from .host_stream import HostBandPropagator, HostPropagator, pinned_near_gpu
from .train import eval_step, gather_eval_results, masked_l1_loss, train_step
from .data import GraphDataset
from .loader import NeighborLoader
from . import optim  # noqa: F401

__version__ = "0.1.0"
__all__ = ["GCNConv", "gcn_conv", "GraphCSR", "build_graph", "get_graph", "clear_graph_cache",
           "grid", "grid_edge_count", "complete_graph", "erdos_renyi_graph", "HostPropagator", "HostBandPropagator", "pinned_near_gpu", "masked_l1_loss", "train_step", "eval_step", "gather_eval_results", "GraphDataset", "NeighborLoader", "GNNConfig",
           "DownConvLayers", "UpConvLayers", "GCNConvLayers", "GNNModel", "loss_func", "ops", "partition", "optim"]
