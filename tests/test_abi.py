"""CPU tests of the C-ABI boundary and host logic: the library loads, exports every symbol the
header declares, validates arguments before touching CUDA, and the Python layer refuses to run
without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import gwen_b200 as gw
from gwen_b200 import _lib
from oracle import gcn_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gwen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gwen_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), "libgwen_b200.so lacks %s" % s
    assert set(syms) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert _lib.lib().gwen_version() == 1


def test_error_convention_without_device():
    L = _lib.lib()
    need = C.c_size_t()
    assert L.gwen_graph_workspace_bytes(-1, 0, 1, C.byref(need)) == -1           # GWEN_E_BADARG
    assert b"negative" in L.gwen_last_error()
    assert L.gwen_graph_workspace_bytes(2 ** 31, 0, 1, C.byref(need)) == -1
    assert L.gwen_graph_workspace_bytes(4, 4, 1, None) == -1
    assert L.gwen_aggregate_fwd(None, None, None, None, None, None, 1, 4, 4, 8, 8, 32, 8, 32, 0, None, 0, None) == -1
    assert L.gwen_aggregate_fwd(None, None, None, None, None, None, 0, 4, 4, 8, 8, 32, 8, 32, 0, None, 0, None) == 0
    assert L.gwen_linear_fwd(None, None, None, 4, 4, 4, 4, 4, 4, 7, None, 0, None) == -3   # GWEN_E_DTYPE
    assert L.gwen_linear_fwd(None, None, None, 4, 4, 4, 4, 4, 4, 0, None, 0, None) == -1
    assert L.gwen_grid_edge_count(3, 4) == 70
    assert L.gwen_grid_edge_count(582, 390) == 2036992
    assert L.gwen_grid_edge_count(1158, 774) == 8055040
    assert L.gwen_grid_edge_count(2048, 2048) == 37724164
    assert L.gwen_grid_edge_count(1, 1) == 1 and L.gwen_grid_edge_count(0, 5) == 0


def test_grid_edge_count_matches_oracle():
    for h, w in ((1, 1), (1, 6), (6, 1), (2, 2), (3, 4), (9, 13)):
        assert gw.grid_edge_count(h, w) == orc.grid(h, w).size(1)


def test_no_cpu_fallback():
    conv = gw.GCNConv(4, 8)
    x = torch.randn(5, 4)
    ei = orc.complete_graph(5)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(x, ei)
    with pytest.raises(RuntimeError, match="CUDA"):
        gw.build_graph(ei, 5)
    with pytest.raises(RuntimeError):
        gw.grid(3, 4, device="cpu")


def test_module_api_and_state_dict_compat():
    torch.manual_seed(0)
    conv = gw.GCNConv(6, 10)
    assert conv.in_channels == 6 and conv.out_channels == 10
    assert set(conv.state_dict()) == {"bias", "lin.weight"}
    assert conv.lin.weight.shape == (10, 6) and conv.bias.shape == (10,)
    a = (6.0 / 16) ** 0.5
    assert conv.lin.weight.abs().max() <= a and torch.all(conv.bias == 0)
    assert gw.GCNConv(6, 10, bias=False).bias is None
    with pytest.raises(TypeError):
        gw.GCNConv(6, 10, foo=1)
    # same module tree / keys as the reference model (and the oracle restatement of it)
    cfg = gw.GNNConfig(nodes_in=2, nodes_out=2, channels_in=12, channels_out=12, hidden_feats=64)
    model = gw.GNNModel(cfg)
    ref = orc.GNNModelOracle(12, 12, 64)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(model.state_dict())
    model.load_state_dict(ref.state_dict())
    # same init stream as PyG's glorot (uniform_ on [out, in], layers in construction order)
    torch.manual_seed(5)
    m1 = gw.GCNConv(7, 3)
    torch.manual_seed(5)
    m2 = orc.GCNConvOracle(7, 3)
    assert torch.equal(m1.lin.weight, m2.lin.weight)


def test_erdos_renyi_rng_side_effect_without_gpu():
    torch.manual_seed(3)
    with pytest.raises(RuntimeError):
        gw.erdos_renyi_graph(6, 1, device="cpu")
    with pytest.raises(NotImplementedError):
        gw.erdos_renyi_graph(6, 0.5)
