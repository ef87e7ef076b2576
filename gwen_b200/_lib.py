"""ctypes binding of ``libgwen_b200.so`` (C ABI declared in ``include/gwen_b200.h``).

The shared library is the product; this module only loads it, declares the prototypes and
turns negative return codes into ``RuntimeError``.  There is no CPU fallback: if the library
is missing the import of :mod:`gwen_b200` fails with instructions to build it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libgwen_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["common.cu", "graph_build.cu", "aggregate.cu", "stencil.cu", "linear.cu", "linear_tc.cu", "linear_tc3.cu", "linear_wgrad_tc.cu", "loss.cu", "gcn_fused.cu", "linear_b2b.cu", "linear_tf32x3.cu", "linear_wgrad_tf32x3.cu", "neighbor.cu", "mesh_mask.cu", "optim.cu", "locality.cu"]
HEADERS = ["common.cuh", "tma.cuh", "tcgen05.cuh", "stencil_common.cuh"]

GWEN_F32, GWEN_BF16 = 0, 1
GRAPH_ADD_SELF_LOOPS, GRAPH_IMPROVED, GRAPH_TRANSPOSE = 1, 2, 4
EPI_NONE, EPI_RELU = 0, 1
GWEN_E_NOSUPPORT = -5
PLAN_GATHER = 1


def nvcc_command(out_path: str = LIB_PATH) -> list:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "--expt-relaxed-constexpr", "--threads", "0", "-Xcompiler", "-fPIC", "-shared",
            # the CUDA runtime is the process's shared one (torch has loaded libcudart.so.12 by the time this
            # library is opened): no second static copy of the runtime inside the product binary
            "--cudart", os.environ.get("GWEN_CUDART", "shared"),
            "-I" + os.path.join(_ROOT, "include"), "-o", out_path] + \
        [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(_ROOT, "include", "gwen_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``gwen_b200/libgwen_b200.so`` (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = nvcc_command(tmp)
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class TilePlanStruct(C.Structure):
    _fields_ = [("num_tiles", C.c_int32), ("run_len", C.c_int32), ("max_tile_runs", C.c_int32),
                ("max_tile_rows", C.c_int32), ("max_tile_msgs", C.c_int32), ("reserved", C.c_int32),
                ("n_dst", C.c_int64), ("tile_ptr", C.c_void_p), ("run_ptr", C.c_void_p),
                ("run_start", C.c_void_p), ("trec", C.c_void_p), ("tmsg", C.c_void_p),
                ("tmsg_base", C.c_void_p)]


class HaloPeersStruct(C.Structure):
    """gwen_halo_peers (include/gwen_b200.h)."""
    _fields_ = [("up_row", C.c_void_p), ("down_row", C.c_void_p), ("up_bstride", C.c_int64),
                ("down_bstride", C.c_int64), ("up_flag", C.c_void_p), ("down_flag", C.c_void_p),
                ("ctl", C.c_void_p)]


_p, _i64, _i32, _u32, _int, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_int, C.c_size_t

# name -> (restype, argtypes); every symbol include/gwen_b200.h declares.
PROTOTYPES = {
    "gwen_version": (_int, []),
    "gwen_set_sm_reserve": (_int, [_int]),
    "gwen_last_error": (C.c_char_p, []),
    "gwen_graph_workspace_bytes": (_int, [_i64, _i64, _u32, C.POINTER(_sz)]),
    "gwen_graph_build": (_int, [_p, _i64, _i64, _u32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gwen_grid_edge_count": (_i64, [_i64, _i64]),
    "gwen_grid_edges": (_int, [_i64, _i64, _p, _p]),
    "gwen_complete_edges": (_int, [_i64, _p, _p]),
    "gwen_aggregate_fwd": (_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64,
                                  _i64, _i64, _int, _p, _int, _p]),
    "gwen_tile_plan_workspace_bytes": (_int, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "gwen_tile_plan_build": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _p, _p, _p, _p,
                                    _p, _p, _p, _sz, _p]),
    "gwen_locality_workspace_bytes": (_int, [_i64, C.POINTER(_sz)]),
    "gwen_locality_tiles": (_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gwen_uniform_tiles": (_int, [_i64, _i32, _p, _p]),
    "gwen_grid_tiles": (_int, [_i64, _i64, _i32, _i32, _p, _p, _p]),
    "gwen_aggregate_tiled_fwd": (_int, [C.POINTER(TilePlanStruct), _p, _p, _i64, _i64, _i64,
                                        _i64, _i64, _i64, _i64, _int, _p, _int, _i32, _i32, _i32,
                                        _p]),
    "gwen_grid_stencil_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                     _i64, _i64, _int, _p, _int, _i32, _i32, _p]),
    "gwen_grid_stencil_peer_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                          _int, _p, _int, _i32, _i32, C.POINTER(HaloPeersStruct), _p]),
    "gwen_gcn_fused_fwd": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _int, _p, _int, _p]),
    "gwen_linear_b2b_supported": (_int, [_i64, _i64, _i64, _i64, _int]),
    "gwen_linear_b2b_fwd": (_int, [_p, _p, _p, _int, _p, _p, _int, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_linear_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _int, _p]),
    "gwen_linear_fwd_workspace_bytes": (_int, [_i64, _i64, _i64, _int, C.POINTER(_sz)]),
    "gwen_linear_fwd_ws": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _int, _p, _sz, _p]),
    "gwen_linear_bwd_data": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_linear_bwd_data_masked": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_linear_batched_bwd_data_masked": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                                   _i64, _i64, _int, _p]),
    "gwen_linear_batched_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int,
                                       _p, _int, _p]),
    "gwen_linear_batched_bwd_data": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                            _int, _p]),
    "gwen_linear_bwd_data_workspace_bytes": (_int, [_i64, _i64, _i64, _int, C.POINTER(_sz)]),
    "gwen_linear_bwd_data_ws": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _sz, _p]),
    "gwen_linear_bwd_weight": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p,
                                      _sz, _p]),
    "gwen_linear_bwd_weight_bias_workspace_bytes": (_int, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "gwen_linear_bwd_weight_bias": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _sz, _p]),
    "gwen_linear_bwd_weight_workspace_bytes": (_int, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "gwen_relu_bwd": (_int, [_p, _p, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_bias_grad": (_int, [_p, _p, _i64, _i64, _i64, _int, _p, _sz, _p]),
    "gwen_bias_grad_workspace_bytes": (_int, [_i64, _i64, C.POINTER(_sz)]),
    "gwen_relu_bias_bwd": (_int, [_p, _p, _p, _p, _i64, _i64, _int, _p, _sz, _p]),
    "gwen_masked_l1_workspace_bytes": (_int, [_i64, C.POINTER(_sz)]),
    "gwen_masked_l1_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _p, _sz, _p]),
    "gwen_masked_l1_bwd": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _int, _p, _p]),
    "gwen_rows_gather": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_rows_scatter": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _p]),
    "gwen_neighbor_workspace_bytes": (_int, [_i64, _i64, C.POINTER(_sz)]),
    "gwen_neighbor_sample_full": (_int, [_p, _p, _p, _i64, _i64, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gwen_neighbor_complete": (_int, [_i64, _i64, _i64, _p, _p, _p, _p]),
    "gwen_gather_rows_bytes": (_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "gwen_mesh_mask_detect": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "gwen_rows_self_fwd": (_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int, _p, _int, _p]),
    "gwen_adam_step": (_int, [_i32, _p, _p, _p, _p, _p, _i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _p]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "gwen_b200: %s is missing -- build it with `python -c \"import __graft_entry__ as g; "
                "g.build()\"` (nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().gwen_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))
