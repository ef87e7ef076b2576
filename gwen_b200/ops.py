"""Functional wrappers over the C ABI: aggregate (K1), linear (K2), the fused layer, backward
pieces, row moves.

Every function takes CUDA tensors, launches on the current stream of the tensors' device and
returns a torch-allocated result.  Nothing here computes with torch ops: a missing library or a
CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import check, lib
from .graph import GraphCSR, TilePlan, _ptr, _require_cuda, _stream

__all__ = ["aggregate", "mesh_stencil", "gcn_fused", "gcn_fused_supported", "gcn_fused_preferred",
           "linear", "linear_b2b", "linear_b2b_supported", "linear_bwd_data", "linear_bwd_data_masked", "linear_bwd_weight", "linear_bwd_weight_bias", "relu_bwd_", "bias_grad", "relu_bias_bwd", "copy_rows_",
           "rows_gather", "rows_scatter_", "dtype_code", "clear_cast_cache"]

_DTYPES = {torch.float32: _lib.GWEN_F32, torch.bfloat16: _lib.GWEN_BF16}

# Which K1 kernel `aggregate` uses:
#   "rows"    generic CSR kernel (any graph), fp32 bit-exact vs the CPU scatter_add_ order
#   "tiled"   TMA-staged CSR kernel for mesh graphs, bitwise equal to "rows"
#   "stencil" separable mesh fast path (plain H x W mesh, or a mesh with cut-out nodes: then one extra
#             pass rewrites the cut-out rows), equal to fp32 rounding
#   "auto"    stencil on a plain or masked mesh with 16-byte rows; tiled for any other graph whose nodes
#             are numbered like a grid (2-D tile plans: improved=True meshes, partition-local graphs)
#             when the plan fits in shared memory; else rows
#   "locality" staged kernel over locality tiles (GraphCSR.locality_plan): graphs whose numbering carries no
#             locality; bitwise equal to "rows".  "auto" takes it for sparse graphs of >= LOCALITY_MIN_NODES nodes
#             that are not grid-numbered, and falls back to rows when the graph has no locality
DEFAULT_AGG_KERNEL = "auto"
LOCALITY_TILES = os.environ.get("GWEN_LOCALITY_TILES", "1") != "0"
LOCALITY_MIN_NODES = 32768        # below this the row kernel's working set sits in L2 anyway
LOCALITY_MAX_DEGREE = 32          # mean messages per destination above which a tile's sources cannot fit


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError("gwen_b200 supports float32 and bfloat16, got %s" % dt) from None


def _as_3d(x: torch.Tensor):
    """[..., N, F] -> contiguous [B, N, F] view + the leading shape."""
    if x.dim() < 2:
        raise ValueError("expected [..., N, F]")
    lead = x.shape[:-2]
    x3 = x.reshape((-1,) + tuple(x.shape[-2:])).contiguous()
    return x3, lead


def _bias32(bias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if bias is None:
        return None
    return _cast_cached(bias, torch.float32)


# Parameter casts (fp32 master weights -> bf16 operand, any bias -> fp32) are cached per parameter
# VERSION: the cast kernel runs once after each optimizer update instead of once per layer call
# (forward, dgrad and the recompute paths all ask for the same cast).  The cache holds weak references.
# The key is autograd's version counter: every in-place torch op (torch.optim included) and
# gwen_b200.optim.Adam bump it; code that writes a parameter behind autograd's back (``p.data.copy_``, raw
# pointers) must call ``torch.autograd.graph.increment_version(p)`` or ``ops.clear_cast_cache()``.
_CAST_CACHE: dict = {}


def clear_cast_cache() -> None:
    _CAST_CACHE.clear()


def _cast_cached(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    t = t.detach()
    if t.dtype == dtype and t.is_contiguous():
        return t
    key = (id(t.untyped_storage()), t.storage_offset(), tuple(t.shape), tuple(t.stride()), dtype, t.device)
    ent = _CAST_CACHE.get(key)
    if ent is not None:
        ref, ver, out = ent
        if ref() is t.untyped_storage() and ver == t._version:
            return out
    import weakref
    out = t.to(dtype).contiguous()
    if len(_CAST_CACHE) > 256:
        _CAST_CACHE.clear()
    _CAST_CACHE[key] = (weakref.ref(t.untyped_storage()), t._version, out)
    return out


def aggregate(graph: GraphCSR, x: torch.Tensor, bias: Optional[torch.Tensor] = None,
              relu: bool = False, kernel: Optional[str] = None, tile=None, slab: int = 0,
              run_len: Optional[int] = None, out: Optional[torch.Tensor] = None,
              tile_range: Optional[tuple] = None, plan: Optional[TilePlan] = None) -> torch.Tensor:
    """out[..., i, :] = epi(sum_s w[s] * x[..., src[s], :] + bias) over the CSR of ``graph``.
    ``tile_range=(begin, count)`` (tiled kernel only) restricts the launch to those tiles of the
    plan; the other rows of ``out`` are left untouched."""
    _require_cuda(x, "x")
    x3, lead = _as_3d(x)
    b, n_src, f = x3.shape
    if n_src != graph.n_src:
        raise ValueError("x has %d rows, graph expects %d source rows" % (n_src, graph.n_src))
    code = dtype_code(x3.dtype)
    kernel = kernel or DEFAULT_AGG_KERNEL
    esz = x3.element_size()
    if kernel == "auto":
        if (graph.is_plain_mesh or graph.is_masked_mesh) and (f * esz) % 16 == 0:
            kernel = "stencil"
        elif graph.grid_shape is not None and graph.n_src == graph.n_dst and (f * esz) % 16 == 0 and \
                plan is None and tile_range is None:
            kernel = "tiled_or_rows"
        elif LOCALITY_TILES and graph.grid_shape is None and graph.n_src == graph.n_dst and \
                (f * esz) % 16 == 0 and plan is None and tile_range is None and \
                graph.n_dst >= LOCALITY_MIN_NODES and graph.num_messages <= LOCALITY_MAX_DEGREE * graph.n_dst:
            # a large sparse graph whose numbering is not a grid's: compact tiles found from the CSR
            # (GraphCSR.locality_plan, once per graph) + the staged kernel's row-gather producer.  The plan costs
            # ~10-50 ms, so it is built at the SECOND call on a graph handle: a graph that is used once (a loader
            # handing a fresh edge_index every iteration, models_gnn.py:359) never pays for it
            graph.auto_calls += 1
            kernel = "locality_or_rows" if graph.auto_calls >= 2 or ("locality", None) in graph._plans else "rows"
        else:
            kernel = "rows"
    bias32 = _bias32(bias)
    with torch.cuda.device(x3.device):
        if out is None:
            out = torch.empty((b, graph.n_dst, f), dtype=x3.dtype, device=x3.device)
        epi = _lib.EPI_RELU if relu else _lib.EPI_NONE
        if kernel in ("locality", "locality_or_rows"):
            lplan = plan if plan is not None else graph.locality_plan()
            rc = _lib.GWEN_E_NOSUPPORT
            if lplan is not None:
                rc = lib().gwen_aggregate_tiled_fwd(C.byref(lplan.struct), _ptr(x3), _ptr(out), b, n_src, f, f,
                                                    n_src * f, f, graph.n_dst * f, code, _ptr(bias32), epi, slab,
                                                    0, 0, _stream())
            if rc == _lib.GWEN_E_NOSUPPORT and kernel == "locality_or_rows":
                kernel = "rows"     # no locality in this graph, or its tiles do not fit in shared memory
            elif lplan is None:
                raise RuntimeError("this graph has no locality plan (GraphCSR.locality_plan() is None)")
            else:
                check(rc, "gwen_aggregate_tiled_fwd")
                return out.reshape(tuple(lead) + (graph.n_dst, f))
        if kernel == "tiled_or_rows":
            # the TMA-staged kernel (bitwise equal to "rows", ~2x faster on grid-numbered graphs) unless
            # its stages do not fit in shared memory for this width
            plan = graph.tile_plan(tile, run_len)
            rc = lib().gwen_aggregate_tiled_fwd(C.byref(plan.struct), _ptr(x3), _ptr(out), b, n_src, f, f,
                                                n_src * f, f, graph.n_dst * f, code, _ptr(bias32), epi, slab,
                                                0, 0, _stream())
            if rc == _lib.GWEN_E_NOSUPPORT:
                kernel = "rows"
            else:
                check(rc, "gwen_aggregate_tiled_fwd")
                return out.reshape(tuple(lead) + (graph.n_dst, f))
        if kernel == "tiled":
            if plan is None:
                plan = graph.tile_plan(tile, run_len)
            check(lib().gwen_aggregate_tiled_fwd(C.byref(plan.struct), _ptr(x3),
                                                 _ptr(out), b, n_src, f, f, n_src * f, f,
                                                 graph.n_dst * f, code, _ptr(bias32), epi, slab,
                                                 tile_range[0] if tile_range else 0,
                                                 tile_range[1] if tile_range else 0,
                                                 _stream()), "gwen_aggregate_tiled_fwd")
        elif kernel == "stencil":
            masked = graph.is_masked_mesh
            if not (graph.is_plain_mesh or masked):
                raise RuntimeError("the stencil kernel needs a plain (or masked) H x W mesh graph")
            h, w = graph.grid_shape
            dpad = graph.dis_padded_masked() if masked else graph.dis_padded()
            check(lib().gwen_grid_stencil_fwd(_ptr(x3), _ptr(out), _ptr(dpad), dpad.shape[1], dpad.shape[0], b, h, h, w,
                                              0, f, f, n_src * f, f, graph.n_dst * f, code,
                                              _ptr(bias32), epi, slab, tile[0] if tile else 0,
                                              _stream()), "gwen_grid_stencil_fwd")
            if masked:   # cut-out nodes keep only their self loop: out = epi(x + bias)
                idx = graph.masked_idx
                check(lib().gwen_rows_self_fwd(_ptr(x3), _ptr(out), _ptr(idx), idx.numel(), b, f, f, n_src * f, f,
                                               graph.n_dst * f, code, _ptr(bias32), epi, _stream()),
                      "gwen_rows_self_fwd")
        elif kernel == "rows":
            check(lib().gwen_aggregate_fwd(_ptr(graph.rowptr), _ptr(graph.src), _ptr(graph.w),
                                           _ptr(graph.order), _ptr(x3), _ptr(out), b, graph.n_dst,
                                           n_src, f, f, n_src * f, f, graph.n_dst * f, code,
                                           _ptr(bias32), epi, _stream()), "gwen_aggregate_fwd")
        else:
            raise ValueError("unknown aggregate kernel %r" % kernel)
    return out.reshape(tuple(lead) + (graph.n_dst, f))


def mesh_stencil(x: torch.Tensor, dis_bordered: torch.Tensor, hs: int, hd: int, w: int, row_off: int,
                 bias: Optional[torch.Tensor] = None, relu: bool = False,
                 out: Optional[torch.Tensor] = None, slab: int = 0, tile_w: int = 0) -> torch.Tensor:
    """Mesh fast path on explicit geometry: x [B, hs*w, F] -> out [B, hd*w, F], destination row r
    reads source rows r + row_off - 1 .. r + row_off + 1 (see gwen_grid_stencil_fwd)."""
    _require_cuda(x, "x")
    x3, lead = _as_3d(x)
    b, n_src, f = x3.shape
    if n_src != hs * w:
        raise ValueError("x has %d rows, expected hs * w = %d" % (n_src, hs * w))
    code = dtype_code(x3.dtype)
    bias32 = _bias32(bias)
    with torch.cuda.device(x3.device):
        if out is None:
            out = torch.empty((b, hd * w, f), dtype=x3.dtype, device=x3.device)
        o3 = out if out.dim() == 3 else out.unsqueeze(0)
        assert o3.is_contiguous() or o3.stride(-1) == 1
        check(lib().gwen_grid_stencil_fwd(_ptr(x3), _ptr(o3), _ptr(dis_bordered), dis_bordered.shape[1],
                                          dis_bordered.shape[0], b, hs, hd, w, row_off, f, f, n_src * f, f, o3.stride(0) if b > 1 else hd * w * f,
                                          code, _ptr(bias32), _lib.EPI_RELU if relu else 0, slab, tile_w,
                                          _stream()), "gwen_grid_stencil_fwd")
    return out


def _rows_view(t: torch.Tensor):
    """(pointer-compatible [B, rows, F] view, B, rows) of a tensor whose last two dims are a dense
    row block (row pitch F) and whose leading dims collapse to one batch stride; None if not."""
    if t.dim() < 2 or t.stride(-1) != 1 or t.stride(-2) != t.shape[-1]:
        return None
    if t.dim() == 2:
        return t.unsqueeze(0)
    if t.dim() == 3:
        return t
    lead = t.shape[:-2]
    try:
        return t.view((-1,) + tuple(t.shape[-2:]))
    except RuntimeError:
        return None


# The fused layer kernel pays off once every CTA pair has a long queue of 8 x 32 tiles (measured on
# 1158 x 774 meshes: -4 % .. -11 % vs the two-kernel path; on 582 x 390 it is slower).
FUSED_MIN_ITEMS = 3500
FUSED_MAX_K = 512       # widest input the fused kernel keeps resident (one A buffer above 256)
import os as _os
FUSED_MIN_K = int(_os.environ.get("GWEN_FUSED_MIN_K", "64"))   # narrowest input that takes the fused kernel


def gcn_fused_supported(graph: GraphCSR, x: torch.Tensor, weight: torch.Tensor) -> bool:
    """True when ``gcn_fused`` can serve this layer: plain mesh, bf16, k_in in {64, 128, .., 512},
    n_out a multiple of 128 (and GWEN_NO_FUSED unset)."""
    import os
    k, n = weight.shape[1], weight.shape[0]
    return (os.environ.get("GWEN_NO_FUSED") is None and graph.is_plain_mesh and x.dtype == torch.bfloat16
            and x.is_cuda and max(64, FUSED_MIN_K) <= k <= FUSED_MAX_K and k % 64 == 0 and n % 128 == 0 and n <= 8192)


def gcn_fused_preferred(graph: GraphCSR, x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Supported AND large enough to beat the two-kernel path (the layer's automatic choice)."""
    if not gcn_fused_supported(graph, x, weight):
        return False
    h, w = graph.grid_shape
    batch = x.numel() // max(1, x.shape[-1] * x.shape[-2])
    return batch * ((h + 7) // 8) * ((w + 31) // 32) >= FUSED_MIN_ITEMS


def gcn_fused(graph: GraphCSR, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
              relu: bool = False, out: Optional[torch.Tensor] = None, pre_bias: Optional[torch.Tensor] = None,
              pre_relu: bool = False) -> torch.Tensor:
    """y = epi((A_hat x) W^T + bias) in one kernel (gwen_gcn_fused_fwd); x [..., N, K] bf16.
    ``pre_bias`` / ``pre_relu``: y = epi(pre(A_hat x + pre_bias) W^T + bias) -- the previous layer's bias and
    ReLU applied to the aggregated row inside the kernel (that layer then only runs its projection)."""
    _require_cuda(x, "x")
    x3, lead = _as_3d(x)
    b, n, k = x3.shape
    h, w = graph.grid_shape
    n_out = weight.shape[0]
    wt = _cast_cached(weight, x3.dtype)
    bias32 = _bias32(bias)
    pre32 = _bias32(pre_bias)
    dpad = graph.dis_padded()
    with torch.cuda.device(x3.device):
        if out is None:
            out = torch.empty((b, n, n_out), dtype=x3.dtype, device=x3.device)
        assert out.is_contiguous()
        check(lib().gwen_gcn_fused_fwd(_ptr(x3), _ptr(wt), _ptr(out), _ptr(dpad), dpad.shape[1], dpad.shape[0], b, h, w, k,
                                       n_out, dtype_code(x3.dtype), _ptr(bias32),
                                       _lib.EPI_RELU if relu else 0, _ptr(pre32), _lib.EPI_RELU if pre_relu else 0,
                                       _stream()), "gwen_gcn_fused_fwd")
    return out.reshape(tuple(lead) + (n, n_out))


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           relu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epi(x @ weight.T + bias); x [..., K], weight [N_out, K].  ``out`` ([..., N_out], rows dense,
    any batch stride -- e.g. the owned rows of a band buffer) receives the result in place: one launch
    per batch slice when x or out is batch-strided."""
    _require_cuda(x, "x")
    _require_cuda(weight, "weight")
    k = x.shape[-1]
    n_out = weight.shape[0]
    if weight.shape[1] != k:
        raise ValueError("weight is %s, x has %d features" % (tuple(weight.shape), k))
    wt = _cast_cached(weight, x.dtype)
    code = dtype_code(x.dtype)
    bias32 = _bias32(bias)
    if out is not None or not x.is_contiguous():
        xv = _rows_view(x)
        if xv is None:
            xv = _rows_view(x.contiguous())
        if out is None:
            out = torch.empty(tuple(x.shape[:-1]) + (n_out,), dtype=x.dtype, device=x.device)
        ov = _rows_view(out)
        if ov is None or ov.shape[:2] != xv.shape[:2] or ov.shape[2] != n_out or out.dtype != x.dtype:
            raise ValueError("out must be [..., N_out] with dense rows matching x")
        with torch.cuda.device(x.device):
            dense = xv.shape[0] == 1 or (xv.stride(0) == xv.shape[1] * k and ov.stride(0) == ov.shape[1] * n_out)
            epi = _lib.EPI_RELU if relu else 0
            if x.dtype == torch.float32 or dense:
                # fp32 goes through the workspace entry (3xTF32 on the tensor cores when it applies), one call
                # per batch slice when the rows are batch-strided
                slices = [(xv, ov, xv.shape[0] * xv.shape[1])] if dense else \
                    [(xv[b], ov[b], xv.shape[1]) for b in range(xv.shape[0])]
                for xs, os_, m in slices:
                    need = C.c_size_t(0)
                    if x.dtype == torch.float32:
                        check(lib().gwen_linear_fwd_workspace_bytes(m, k, n_out, code, C.byref(need)), "linear ws")
                    ws = torch.empty(need.value, dtype=torch.uint8, device=x.device) if need.value else None
                    check(lib().gwen_linear_fwd_ws(_ptr(xs), _ptr(wt), _ptr(os_), m, k, n_out, k, k, n_out, code,
                                                   _ptr(bias32), epi, _ptr(ws), need.value, _stream()),
                          "gwen_linear_fwd")
            else:       # bf16, batch-strided rows: one launch, the batch index is a TMA coordinate
                check(lib().gwen_linear_batched_fwd(_ptr(xv), _ptr(wt), _ptr(ov), xv.shape[0], xv.shape[1], k,
                                                    n_out, k, k, n_out, xv.stride(0), ov.stride(0), code,
                                                    _ptr(bias32), epi, _stream()), "gwen_linear_batched_fwd")
        return out
    x2 = x.reshape(-1, k).contiguous()
    with torch.cuda.device(x2.device):
        y = torch.empty((x2.shape[0], n_out), dtype=x2.dtype, device=x2.device)
        need = C.c_size_t(0)
        if x2.dtype == torch.float32:      # fp32 on the tensor cores (3xTF32) needs scratch for the split operands
            check(lib().gwen_linear_fwd_workspace_bytes(x2.shape[0], k, n_out, code, C.byref(need)), "linear ws")
        ws = torch.empty(need.value, dtype=torch.uint8, device=x2.device) if need.value else None
        check(lib().gwen_linear_fwd_ws(_ptr(x2), _ptr(wt), _ptr(y), x2.shape[0], k, n_out, k, k, n_out,
                                       code, _ptr(bias32), _lib.EPI_RELU if relu else 0, _ptr(ws), need.value,
                                       _stream()), "gwen_linear_fwd")
    return y.reshape(tuple(x.shape[:-1]) + (n_out,))


def linear_b2b_supported(x: torch.Tensor, weight1: torch.Tensor, weight2: torch.Tensor) -> bool:
    """True when ``linear_b2b`` can serve this pair of projections (bf16, k1 in {64, .., 512}, n1 % 256 == 0,
    n2 % 64 == 0 and n2 <= 256 or n2 % 256 == 0, at least 256 rows; GWEN_NO_B2B unset)."""
    import os
    if os.environ.get("GWEN_NO_B2B") is not None or not x.is_cuda or x.dtype != torch.bfloat16:
        return False
    m = x.numel() // max(1, x.shape[-1])
    return bool(lib().gwen_linear_b2b_supported(m, weight1.shape[1], weight1.shape[0], weight2.shape[0],
                                                dtype_code(x.dtype)))


def linear_b2b(x: torch.Tensor, weight1: torch.Tensor, bias1: Optional[torch.Tensor], relu1: bool,
               weight2: torch.Tensor, bias2: Optional[torch.Tensor] = None, relu2: bool = False,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epi2(epi1(x @ weight1.T + bias1) @ weight2.T + bias2) in one kernel (gwen_linear_b2b_fwd): the hidden
    tensor ``epi1(...)`` ([..., N1]) never reaches HBM.  x [..., K1] bf16, weight1 [N1, K1], weight2 [N2, N1]."""
    _require_cuda(x, "x")
    _require_cuda(weight1, "weight1")
    _require_cuda(weight2, "weight2")
    k1 = x.shape[-1]
    n1, n2 = weight1.shape[0], weight2.shape[0]
    if weight1.shape[1] != k1 or weight2.shape[1] != n1:
        raise ValueError("weights are %s and %s, x has %d features" % (tuple(weight1.shape), tuple(weight2.shape), k1))
    w1 = _cast_cached(weight1, x.dtype)
    w2 = _cast_cached(weight2, x.dtype)
    b1, b2 = _bias32(bias1), _bias32(bias2)
    x2 = x.reshape(-1, k1)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    m = x2.shape[0]
    with torch.cuda.device(x2.device):
        if out is None:
            out = torch.empty(tuple(x.shape[:-1]) + (n2,), dtype=x.dtype, device=x.device)
        if not out.is_contiguous() or out.dtype != x.dtype or out.numel() != m * n2:
            raise ValueError("out must be a contiguous [..., N2] tensor of x's dtype")
        check(lib().gwen_linear_b2b_fwd(_ptr(x2), _ptr(w1), _ptr(b1), _lib.EPI_RELU if relu1 else 0, _ptr(w2), _ptr(b2),
                                        _lib.EPI_RELU if relu2 else 0, _ptr(out), m, k1, n1, n2, k1, n2,
                                        dtype_code(x.dtype), _stream()), "gwen_linear_b2b_fwd")
    return out


def linear_bwd_data(dy: torch.Tensor, weight: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = dy @ weight; dy [..., N_out], weight [N_out, K].  ``out`` as in :func:`linear`."""
    n_out, k = weight.shape
    if out is not None:
        dv, ov = _rows_view(dy if dy.is_contiguous() else dy.contiguous()), _rows_view(out)
        if ov is None or ov.shape[:2] != dv.shape[:2] or ov.shape[2] != k or out.dtype != dy.dtype:
            raise ValueError("out must be [..., K] with dense rows matching dy")
        wt = _cast_cached(weight, dy.dtype)
        with torch.cuda.device(dy.device):
            if dy.dtype == torch.float32:     # workspace entry per slice (3xTF32 when it applies)
                for b in range(dv.shape[0]):
                    need = C.c_size_t(0)
                    check(lib().gwen_linear_bwd_data_workspace_bytes(dv.shape[1], k, n_out, dtype_code(dy.dtype),
                                                                     C.byref(need)), "dgrad ws")
                    ws = torch.empty(need.value, dtype=torch.uint8, device=dy.device) if need.value else None
                    check(lib().gwen_linear_bwd_data_ws(_ptr(dv[b]), _ptr(wt), _ptr(ov[b]), dv.shape[1], k, n_out,
                                                        n_out, k, k, dtype_code(dy.dtype), _ptr(ws), need.value,
                                                        _stream()), "gwen_linear_bwd_data")
            else:
                check(lib().gwen_linear_batched_bwd_data(_ptr(dv), _ptr(wt), _ptr(ov), dv.shape[0], dv.shape[1], k,
                                                         n_out, n_out, k, k, dv.stride(0), ov.stride(0),
                                                         dtype_code(dy.dtype), _stream()),
                      "gwen_linear_batched_bwd_data")
        return out
    dy2 = dy.reshape(-1, n_out).contiguous()
    wt = _cast_cached(weight, dy2.dtype)
    with torch.cuda.device(dy2.device):
        dx = torch.empty((dy2.shape[0], k), dtype=dy2.dtype, device=dy2.device)
        need = C.c_size_t(0)
        if dy2.dtype == torch.float32:     # fp32 on the tensor cores (3xTF32) needs scratch for the split W^T
            check(lib().gwen_linear_bwd_data_workspace_bytes(dy2.shape[0], k, n_out, dtype_code(dy2.dtype),
                                                             C.byref(need)), "dgrad ws")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dy2.device) if need.value else None
        check(lib().gwen_linear_bwd_data_ws(_ptr(dy2), _ptr(wt), _ptr(dx), dy2.shape[0], k, n_out, n_out,
                                            k, k, dtype_code(dy2.dtype), _ptr(ws), need.value, _stream()),
              "gwen_linear_bwd_data")
    return dx.reshape(tuple(dy.shape[:-1]) + (k,))


def linear_bwd_data_masked(dy: torch.Tensor, weight: torch.Tensor, mask: torch.Tensor,
                           out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """``(dy @ weight) * (mask > 0)`` in ONE kernel (``gwen_linear_bwd_data_masked``: the previous layer's ReLU
    backward in the dgrad epilogue), ``mask`` = the layer input (bf16, shape of the result).  Returns None when the
    fused kernel does not serve the problem (not bf16, shapes outside the tcgen05 pair kernel): the caller then runs
    ``linear_bwd_data`` and leaves the mask to the previous layer's ``relu_bias_bwd`` -- the same bits."""
    n_out, k = weight.shape
    if dy.dtype != torch.bfloat16 or mask.dtype != torch.bfloat16 or mask.shape[-1] != k:
        return None
    if out is not None:   # batch-strided destination (dense rows), e.g. the owned rows of a band buffer: one launch
        dv, mv, ov = _rows_view(dy if dy.is_contiguous() else dy.contiguous()), _rows_view(mask), _rows_view(out)
        if dv is None or mv is None or ov is None or out.dtype != dy.dtype or ov.shape[2] != k or \
                not (dv.shape[:2] == mv.shape[:2] == ov.shape[:2]):
            return None
        wt = _cast_cached(weight, dy.dtype)
        nb = dv.shape[0]
        with torch.cuda.device(dy.device):
            rc = lib().gwen_linear_batched_bwd_data_masked(
                _ptr(dv), _ptr(wt), _ptr(ov), _ptr(mv), nb, dv.shape[1], k, n_out, n_out, k, k, k,
                dv.stride(0) if nb > 1 else 0, ov.stride(0) if nb > 1 else 0, mv.stride(0) if nb > 1 else 0,
                dtype_code(dy.dtype), _stream())
        if rc == _lib.GWEN_E_NOSUPPORT:
            return None
        check(rc, "gwen_linear_batched_bwd_data_masked")
        return out
    dy2 = dy.reshape(-1, n_out).contiguous()
    m2 = mask.reshape(-1, k)
    if m2.shape[0] != dy2.shape[0] or not m2.is_contiguous():
        return None
    wt = _cast_cached(weight, dy2.dtype)
    with torch.cuda.device(dy2.device):
        dx = torch.empty((dy2.shape[0], k), dtype=dy2.dtype, device=dy2.device)
        rc = lib().gwen_linear_bwd_data_masked(_ptr(dy2), _ptr(wt), _ptr(dx), _ptr(m2), dy2.shape[0], k, n_out, n_out,
                                               k, k, k, dtype_code(dy2.dtype), _stream())
    if rc == _lib.GWEN_E_NOSUPPORT:
        return None
    check(rc, "gwen_linear_bwd_data_masked")
    return dx.reshape(tuple(dy.shape[:-1]) + (k,))


def linear_bwd_weight(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[n, k] = sum_m dy[m, n] x[m, k] in fp32 (fixed split order -> deterministic).  ``out``: a
    contiguous fp32 [n, k] destination (e.g. a slice of a flat gradient bucket)."""
    n_out, k = dy.shape[-1], x.shape[-1]
    dy2 = dy.reshape(-1, n_out).contiguous()
    x2 = x.reshape(-1, k).contiguous()
    m = dy2.shape[0]
    with torch.cuda.device(dy2.device):
        need = C.c_size_t()
        check(lib().gwen_linear_bwd_weight_workspace_bytes(m, k, n_out, C.byref(need)), "wgrad ws")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dy2.device)
        if out is not None:
            if out.dtype != torch.float32 or tuple(out.shape) != (n_out, k) or not out.is_contiguous():
                raise ValueError("out must be a contiguous fp32 [%d, %d] tensor" % (n_out, k))
            dw = out
        else:
            dw = torch.empty((n_out, k), dtype=torch.float32, device=dy2.device)
        check(lib().gwen_linear_bwd_weight(_ptr(dy2), _ptr(x2), _ptr(dw), m, k, n_out, n_out, k, k,
                                           dtype_code(dy2.dtype), _ptr(ws), need.value, _stream()),
              "gwen_linear_bwd_weight")
    return dw


def linear_bwd_weight_bias(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None):
    """``(dw, db)``: :func:`linear_bwd_weight` and ``db[n] = sum_m dy[m, n]`` (fp32) out of the SAME pass over ``dy``
    (``gwen_linear_bwd_weight_bias``: a ones-column MMA in the tcgen05 wgrad).  Returns None when the fused kernel
    does not serve the problem (not bf16 / shapes outside the tensor-core wgrad): run ``linear_bwd_weight`` and
    ``relu_bias_bwd`` / ``bias_grad`` instead."""
    n_out, k = dy.shape[-1], x.shape[-1]
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
        return None
    dy2 = dy.reshape(-1, n_out).contiguous()
    x2 = x.reshape(-1, k).contiguous()
    m = dy2.shape[0]
    with torch.cuda.device(dy2.device):
        need = C.c_size_t()
        check(lib().gwen_linear_bwd_weight_bias_workspace_bytes(m, k, n_out, C.byref(need)), "wgrad+bias ws")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dy2.device)
        if out is not None:
            if out.dtype != torch.float32 or tuple(out.shape) != (n_out, k) or not out.is_contiguous():
                raise ValueError("out must be a contiguous fp32 [%d, %d] tensor" % (n_out, k))
            dw = out
        else:
            dw = torch.empty((n_out, k), dtype=torch.float32, device=dy2.device)
        db = torch.empty(n_out, dtype=torch.float32, device=dy2.device)
        rc = lib().gwen_linear_bwd_weight_bias(_ptr(dy2), _ptr(x2), _ptr(dw), _ptr(db), m, k, n_out, n_out, k, k,
                                               dtype_code(dy2.dtype), _ptr(ws), need.value, _stream())
    if rc == _lib.GWEN_E_NOSUPPORT:
        return None
    check(rc, "gwen_linear_bwd_weight_bias")
    return dw, db


def relu_bwd_(y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """In place: dy[y <= 0] = 0."""
    f = y.shape[-1]
    y2, dy2 = y.reshape(-1, f), dy.reshape(-1, f)
    assert y2.is_contiguous() and dy2.is_contiguous()
    with torch.cuda.device(y.device):
        check(lib().gwen_relu_bwd(_ptr(y2), _ptr(dy2), y2.shape[0], f, f, f, dtype_code(y.dtype),
                                  _stream()), "gwen_relu_bwd")
    return dy


def bias_grad(dy: torch.Tensor) -> torch.Tensor:
    f = dy.shape[-1]
    dy2 = dy.reshape(-1, f).contiguous()
    with torch.cuda.device(dy2.device):
        need = C.c_size_t()
        check(lib().gwen_bias_grad_workspace_bytes(dy2.shape[0], f, C.byref(need)), "bias ws")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dy2.device)
        db = torch.empty(f, dtype=torch.float32, device=dy2.device)
        check(lib().gwen_bias_grad(_ptr(dy2), _ptr(db), dy2.shape[0], f, f, dtype_code(dy2.dtype),
                                   _ptr(ws), need.value, _stream()), "gwen_bias_grad")
    return db


def copy_rows_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[...] = src[...] for [..., N, F] tensors with dense rows and any batch stride: one dense
    block copy per batch slice (torch's generic strided copy kernel is ~5x slower on these)."""
    dv, sv = _rows_view(dst), _rows_view(src)
    if dv is None or sv is None or dv.shape != sv.shape:
        dst.copy_(src)
        return dst
    for b in range(dv.shape[0]):
        dv[b].copy_(sv[b])
    return dst


def relu_bias_bwd(dy: torch.Tensor, y: Optional[torch.Tensor], want_db: bool,
                  out: Optional[torch.Tensor] = None):
    """(dz, db): dz = dy * (y > 0) (dy itself when ``y`` is None) and db = column sums of dz in fp32
    (None unless ``want_db``), fused in one pass when the width allows, else the two kernels.
    ``dy`` / ``y`` / ``out`` may be batch-strided (dense rows): the pass then runs per batch slice, and
    ``out`` (e.g. the owned rows of a band buffer) receives dz without a staging copy."""
    f = dy.shape[-1]
    vn = 16 // dy.element_size()
    groups = f // vn
    fused = f % vn == 0 and 0 < groups <= 256 and 256 % groups == 0
    if y is None and not want_db:
        return (dy if out is None else copy_rows_(out, dy)), None
    dv = _rows_view(dy)
    yv = None if y is None else _rows_view(y)
    ov = None if out is None else _rows_view(out)
    if not fused or dv is None or (y is not None and yv is None) or (out is not None and ov is None):
        dyc = dy.contiguous()
        dz = relu_bwd_(y.contiguous(), dyc.clone()) if y is not None else dyc
        db = bias_grad(dz) if want_db else None
        return (dz if out is None else copy_rows_(out, dz)), db
    if yv is None and ov is None:
        dz_full = dy                                   # no mask, no destination: dz is dy itself
    else:
        dz_full = out if out is not None else torch.empty(dy.shape, dtype=dy.dtype, device=dy.device)
    zv = _rows_view(dz_full)
    nb, rows = dv.shape[0], dv.shape[1]
    dense = all(t is None or nb == 1 or t.stride(0) == rows * f for t in (dv, yv, zv))
    db_total = None
    with torch.cuda.device(dy.device):
        need = C.c_size_t(0)
        ws = None
        n_rows = nb * rows if dense else rows
        if want_db:
            check(lib().gwen_bias_grad_workspace_bytes(n_rows, f, C.byref(need)), "bias ws")
            ws = torch.empty(need.value, dtype=torch.uint8, device=dy.device)
        for i in range(1 if dense else nb):
            d_i = dv.reshape(-1, f) if dense else dv[i]
            y_i = None if yv is None else (yv.reshape(-1, f) if dense else yv[i])
            z_i = zv.reshape(-1, f) if dense else zv[i]
            if yv is None and z_i.data_ptr() != d_i.data_ptr():
                z_i.copy_(d_i)                         # no mask: dz = dy into the destination
            db = torch.empty(f, dtype=torch.float32, device=dy.device) if want_db else None
            check(lib().gwen_relu_bias_bwd(_ptr(y_i), _ptr(d_i), _ptr(z_i) if y_i is not None else None, _ptr(db),
                                           n_rows, f, dtype_code(dy.dtype), _ptr(ws), need.value, _stream()),
                  "gwen_relu_bias_bwd")
            if want_db:
                db_total = db if db_total is None else db_total + db
    return dz_full, db_total


def rows_gather(x: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """buf[b, j, :] = x[b, idx[j], :] for x [B, N, F] (halo pack)."""
    b, n, f = x.shape
    assert x.is_contiguous() and idx.dtype == torch.int32
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty((b, idx.numel(), f), dtype=x.dtype, device=x.device)
        check(lib().gwen_rows_gather(_ptr(x), _ptr(idx), _ptr(out), b, idx.numel(), f, f, n * f,
                                     dtype_code(x.dtype), _stream()), "gwen_rows_gather")
    return out


def rows_scatter_(x: torch.Tensor, idx: torch.Tensor, buf: torch.Tensor) -> torch.Tensor:
    """x[b, idx[j], :] = buf[b, j, :] (halo unpack)."""
    b, n, f = x.shape
    assert x.is_contiguous() and buf.is_contiguous() and idx.dtype == torch.int32
    with torch.cuda.device(x.device):
        check(lib().gwen_rows_scatter(_ptr(buf), _ptr(idx), _ptr(x), b, idx.numel(), f, f, n * f,
                                      dtype_code(x.dtype), _stream()), "gwen_rows_scatter")
    return x
