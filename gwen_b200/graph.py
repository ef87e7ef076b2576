"""Graph side of the hot path: edge-list builders, the one-time CSR preprocessor (K0) and tile plans.

Host-side mirror of what PyG does inside ``GCNConv.forward`` before the arithmetic starts
(``gcn_norm`` / ``add_remaining_self_loops``; reference call sites ``src/gwen/models_gnn.py:147-149,
204-206``) and of the graph builder the reference dataset uses (``erdos_renyi_graph(N, 1)`` at
``src/gwen/utils.py:176``).  All work happens in ``libgwen_b200.so`` on the GPU; tensors are only
used to own device memory.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib

__all__ = ["GraphCSR", "build_graph", "get_graph", "clear_graph_cache", "grid", "grid_edge_count",
           "erdos_renyi_graph", "complete_graph", "TilePlan"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("gwen_b200: %s must be a CUDA tensor (there is no CPU fallback)" % name)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# builders
# ---------------------------------------------------------------------------------------------
def grid_edge_count(height: int, width: int) -> int:
    return int(lib().gwen_grid_edge_count(height, width))


def grid(height: int, width: int, device="cuda") -> torch.Tensor:
    """``torch_geometric.utils.grid(height, width)`` edge_index (8-neighbour mesh + self loops,
    sorted by (row, col)), generated on the device.  SURVEY.md Appendix B.2."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gwen_b200.grid builds on a CUDA device only")
    e = grid_edge_count(height, width)
    with torch.cuda.device(device):
        out = torch.empty((2, e), dtype=torch.int64, device=device)
        check(lib().gwen_grid_edges(height, width, out.data_ptr(), _stream()), "gwen_grid_edges")
    return out


def complete_graph(num_nodes: int, device="cuda") -> torch.Tensor:
    """All ordered pairs (i, j), i != j, sorted by (row, col): the value of
    ``erdos_renyi_graph(N, edge_prob=1)``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gwen_b200.complete_graph builds on a CUDA device only")
    with torch.cuda.device(device):
        out = torch.empty((2, num_nodes * (num_nodes - 1)), dtype=torch.int64, device=device)
        check(lib().gwen_complete_edges(num_nodes, _ptr(out), _stream()), "gwen_complete_edges")
    return out


def erdos_renyi_graph(num_nodes: int, edge_prob: float, directed: bool = False,
                      device="cuda") -> torch.Tensor:
    """Drop-in for the call at reference ``src/gwen/utils.py:176``.  Only ``edge_prob >= 1`` (the
    value GWEN uses) is served; like PyG it draws ``N(N-1)/2`` numbers from the global CPU RNG so
    that a seeded script initialises its weights identically afterwards."""
    if edge_prob < 1.0 or directed:
        raise NotImplementedError("gwen_b200.erdos_renyi_graph serves edge_prob=1, undirected "
                                  "(the only form the reference calls)")
    torch.rand(num_nodes * (num_nodes - 1) // 2)  # RNG side effect of the reference builder
    return complete_graph(num_nodes, device)


# ---------------------------------------------------------------------------------------------
# CSR handle
# ---------------------------------------------------------------------------------------------
class TilePlan:
    """Device arrays of a tile plan + the host struct handed to gwen_aggregate_tiled_fwd."""

    def __init__(self, order, tile_ptr, run_ptr, run_start, trec, tmsg, tmsg_base, n_dst, run_len,
                 max_tile_runs, max_tile_rows, max_tile_msgs, total_runs, total_src, flags: int = 0):
        self.order, self.tile_ptr, self.run_ptr, self.run_start = order, tile_ptr, run_ptr, run_start
        self.trec, self.tmsg, self.tmsg_base = trec, tmsg, tmsg_base
        self.num_tiles = tile_ptr.numel() - 1
        self.n_dst, self.run_len, self.max_tile_runs = n_dst, run_len, max_tile_runs
        self.max_tile_rows, self.max_tile_msgs = max_tile_rows, max_tile_msgs
        self.total_runs, self.total_src, self.flags = total_runs, total_src, flags
        self.struct = _lib.TilePlanStruct(self.num_tiles, run_len, max_tile_runs, max_tile_rows,
                                          max_tile_msgs, flags, n_dst, _ptr(tile_ptr), _ptr(run_ptr),
                                          _ptr(run_start), _ptr(trec), _ptr(tmsg), _ptr(tmsg_base))

    @property
    def staging_efficiency(self) -> float:
        """distinct staged sources / staged rows (1.0 = every staged row is used)."""
        return self.total_src / max(1, self.total_runs * self.run_len)

    @property
    def amplification(self) -> float:
        """staged rows per destination row: the L2->SM read traffic relative to one pass."""
        return self.total_runs * self.run_len / max(1, self.n_dst)


def bordered_dis(dis2d: torch.Tensor) -> torch.Tensor:
    """[H, W] -> zero-bordered fp32 [H + 12, pitch] as gwen_grid_stencil_fwd expects.  The kernel
    always bulk-copies the 10 dis rows of an 8-row tile, also for a sub-range launch whose last tile
    starts at destination row ``round_up(hd, 8) - 8`` with ``row_off + hd <= H + 1``: the last row it
    touches is ``round_up(hd, 8) + row_off + 1 <= H + 9``, so H + 12 rows cover every legal launch
    (the C entry point checks ``dis_rows``)."""
    h, w = dis2d.shape
    pitch = (w + 128 + 8 + 3) // 4 * 4
    d = torch.zeros((h + 12, pitch), dtype=torch.float32, device=dis2d.device)
    d[1:h + 1, 1:w + 1] = dis2d
    return d


class GraphCSR:
    """Destination-sorted CSR of ``edge_index'`` (after self-loop normalisation) with GCN weights.

    ``rowptr int32[N+1]``, ``src int32[E']``, ``w fp32[E']``, ``dis fp32[N]``, ``perm int64[E']``.
    ``n_src`` may exceed ``n_dst`` for a partition-local graph (owned rows + halo rows).
    """

    def __init__(self, rowptr, src, w, dis, perm, n_dst, n_src, num_messages, flags,
                 edge_index=None, grid_shape=None):
        self.rowptr, self.src, self.w, self.dis, self.perm = rowptr, src, w, dis, perm
        self.n_dst, self.n_src, self.num_messages, self.flags = n_dst, n_src, num_messages, flags
        self.edge_index = edge_index
        self.grid_shape = grid_shape
        # set by build_graph's detection: "plain" (the full H x W mesh), "masked" (mesh minus cut-out
        # nodes: mesh_valid / masked_idx below), or None (only the node numbering is grid-like, or unknown)
        self.mesh_kind: Optional[str] = None
        self.mesh_valid: Optional[torch.Tensor] = None     # bool [N]: node kept in the masked mesh
        self.masked_idx: Optional[torch.Tensor] = None     # int32 ids of the cut-out nodes
        self._transposed: Optional["GraphCSR"] = None
        self._plans: Dict[Tuple, TilePlan] = {}
        self.order: Optional[torch.Tensor] = None  # locality order for the row kernel
        self.auto_calls = 0     # ops.aggregate(kernel="auto") calls on a graph that is not grid-numbered (see ops.py)

    @property
    def device(self):
        return self.rowptr.device

    # -- mesh fast path ------------------------------------------------------------------------
    @property
    def is_plain_mesh(self) -> bool:
        """True when this is exactly the H x W 8-neighbour mesh with default GCN normalisation
        (self loops added, weight 1) over all of its nodes: the stencil kernel applies."""
        return (self.mesh_kind == "plain" and self.grid_shape is not None and self.dis is not None
                and self.n_src == self.n_dst and self.flags == _lib.GRAPH_ADD_SELF_LOOPS
                and self.grid_shape[0] * self.grid_shape[1] == self.n_dst)

    @property
    def is_masked_mesh(self) -> bool:
        """True for an H x W mesh with cut-out nodes (see csrc/mesh_mask.cu): the stencil kernel applies
        with dis zeroed at the cut-out nodes, plus one pass over the cut-out rows (out = x)."""
        return (self.mesh_kind == "masked" and self.mesh_valid is not None and self.dis is not None
                and self.n_src == self.n_dst and self.flags == _lib.GRAPH_ADD_SELF_LOOPS)

    def dis_padded_masked(self) -> torch.Tensor:
        """Bordered dis of a masked mesh: dis at valid nodes, 0 at cut-out nodes; cached."""
        if getattr(self, "_dis_padded_m", None) is None:
            h, w = self.grid_shape
            d = torch.where(self.mesh_valid, self.dis, torch.zeros_like(self.dis))
            self._dis_padded_m = bordered_dis(d.view(h, w))
        return self._dis_padded_m

    def dis_padded(self) -> torch.Tensor:
        """dis with a one-element zero border, [H + 2, pitch] (element [r+1][c+1] = dis[r][c]),
        pitch sized for any stencil tile width up to 128; cached."""
        if getattr(self, "_dis_padded", None) is None:
            h, w = self.grid_shape
            self._dis_padded = bordered_dis(self.dis.view(h, w))
        return self._dis_padded

    # -- backward graph ----------------------------------------------------------------------
    def transposed(self) -> "GraphCSR":
        """CSR of the transposed graph with the SAME per-edge weights (Appendix A.7)."""
        if self._transposed is None and (self.is_plain_mesh or self.is_masked_mesh):
            # the normalised mesh operator is symmetric (undirected edges, w = dis[s] * dis[d]):
            # A_hat^T = A_hat, and the stencil fast path serves the backward pass too
            self._transposed = self
        if self._transposed is None:
            if self.edge_index is None:
                raise RuntimeError("transposed graph needs the original edge_index")
            self._transposed = _build(self.edge_index, self.n_dst, self.flags | _lib.GRAPH_TRANSPOSE,
                                      dis_in=self.dis)
            self._transposed.grid_shape = self.grid_shape
        return self._transposed

    # -- tile plans ----------------------------------------------------------------------------
    def tile_plan_from(self, order: Optional[torch.Tensor], tile_ptr: torch.Tensor, run_len: int,
                       key: Tuple) -> TilePlan:
        """Plan for an explicit tile layout: ``order`` int32[n_dst] (or None) and ``tile_ptr``
        int32[num_tiles + 1] on this graph's device; memoised under ``key``."""
        if key in self._plans:
            return self._plans[key]
        plan = self._build_plan(order, tile_ptr, run_len)
        self._plans[key] = plan
        return plan

    def tile_plan(self, tile: Optional[Tuple[int, ...]] = None, run_len: Optional[int] = None) -> TilePlan:
        """``tile=(th, tw)`` -> 2-D blocks of the grid (needs ``grid_shape``; runs of tw + 2 rows);
        ``tile=(rows,)`` -> contiguous destination ranges.  Default: (8, 16) blocks on a grid,
        else 128-row ranges with 32-row runs."""
        if tile is None:
            tile = (8, 16) if self.grid_shape is not None else (128,)
        tile = tuple(int(t) for t in tile)
        if run_len is None:
            run_len = tile[1] + 2 if len(tile) == 2 else 32
        key = tile + (run_len,)
        if key in self._plans:
            return self._plans[key]
        L, st = lib(), _stream()
        dev = self.device
        with torch.cuda.device(dev):
            if len(tile) == 2:
                if self.grid_shape is None:
                    raise RuntimeError("2-D tiles need a grid-shaped graph")
                h, w = self.grid_shape
                th, tw = tile
                nt = -(-h // th) * -(-w // tw)
                order = torch.empty(self.n_dst, dtype=torch.int32, device=dev)
                tile_ptr = torch.empty(nt + 1, dtype=torch.int32, device=dev)
                check(L.gwen_grid_tiles(h, w, th, tw, _ptr(order), _ptr(tile_ptr), st),
                      "gwen_grid_tiles")
            else:
                rows = tile[0]
                nt = -(-self.n_dst // rows)
                order = None
                tile_ptr = torch.empty(nt + 1, dtype=torch.int32, device=dev)
                check(L.gwen_uniform_tiles(self.n_dst, rows, _ptr(tile_ptr), st),
                      "gwen_uniform_tiles")
        plan = self._build_plan(order, tile_ptr, run_len)
        self._plans[key] = plan
        return plan

    # -- locality tiles (graphs whose numbering carries no locality) ------------------------------
    LOCALITY_TARGET_ROWS = 200      # mean cell size aimed for (a cell of ~200 nodes stages ~1.4 rows per row on a mesh)
    LOCALITY_MERGE_ROWS, LOCALITY_CAP_ROWS = 200, 256
    LOCALITY_MAX_AMPLIFICATION = 2.5

    def locality_tiles(self, radius: int, rounds: int = 12, merge_rows: Optional[int] = None,
                       cap_rows: Optional[int] = None, deal: Optional[int] = None):
        """``gwen_locality_tiles`` on this graph: ``(order, tile_ptr, cell, depth, status)`` with
        ``status = [cells, tiles, largest cell, unreached nodes]`` read back (one host sync)."""
        merge_rows = merge_rows or self.LOCALITY_MERGE_ROWS
        cap_rows = cap_rows or self.LOCALITY_CAP_ROWS
        L, st, dev, n = lib(), _stream(), self.device, self.n_dst
        if deal is None:   # the grid of the staged kernel: one persistent CTA per SM
            deal = torch.cuda.get_device_properties(dev).multi_processor_count
        with torch.cuda.device(dev):
            need = C.c_size_t()
            check(L.gwen_locality_workspace_bytes(n, C.byref(need)), "gwen_locality_workspace_bytes")
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            order = torch.empty(n, dtype=torch.int32, device=dev)
            tile_ptr = torch.empty(n + 2, dtype=torch.int32, device=dev)
            cell = torch.empty(n, dtype=torch.int32, device=dev)
            depth = torch.empty(n, dtype=torch.int32, device=dev)
            status = torch.zeros(4, dtype=torch.int32, device=dev)
            check(L.gwen_locality_tiles(_ptr(self.rowptr), _ptr(self.src), n, radius, rounds, merge_rows, cap_rows,
                                        deal, _ptr(order), _ptr(tile_ptr), _ptr(cell), _ptr(depth), _ptr(status),
                                        _ptr(ws), need.value, st), "gwen_locality_tiles")
            stat = status.tolist()
            tile_ptr = tile_ptr[:stat[1] + 1].clone()
        return order, tile_ptr, cell, depth, stat

    def locality_plan(self, radius: Optional[int] = None) -> Optional[TilePlan]:
        """Tile plan over locality tiles (csrc/locality.cu), single-row runs copied by the gather producer of
        the tiled kernel; memoised.  ``radius=None`` searches the seed spacing whose mean cell size is nearest
        LOCALITY_TARGET_ROWS (cell size grows with the square of the radius on a surface mesh; at most four
        tries).  Returns None when the graph has no locality to exploit (staged rows per destination row above
        LOCALITY_MAX_AMPLIFICATION: an expander, a dense graph) -- the row kernel serves those."""
        key = ("locality", radius)
        if key in self._plans:
            return self._plans[key]
        plan = None
        if self.n_src == self.n_dst and self.n_dst >= 2 and self.num_messages > 0:
            if radius is not None:
                order, tile_ptr = self.locality_tiles(radius)[:2]
            else:
                tried, r = {}, 8
                while r not in tried and len(tried) < 4:
                    res = self.locality_tiles(r)
                    tried[r] = res
                    mean = self.n_dst / max(1, res[4][0])
                    r = int(min(64, max(1, round(r * (self.LOCALITY_TARGET_ROWS / mean) ** 0.5))))
                best = min(tried, key=lambda q: abs(self.n_dst / max(1, tried[q][4][0]) - self.LOCALITY_TARGET_ROWS))
                order, tile_ptr = tried[best][:2]
                self.locality_radius = best
                del tried
            plan = self._build_plan(order, tile_ptr, 1, flags=_lib.PLAN_GATHER)
            if plan.amplification > self.LOCALITY_MAX_AMPLIFICATION:
                plan = None
        self._plans[key] = plan
        return plan

    def _build_plan(self, order, tile_ptr, run_len: int, flags: int = 0) -> TilePlan:
        L, st = lib(), _stream()
        dev = self.device
        nt = tile_ptr.numel() - 1
        with torch.cuda.device(dev):
            m = self.num_messages
            need = C.c_size_t()
            check(L.gwen_tile_plan_workspace_bytes(self.n_dst, m, nt, C.byref(need)),
                  "gwen_tile_plan_workspace_bytes")
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            run_ptr = torch.empty(nt + 1, dtype=torch.int32, device=dev)
            run_full = torch.empty(m, dtype=torch.int32, device=dev)
            trec = torch.empty((self.n_dst, 4), dtype=torch.int32, device=dev)
            tmsg = torch.empty(m + 2, dtype=torch.int64, device=dev)
            tmsg_base = torch.empty(nt + 1, dtype=torch.int32, device=dev)
            status = torch.zeros(8, dtype=torch.int32, device=dev)
            check(L.gwen_tile_plan_build(_ptr(self.rowptr), _ptr(self.src), _ptr(self.w),
                                         _ptr(order), _ptr(tile_ptr), nt, self.n_dst, m, run_len,
                                         _ptr(run_ptr), _ptr(run_full), _ptr(trec), _ptr(tmsg),
                                         _ptr(tmsg_base), _ptr(status), _ptr(ws), need.value, st),
                  "gwen_tile_plan_build")
            total_runs, max_runs, total_src, max_msgs, max_rows = status.tolist()[:5]  # one-time sync
            run_start = torch.zeros(total_runs + 4, dtype=torch.int32, device=dev)   # 4 readable entries of padding
            run_start[:total_runs].copy_(run_full[:total_runs])
            del run_full, ws
        return TilePlan(order, tile_ptr, run_ptr, run_start, trec, tmsg, tmsg_base, self.n_dst,
                        run_len, max_runs, max_rows, max_msgs, total_runs, total_src, flags)


def _build(edge_index: torch.Tensor, num_nodes: int, flags: int,
           dis_in: Optional[torch.Tensor] = None) -> GraphCSR:
    _require_cuda(edge_index, "edge_index")
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
    ei = edge_index.contiguous()
    e = ei.size(1)
    dev = ei.device
    L = lib()
    with torch.cuda.device(dev):
        st = _stream()
        need = C.c_size_t()
        check(L.gwen_graph_workspace_bytes(num_nodes, e, flags, C.byref(need)),
              "gwen_graph_workspace_bytes")
        cap = e + (num_nodes if flags & _lib.GRAPH_ADD_SELF_LOOPS else 0)
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
        src = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        perm = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
        w = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        dis = dis_in if dis_in is not None else torch.empty(num_nodes, dtype=torch.float32, device=dev)
        status = torch.zeros(2, dtype=torch.int32, device=dev)
        check(L.gwen_graph_build(_ptr(ei), e, num_nodes, flags, _ptr(rowptr), _ptr(src), _ptr(perm),
                                 _ptr(dis), _ptr(w), _ptr(status), _ptr(ws), need.value, st),
              "gwen_graph_build")
        bad, m = status.tolist()  # one-time sync per graph
    if bad:
        raise IndexError("edge_index has %d entries outside [0, %d)" % (bad, num_nodes))
    return GraphCSR(rowptr, src[:m], w[:m], dis, perm[:m], num_nodes, num_nodes, m, flags,
                    edge_index=ei)


def _detect_grid(edge_index: torch.Tensor, num_nodes: int) -> Optional[Tuple[int, int]]:
    """(H, W) if edge_index is exactly grid(H, W) (with or without its self loops), else None."""
    e = edge_index.size(1)
    if num_nodes < 4 or e < 12:
        return None
    head = edge_index[:, :4].tolist()
    if head[0][:3] != [0, 0, 0]:
        return None
    with_loops = head[1][0] == 0
    w = head[1][2] if with_loops else head[1][1]
    if w < 2 or num_nodes % w:
        return None
    h = num_nodes // w
    if h < 2:
        return None
    full = grid_edge_count(h, w)
    if e != (full if with_loops else full - num_nodes):
        return None
    ref = grid(h, w, edge_index.device)
    if not with_loops:
        ref = ref[:, ref[0] != ref[1]]
    return (h, w) if torch.equal(ref, edge_index) else None


def _classify_mesh(g: GraphCSR, h: int, w: int) -> Optional[str]:
    """"plain" / "masked" / None for the CSR ``g`` read as an ``h x w`` mesh (gwen_mesh_mask_detect);
    fills ``mesh_valid`` / ``masked_idx`` for a masked mesh.  One host sync."""
    if g.flags != _lib.GRAPH_ADD_SELF_LOOPS or g.n_src != g.n_dst or h * w != g.n_dst or h < 1 or w < 1:
        return None
    dev = g.device
    with torch.cuda.device(dev):
        valid = torch.empty(g.n_dst, dtype=torch.uint8, device=dev)
        status = torch.zeros(2, dtype=torch.int32, device=dev)
        check(lib().gwen_mesh_mask_detect(_ptr(g.rowptr), _ptr(g.src), g.n_dst, h, w, _ptr(valid), _ptr(status),
                                          _stream()), "gwen_mesh_mask_detect")
        bad, cut = status.tolist()
    if bad:
        return None
    if cut == 0:
        return "plain"
    g.mesh_valid = valid.to(torch.bool)
    g.masked_idx = torch.nonzero(~g.mesh_valid).flatten().to(torch.int32)
    return "masked"


def _detect_mesh(g: GraphCSR) -> Optional[Tuple[int, int]]:
    """(H, W) if ``g`` is an 8-neighbour mesh -- complete or with cut-out nodes, edges in ANY order --
    else None.  The width follows from the largest id difference along an edge (W + 1 as soon as one
    diagonal edge exists, W when only vertical ones do)."""
    ei, n = g.edge_index, g.n_dst
    if ei is None or ei.size(1) == 0 or n < 4 or g.flags != _lib.GRAPH_ADD_SELF_LOOPS:
        return None
    d = int((ei[0] - ei[1]).abs().max().item())
    for w in (d - 1, d):
        if w >= 2 and n % w == 0 and n // w >= 2:
            kind = _classify_mesh(g, n // w, w)
            if kind is not None:
                g.mesh_kind = kind
                return (n // w, w)
    return None


def build_graph(edge_index: torch.Tensor, num_nodes: int, add_self_loops: bool = True,
                improved: bool = False, grid_shape: Optional[Tuple[int, int]] = "auto") -> GraphCSR:
    """Run the K0 preprocessor.  ``grid_shape``: "auto" detects an H x W mesh (PyG ``grid`` order, any other
    edge order, or a mesh with cut-out nodes: ``GraphCSR.mesh_kind``); an explicit ``(H, W)`` is verified
    the same way and kept for 2-D tile plans even when the graph is not a mesh operator."""
    flags = (_lib.GRAPH_ADD_SELF_LOOPS if add_self_loops else 0) | (_lib.GRAPH_IMPROVED if improved else 0)
    g = _build(edge_index, num_nodes, flags)
    if grid_shape == "auto":
        grid_shape = _detect_grid(g.edge_index, num_nodes)
        if grid_shape is not None:
            g.mesh_kind = "plain" if flags == _lib.GRAPH_ADD_SELF_LOOPS else None
        else:
            grid_shape = _detect_mesh(g)
    elif grid_shape is not None:
        g.mesh_kind = _classify_mesh(g, int(grid_shape[0]), int(grid_shape[1]))
    g.grid_shape = grid_shape
    return g


# ---------------------------------------------------------------------------------------------
# cache: GCNConv(cached=False) semantics without recomputing the graph every layer call
# ---------------------------------------------------------------------------------------------
class _CacheEntry:
    __slots__ = ("tensor", "version", "num_nodes", "flags", "graph")


_CACHE: list = []
_CACHE_SIZE = 8


def clear_graph_cache() -> None:
    _CACHE.clear()


def get_graph(edge_index: torch.Tensor, num_nodes: int, add_self_loops: bool = True,
              improved: bool = False) -> GraphCSR:
    """CSR for ``edge_index``, memoised on the tensor's storage + version counter.

    PyG recomputes the normalisation on every call (``cached=False``, the GWEN default).  The
    result only depends on ``edge_index``; the cache keeps a strong reference to the tensor it
    was built from, so a hit means "same live memory, same version" and stays correct when a
    loader hands a different ``edge_index`` every iteration (reference models_gnn.py:359).
    """
    flags = (_lib.GRAPH_ADD_SELF_LOOPS if add_self_loops else 0) | (_lib.GRAPH_IMPROVED if improved else 0)
    for i, ent in enumerate(_CACHE):
        t = ent.tensor
        if (t.data_ptr() == edge_index.data_ptr() and t.shape == edge_index.shape
                and t.stride() == edge_index.stride() and t.device == edge_index.device
                and ent.version == edge_index._version and ent.num_nodes == num_nodes
                and ent.flags == flags):
            if i:
                _CACHE.insert(0, _CACHE.pop(i))
            return ent.graph
    g = build_graph(edge_index, num_nodes, add_self_loops, improved)
    ent = _CacheEntry()
    ent.tensor, ent.version, ent.num_nodes, ent.flags, ent.graph = edge_index, edge_index._version, num_nodes, flags, g
    _CACHE.insert(0, ent)
    del _CACHE[_CACHE_SIZE:]
    return g
