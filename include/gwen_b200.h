/*
 * gwen_b200.h -- C ABI of libgwen_b200.so: the B200 (sm_100a) implementation of the GWEN
 * GCN message-passing hot path.
 *
 * The reference (MeteoSwiss/GWEN) has no FFI for this path: the boundary is the Python class
 * torch_geometric.nn.GCNConv imported at reference src/gwen/models_gnn.py:19, constructed at
 * :118-130 / :172-184 and called as conv(x, edge_index) at :147-149 / :204-206.  Each entry
 * point below names the step of that call it replaces (the
 * PyG-2.3.1 implementation itself is an un-vendored dependency, requirements/environment.yml:552).
 * gwen_b200/nn.py is the ctypes host side that mirrors GCNConv on top of these functions;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Every pointer is a DEVICE pointer unless
 *     the parameter name ends in _host.  The caller (PyTorch's caching allocator) owns every
 *     buffer; the library never allocates device memory that outlives a call.
 *   - every launch function takes the CUDA stream (a cudaStream_t passed as void*); calls are
 *     stream-ordered, never synchronise, and are CUDA-graph capturable.
 *   - return value: GWEN_OK (0) or a negative GWEN_E_* code; the message is available from
 *     gwen_last_error() (thread local).  Nothing throws across the ABI.  There is no CPU
 *     fallback: without a device every launch function returns GWEN_E_CUDA.
 *   - matrices are row-major; "ld" is the row pitch in ELEMENTS; batch strides are in elements.
 */
#ifndef GWEN_B200_H_
#define GWEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWEN_ABI_VERSION 1

enum {
  GWEN_OK = 0,
  GWEN_E_BADARG = -1,    /* null pointer, negative size, index range ... */
  GWEN_E_ALIGN = -2,     /* pointer / pitch not aligned as documented     */
  GWEN_E_DTYPE = -3,     /* unknown dtype code                             */
  GWEN_E_CUDA = -4,      /* CUDA runtime / launch error                    */
  GWEN_E_NOSUPPORT = -5, /* valid request this build cannot serve          */
  GWEN_E_WORKSPACE = -6  /* workspace too small                            */
};

enum { GWEN_F32 = 0, GWEN_BF16 = 1 };

/* flags of gwen_graph_build */
enum {
  GWEN_GRAPH_ADD_SELF_LOOPS = 1u, /* GCNConv(add_self_loops=True), the default GWEN uses      */
  GWEN_GRAPH_IMPROVED = 2u,       /* GCNConv(improved=True): self-loop weight 2               */
  GWEN_GRAPH_TRANSPOSE = 4u       /* build the CSR of the transposed graph (backward pass)    */
};

/* flags of the epilogue argument of aggregate / linear */
enum { GWEN_EPI_NONE = 0, GWEN_EPI_RELU = 1 };

int gwen_version(void);
/* The persistent kernels (tiled aggregation, tensor-core GEMM) launch one CTA per SM.  A caller
 * that overlaps them with small kernels on another stream (halo packs, NCCL send/recv) reserves
 * n SMs for those: subsequent persistent launches use (SM count - n) CTAs.  Returns the previous
 * value; process-wide. */
int gwen_set_sm_reserve(int n);
/* Last error message of the calling thread ("" if none).  Pointer stays valid until the next
 * failing call on this thread. */
const char* gwen_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * K0  graph preprocessor.  Replaces, once per edge_index instead of once per layer call,
 * add_remaining_self_loops + ones + scatter(deg) + pow(-0.5) + dis[row]*ew*dis[col]
 *  and produces the destination-sorted CSR the
 * aggregation kernels walk (Appendix B.3).
 *
 * edge_index: int64 [2, E] row-major, row 0 = source, row 1 = destination (PyG flow
 *             source_to_target).  With GWEN_GRAPH_TRANSPOSE the roles are swapped.
 * Outputs (capacity Ecap = E + N when ADD_SELF_LOOPS else E):
 *   rowptr int32[N+1]   segment offsets per destination; rowptr[N] = E' (messages)
 *   src    int32[Ecap]  source node of each message, destination-major, STABLE in edge_index'
 *                       order (surviving input edges in list order, the self loop last)
 *   perm   int64[Ecap]  (nullable) position of the message in edge_index' (after A.2)
 *   dis    fp32[N]      deg^-1/2, fp64-computed and rounded once (0 for deg 0)
 *   w      fp32[Ecap]   (nullable) dis[src] * fill * dis[dst] per message, CSR order
 *   status int32[2]     [0] = number of out-of-range indices seen (must be 0), [1] = E'
 * ws: scratch of at least gwen_graph_workspace_bytes(...) bytes, 256-byte aligned.
 * Requires N < 2^31 - 1 and E + N < 2^31.
 */
int gwen_graph_workspace_bytes(int64_t num_nodes, int64_t num_edges, uint32_t flags,
                               size_t* bytes_out_host);
int gwen_graph_build(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                     uint32_t flags, int32_t* rowptr, int32_t* src, int64_t* perm, float* dis,
                     float* w, int32_t* status, void* ws, size_t ws_bytes, void* stream);

/* Graph builders whose edge ORDER is part of the parity contract.
 * gwen_complete_edges: erdos_renyi_graph(N, edge_prob=1) as called at reference
 *   src/gwen/utils.py:176 -> all (i, j), i != j, sorted by (row, col); out int64 [2, N(N-1)].
 * gwen_grid_edges: PyG grid(H, W) 8-neighbour mesh incl. self loops, node id = r*W + c, sorted
 *   by (row, col); out int64 [2, gwen_grid_edge_count(H, W)]. */
int64_t gwen_grid_edge_count(int64_t height, int64_t width);
int gwen_grid_edges(int64_t height, int64_t width, int64_t* edge_index_out, void* stream);
int gwen_complete_edges(int64_t num_nodes, int64_t* edge_index_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  deterministic segment-reduce aggregation.  Replaces index_select + message multiply +
 * scatter_add_ + bias + relu:
 *   out[b, i, :] = epi( sum_{s in [rowptr[i], rowptr[i+1])} w[s] * x[b, src[s], :]  + bias )
 * Summation runs in CSR order with fp32 accumulation; for fp32 the products and adds are
 * formed unfused (mul.rn then add.rn) so the result equals the CPU scatter_add_ order bitwise.
 * x: [B, n_src, F] (pitch ldx, batch stride x_bstride), out: [B, n_dst, F] (ldo, o_bstride).
 * bias: fp32[F] or NULL.  dtype: GWEN_F32 / GWEN_BF16 (x and out share it).
 * order: int32[n_dst] or NULL -- the order in which destination rows are assigned to warps
 *        (a locality hint, e.g. the 2-D tile order of gwen_grid_tiles); results do not depend on it.
 */
int gwen_aggregate_fwd(const int32_t* rowptr, const int32_t* src, const float* w,
                       const int32_t* order, const void* x, void* out, int64_t batch,
                       int64_t n_dst, int64_t n_src, int64_t feat, int64_t ldx, int64_t x_bstride,
                       int64_t ldo, int64_t o_bstride, int dtype, const float* bias, int epilogue,
                       void* stream);

/* Tiled variant (mesh-like graphs): destination rows are processed in tiles; the source rows a
 * tile needs are covered by a few RUNS of run_len consecutive row ids, and each run's feature
 * slab is brought into shared memory by ONE TMA tensor copy (cp.async.bulk.tensor, mbarrier
 * completion), double-buffered against the reduction.  A source row slab then crosses L2->SM
 * once per tile instead of once per message.  The plan is built once per graph from the CSR by
 * gwen_tile_plan_build:
 *   order     int32[n_dst]        (nullable = identity) destination processing order
 *   tile_ptr  int32[num_tiles+1]  tile t covers positions tile_ptr[t] .. tile_ptr[t+1] of order
 *   run_ptr   int32[num_tiles+1]  offsets into run_start
 *   run_start int32[...]          first source row of each run (greedy cover of the tile's
 *                                 ascending distinct sources with windows of run_len rows)
 *   trec      int4[n_dst]         per position p of the order: {destination row, offset of its
 *                                 first message inside the tile's staged message block, degree, 0}
 *   tmsg      uint64[E'+2]        messages in processing order (tile-contiguous): low 32 bits =
 *                                 staged row of the source inside its tile (run index * run_len
 *                                 + offset), high 32 bits = the fp32 weight
 *   tmsg_base int32[num_tiles+1]  first message of each tile in tmsg
 * All arrays are caller-allocated device memory (run_start capacity = E' is always enough;
 * trec and tmsg 16-byte aligned).  The result is bitwise identical to gwen_aggregate_fwd. */
#define GWEN_PLAN_GATHER 1 /* gwen_tile_plan.reserved: run_len == 1 and run_start has >= 4 readable entries
                             past its last run -> source rows are copied one by one with 16-byte cp.async by
                             the producer warp (tiles of gwen_locality_tiles) instead of one TMA box per run */
typedef struct gwen_tile_plan {
  int32_t num_tiles;
  int32_t run_len;         /* rows per run = TMA box height (<= 256)                        */
  int32_t max_tile_runs;   /* largest run_ptr[t+1] - run_ptr[t]; sizes the data stages      */
  int32_t max_tile_rows;   /* most destination rows in a tile                               */
  int32_t max_tile_msgs;   /* most staged messages of a tile (status[3])                    */
  int32_t reserved;        /* flags: GWEN_PLAN_GATHER                                        */
  int64_t n_dst;
  const int32_t* tile_ptr;  /* device */
  const int32_t* run_ptr;   /* device */
  const int32_t* run_start; /* device */
  const void* trec;         /* device, int4[n_dst] */
  const uint64_t* tmsg;     /* device */
  const int32_t* tmsg_base; /* device */
} gwen_tile_plan;

int gwen_tile_plan_workspace_bytes(int64_t n_dst, int64_t num_messages, int64_t num_tiles,
                                   size_t* bytes_out_host);
/* status int32[8]: [0] = total runs = run_ptr[num_tiles], [1] = max runs of any tile,
 *   [2] = total distinct (tile, source) pairs (staging efficiency = [2] / ([0] * run_len)),
 *   [3] = max staged messages of a tile, [4] = max destination rows of a tile. */
int gwen_tile_plan_build(const int32_t* rowptr, const int32_t* src, const float* w,
                         const int32_t* order, const int32_t* tile_ptr, int64_t num_tiles,
                         int64_t n_dst, int64_t num_messages, int32_t run_len, int32_t* run_ptr,
                         int32_t* run_start, void* trec, uint64_t* tmsg, int32_t* tmsg_base,
                         int32_t* status, void* ws, size_t ws_bytes, void* stream);
/* Tile layouts.  uniform: identity order, tile_rows destinations per tile (tile_ptr only).
 * grid: 2-D th x tw blocks of an H x W grid graph (node id r*W + c), tiles row-major; writes
 * order[H*W] and tile_ptr[ceil(H/th)*ceil(W/tw) + 1]. */
int gwen_uniform_tiles(int64_t n_dst, int32_t tile_rows, int32_t* tile_ptr_out, void* stream);
int gwen_grid_tiles(int64_t height, int64_t width, int32_t th, int32_t tw, int32_t* order_out,
                    int32_t* tile_ptr_out, void* stream);
/* Locality tiles (csrc/locality.cu) for a graph whose node numbering carries no locality (a mesh with permuted
 * ids, an unstructured grid in file order): compact patches of the graph found from the CSR alone -- seeds = a
 * maximal independent set of the radius-th power of the graph (`rounds` Luby rounds with hashed priorities),
 * cells = their Voronoi regions (multi-source BFS), order = nodes sorted by (cell, BFS depth, id); cells above
 * cap_rows are cut into equal chunks, consecutive cells are packed into one tile while it stays within merge_rows.
 * Replaces nothing in the reference (PyG's scatter_add has no locality pass); it is the north_star's "one-time
 * CSR/dst-sorted graph preprocessor" extended so the TMA/shared-memory staged kernel serves such graphs.
 *   order    int32[n]      out: processing order (a permutation of 0..n-1)
 *   tile_ptr int32[n + 2]  out: tile t covers positions tile_ptr[t] .. tile_ptr[t+1]; status[1] tiles are written
 *   cell, depth int32[n]   out: cell index and BFS depth of every node
 *   status   int32[4]      [0] cells, [1] tiles, [2] largest cell, [3] nodes no seed reached (asymmetric edge lists;
 *                          they share one extra cell)
 * deal > 0: the tiles are renumbered by size -- descending row count, dealt in snake order over `deal` CTAs -- because
 * the staged kernel gives tile t to CTA t mod grid (pass the grid it will run with, the SM count); 0 keeps cell order.
 * Deterministic; no host synchronisation.  Feed order / tile_ptr to gwen_tile_plan_build with run_len = 1 and set
 * GWEN_PLAN_GATHER in the plan. */
int gwen_locality_workspace_bytes(int64_t n, size_t* bytes_out_host);
int gwen_locality_tiles(const int32_t* rowptr, const int32_t* src, int64_t n, int32_t radius, int32_t rounds,
                        int32_t merge_rows, int32_t cap_rows, int32_t deal, int32_t* order, int32_t* tile_ptr, int32_t* cell,
                        int32_t* depth, int32_t* status, void* ws, size_t ws_bytes, void* stream);
/* slab_elems: feature columns staged per work item: 8, 16 or 32 sixteen-byte chunks
 * (0 = widest that fits); two stages of max_tile_runs * run_len * slab bytes must fit in the
 * 227 KB of shared memory, else GWEN_E_NOSUPPORT (use gwen_aggregate_fwd). */
/* tile_begin / tile_count: process only tiles [tile_begin, tile_begin + tile_count) of the plan
 * (0, 0 = all) -- lets a partitioned run launch interior tiles while the halo is in flight. */
int gwen_aggregate_tiled_fwd(const gwen_tile_plan* plan_host, const void* x, void* out,
                             int64_t batch, int64_t n_src, int64_t feat, int64_t ldx,
                             int64_t x_bstride, int64_t ldo, int64_t o_bstride, int dtype,
                             const float* bias, int epilogue, int32_t slab_elems,
                             int32_t tile_begin, int32_t tile_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1s  mesh fast path of the aggregation (same reference steps as K1: rows 7-11).
 * For the exact 8-neighbour H x W mesh with GCN normalisation (what gwen_grid_edges builds and
 * K0 normalises; the host layer checks this before choosing it):
 *   out[b, r, c, :] = epi( dis[r', c] * sum_{dr, dc in {-1,0,1}} dis[r'+dr, c+dc] * x[b, r'+dr, c+dc, :] + bias )
 * with r' = r + row_off and nodes outside [0, hs) x [0, w) contributing 0.  Evaluated as a
 * separable box filter (column sums in registers), sources staged by one 4-D TMA box per tile.
 *   x   : [B, hs, w, F] (node pitch ldx, batch stride x_bstride)      out : [B, hd, w, F]
 *   dis_padded : fp32 [dis_rows >= round_up(hd, 8) + row_off + 2, dis_pitch] (checked: the kernel
 *                bulk-copies 10 dis rows per 8-row tile, GWEN_E_BADARG if they would not fit) with a one-element ZERO
 *                border (and zero padding): element [r+1][c+1] = dis[r][c]; dis_pitch % 4 == 0 and >= (ceil(w / tile_w) - 1) * tile_w +
 *                round_up(tile_w + 2, 4); 16-byte aligned (rows are fetched with bulk copies)
 *   row_off    : source row of destination row 0 (0 for a whole mesh; 1 for a row band whose x
 *                carries one halo row above and below: hs = hd + 2)
 * Agrees with gwen_aggregate_fwd to fp32 rounding (different summation order), deterministic,
 * independent of tiling and partitioning. */
int gwen_grid_stencil_fwd(const void* x, void* out, const float* dis_padded, int64_t dis_pitch,
                          int64_t dis_rows, int64_t batch, int64_t hs, int64_t hd, int64_t w, int64_t row_off,
                          int64_t feat, int64_t ldx, int64_t x_bstride, int64_t ldo,
                          int64_t o_bstride, int dtype, const float* bias, int epilogue,
                          int32_t slab_elems, int32_t tile_w, void* stream);

/* K1s + halo exchange in ONE kernel (multi-GPU row bands, SURVEY.md section 8e: no reference
 * counterpart -- the reference runs identical replicas).  x is this rank's band
 * [B, hd + 2, w, feat]: local row 0 and row hd + 1 are the halo rows, rows 1 .. hd are owned.  The
 * launch (1) stores its epoch into the neighbours' flag words (peer memory over NVLink), (2) one
 * extra warp per CTA waits for the neighbours' flags and copies its share of their boundary rows
 * from PEER memory into the local halo rows while the other warps aggregate the interior tile
 * rows, (3) the tile rows that read a halo row are processed last, after all CTAs have published
 * their share.  Results are bitwise equal to gwen_grid_stencil_fwd on exchanged halos.
 * All ranks must launch the same sequence of peer calls (epochs are counted on the device, so a
 * CUDA-graph replay works).  The neighbour reads this rank's rows 1 and hd of x during ITS launch
 * of the same epoch: do not overwrite x before the next peer launch (ping-pong two buffers).
 * up_row / down_row : the up neighbour's LAST owned row / the down neighbour's FIRST owned row
 *                     (pointers into the neighbour's x, mapped into this process), NULL at the
 *                     mesh edge (the local halo row must then be zero)
 * up/down_bstride   : the neighbour's batch stride in elements
 * up_flag/down_flag : address (peer memory) of the neighbour's ctl[1] / ctl[0]
 * ctl               : >= 8 zero-initialised uint32 in LOCAL device memory, owned by the protocol:
 *                     [0] epoch announced by the up neighbour, [1] by the down neighbour,
 *                     [2] epochs completed, [3] CTAs that published their halo share, [4] CTAs done,
 *                     [5] STICKY ERROR word: every wait of the protocol is bounded (GWEN_PEER_TIMEOUT_MS,
 *                     default 10 s); on expiry bit 0 (a neighbour never announced this epoch: it launched a
 *                     different sequence of peer calls, or died) or bit 1 (a CTA of this grid never published
 *                     its halo share) is set and the kernel finishes with undefined halo-dependent rows
 *                     instead of hanging the GPU.  The host reads it when it next synchronises.
 * The launch is cooperative (cudaLaunchAttributeCooperative) with the grid clamped to the occupancy
 * limit, so all CTAs are co-resident or the launch fails with GWEN_E_CUDA. */
typedef struct gwen_halo_peers {
  const void* up_row;
  const void* down_row;
  int64_t up_bstride, down_bstride;
  uint32_t* up_flag;
  uint32_t* down_flag;
  uint32_t* ctl;
} gwen_halo_peers;
int gwen_grid_stencil_peer_fwd(void* x, void* out, const float* dis_padded, int64_t dis_pitch,
                               int64_t dis_rows, int64_t batch, int64_t hd, int64_t w, int64_t feat,
                               int64_t x_bstride, int64_t ldo, int64_t o_bstride, int dtype,
                               const float* bias, int epilogue, int32_t slab_elems, int32_t tile_w,
                               const gwen_halo_peers* peers, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1 + K2 fused (mesh graphs, bf16): y = epi( (A_hat x) W^T + bias ) in ONE kernel -- the whole
 * GCNConv.forward for the layers that aggregate before they project
 * (in_channels < out_channels).  CUDA-core warps run the mesh stencil of gwen_grid_stencil_fwd on
 * TMA-staged source boxes and write the aggregated rows into shared memory as the tcgen05 A operand;
 * the projection is the CTA-pair GEMM of gwen_linear_fwd.  The aggregated intermediate is rounded to
 * bf16 exactly as the two-kernel path stores it, so results match that path.
 *   x : bf16 [B, h*w, k_in] contiguous      weight : bf16 [n_out, k_in]      y : bf16 [B, h*w, n_out]
 *   dis_padded / dis_pitch / dis_rows : as for gwen_grid_stencil_fwd (dis_rows >= round_up(h, 8) + 2)
 *   k_in in {64, 128, .., 512}, n_out % 128 == 0 (else GWEN_E_NOSUPPORT: use the two kernels)
 *   pre_bias fp32 [k_in] (nullable) / pre_epilogue (GWEN_EPI_RELU or 0): applied to the AGGREGATED row before it
 *   is projected, a = epi_pre(A_hat x + pre_bias) rounded to bf16 as gwen_grid_stencil_fwd would store it --
 *   the previous layer's bias / ReLU when that layer aggregates last and hands over its un-aggregated
 *   projection:  y = epi( epi_pre(A_hat x + pre_bias) W^T + bias ). */
int gwen_gcn_fused_fwd(const void* x, const void* weight, void* y, const float* dis_padded,
                       int64_t dis_pitch, int64_t dis_rows, int64_t batch, int64_t h, int64_t w, int64_t k_in,
                       int64_t n_out, int dtype, const float* bias, int epilogue, const float* pre_bias,
                       int pre_epilogue, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Two projections back to back (bf16): y = epi2( epi1(x W1^T + bias1) W2^T + bias2 ) in ONE kernel.
 * Where a GCNConv that projects last (in < out: conv1, upconv4; models_gnn.py:147,204) is followed by one
 * that projects first (in > out: conv2, upconv5; torch_geometric GCNConv.forward `x = self.lin(x)`), the wide
 * hidden tensor between them never reaches HBM: per 256-column chunk the first product's accumulator is turned
 * into bf16 (rounded exactly as gwen_linear_fwd stores it) in shared memory and consumed as the A operand of
 * the second product.  Results equal gwen_linear_fwd(epi1) followed by gwen_linear_fwd(epi2).
 *   x : bf16 [m, k1] (row pitch ldx)   w1 : bf16 [n1, k1]   w2 : bf16 [n2, n1]   y : bf16 [m, n2] (pitch ldy)
 *   bias1 fp32 [n1] / bias2 fp32 [n2] nullable; epilogue1 / epilogue2: GWEN_EPI_RELU or 0
 *   k1 in {64, .., 512}, n1 % 256 == 0, n2 % 64 == 0 and (n2 <= 256 or n2 % 256 == 0), m >= 256:
 *   gwen_linear_b2b_supported returns 1, else 0 (then gwen_linear_b2b_fwd returns GWEN_E_NOSUPPORT). */
int gwen_linear_b2b_supported(int64_t m, int64_t k1, int64_t n1, int64_t n2, int dtype);
int gwen_linear_b2b_fwd(const void* x, const void* w1, const float* bias1, int epilogue1, const void* w2,
                        const float* bias2, int epilogue2, void* y, int64_t m, int64_t k1, int64_t n1,
                        int64_t n2, int64_t ldx, int64_t ldy, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  dense projection.  Replaces Linear(bias=False) inside GCNConv (F.linear -> cuBLAS,
 * SURVEY.md table 2.3 row 6) and, through the epilogue, the bias add and ReLU when the
 * projection runs after the aggregation:
 *   y[m, n] = epi( sum_k x[m, k] * weight[n, k] + bias[n] )        x:[M,K]  weight:[Nout,K]
 * GWEN_F32: fp32 in / fp32 accumulate / fp32 out.
 * GWEN_BF16: bf16 in, fp32 accumulate (tcgen05 tensor cores, TMEM accumulators), bf16 out.
 * gwen_linear_bwd_* are the two GEMMs of the backward pass (Appendix A.7):
 *   dgrad: dx[m, k] = sum_n dy[m, n] * weight[n, k]
 *   wgrad: dw[n, k] = sum_m dy[m, n] * x[m, k]      (fp32 output, deterministic split order)
 */
int gwen_linear_fwd(const void* x, const void* weight, void* y, int64_t m, int64_t k,
                    int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int dtype,
                    const float* bias, int epilogue, void* stream);
/* gwen_linear_fwd with a scratch buffer: fp32 problems with m >= 4096, k % 4 == 0, n_out % 64 == 0 and
 * dense rows then run on the tensor cores as three kind::tf32 products per K step (operands split once
 * into hi / lo TF32 parts in the workspace, fp32 accumulation: fp32-level accuracy, ~1e-6 relative);
 * everything else (and a NULL / too small workspace, or GWEN_FP32_SIMT set) is gwen_linear_fwd. */
int gwen_linear_fwd_workspace_bytes(int64_t m, int64_t k, int64_t n_out, int dtype, size_t* bytes_out_host);
int gwen_linear_fwd_ws(const void* x, const void* weight, void* y, int64_t m, int64_t k, int64_t n_out,
                       int64_t ldx, int64_t ldw, int64_t ldy, int dtype, const float* bias, int epilogue,
                       void* ws, size_t ws_bytes, void* stream);
int gwen_linear_bwd_data(const void* dy, const void* weight, void* dx, int64_t m, int64_t k,
                         int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, int dtype,
                         void* stream);
/* The same with the PREVIOUS layer's ReLU backward folded into the epilogue: dx = (dy W) * (mask > 0), mask = a bf16
 * tensor of dx's shape (the layer input x, i.e. the previous layer's ReLU output; replaces the `torch.relu` backward
 * between two GCNConv calls, reference models_gnn.py:147-149,204-206).  bf16 on the tcgen05 CTA-pair kernel only:
 * GWEN_E_NOSUPPORT otherwise, and gwen_linear_bwd_data + gwen_relu_bias_bwd give the same bits. */
int gwen_linear_bwd_data_masked(const void* dy, const void* weight, void* dx, const void* mask, int64_t m, int64_t k,
                                int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, int64_t ldmask, int dtype,
                                void* stream);
/* ... over `batch` slices in one launch (batch-strided dy / dx / mask: e.g. dx = the owned rows of a band buffer). */
int gwen_linear_batched_bwd_data_masked(const void* dy, const void* weight, void* dx, const void* mask, int64_t batch,
                                        int64_t m, int64_t k, int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx,
                                        int64_t ldmask, int64_t dy_bstride, int64_t dx_bstride, int64_t mask_bstride,
                                        int dtype, void* stream);
/* Batched forms: `batch` independent row blocks x[b] = x + b * x_bstride ([m, k], row pitch ldx) against the
 * SAME weight, results to y[b] = y + b * y_bstride -- e.g. the owned rows of a band buffer per ensemble
 * member.  bf16 problems the CTA-pair kernel takes run as ONE launch (the batch index is the third TMA
 * coordinate); everything else is a loop over gwen_linear_fwd / gwen_linear_bwd_data. */
int gwen_linear_batched_fwd(const void* x, const void* weight, void* y, int64_t batch, int64_t m, int64_t k,
                            int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int64_t x_bstride,
                            int64_t y_bstride, int dtype, const float* bias, int epilogue, void* stream);
int gwen_linear_batched_bwd_data(const void* dy, const void* weight, void* dx, int64_t batch, int64_t m,
                                 int64_t k, int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx,
                                 int64_t dy_bstride, int64_t dx_bstride, int dtype, void* stream);
/* gwen_linear_bwd_data with scratch: fp32 problems (m >= 4096, n_out % 4 == 0, k % 64 == 0, dense rows)
 * run on the tensor cores through the 3xTF32 forward kernel on the split transpose of W. */
int gwen_linear_bwd_data_workspace_bytes(int64_t m, int64_t k, int64_t n_out, int dtype,
                                         size_t* bytes_out_host);
int gwen_linear_bwd_data_ws(const void* dy, const void* weight, void* dx, int64_t m, int64_t k,
                            int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, int dtype, void* ws,
                            size_t ws_bytes, void* stream);
int gwen_linear_bwd_weight(const void* dy, const void* x, float* dw, int64_t m, int64_t k,
                           int64_t n_out, int64_t lddy, int64_t ldx, int64_t lddw, int dtype,
                           void* ws, size_t ws_bytes, void* stream);
/* The same AND db[n] = sum_m dy[m, n] (fp32) out of the same pass over dy: one more tcgen05.mma per 16 rows against a
 * tile of ones in the CTAs of the first K tile.  With the ReLU mask applied by gwen_linear_bwd_data_masked this leaves
 * no separate pass over the gradient of a layer that aggregates first.  bf16 tcgen05 kernel only (GWEN_E_NOSUPPORT
 * otherwise); workspace from gwen_linear_bwd_weight_bias_workspace_bytes. */
int gwen_linear_bwd_weight_bias_workspace_bytes(int64_t m, int64_t k, int64_t n_out, size_t* bytes_out_host);
int gwen_linear_bwd_weight_bias(const void* dy, const void* x, float* dw, float* db, int64_t m, int64_t k,
                                int64_t n_out, int64_t lddy, int64_t ldx, int64_t lddw, int dtype, void* ws,
                                size_t ws_bytes, void* stream);
int gwen_linear_bwd_weight_workspace_bytes(int64_t m, int64_t k, int64_t n_out,
                                           size_t* bytes_out_host);

/* Elementwise helpers of the backward pass: relu mask (dy *= y > 0) and bias gradient
 * (db[f] = sum_rows dy[:, f], fixed-order tree, fp32 out). */
int gwen_relu_bwd(const void* y, void* dy, int64_t rows, int64_t feat, int64_t ldy, int64_t lddy,
                  int dtype, void* stream);
int gwen_bias_grad(const void* dy, float* db, int64_t rows, int64_t feat, int64_t lddy, int dtype,
                   void* ws, size_t ws_bytes, void* stream);
int gwen_bias_grad_workspace_bytes(int64_t rows, int64_t feat, size_t* bytes_out_host);
/* Both in one pass over contiguous [rows, feat] tensors: dz = dy * (y > 0) (skipped when y is
 * NULL; dz may alias dy) and db[f] = sum_rows dz[:, f] (skipped when db is NULL; workspace as for
 * gwen_bias_grad).  Needs feat = (16 / element size) * d with d a divisor of 256, else
 * GWEN_E_NOSUPPORT (use the two calls above). */
int gwen_relu_bias_bwd(const void* y, const void* dy, void* dz, float* db, int64_t rows,
                       int64_t feat, int dtype, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Caller-side loss of the reference training loop: loss_func(output, target, mask) =
 * L1Loss()(output[mask], target[mask]) (src/gwen/models_gnn.py:261-265, called at :370 with
 * target = the input features).  Replaces nonzero + two boolean-mask gathers + l1_loss (+ the host
 * synchronisation the data-dependent size of output[mask] forces) by one pass for the value and
 * one for the gradient; nothing here synchronises the host.
 *   y, target : [batch, n, feat] contiguous;  mask : uint8 [n] (torch.bool), shared by the batch
 *   loss_and_scale : float[2] device: [0] = mean |y - target| over the masked rows (NaN if the mask
 *                    is empty, like torch), [1] = 1 / (batch * count * feat), read by the backward
 *   dloss : device scalar gradient of the loss (NULL = 1);  dy = sign(y - target) * mask * dloss * scale
 *   count_override : NULL, or a device float holding the number of masked nodes to divide by instead
 *                    of this call's own count -- a rank of a partitioned mesh passes the GLOBAL count, so
 *                    its value is its share  local_sum / (batch * global_count * feat)  of the loss over
 *                    the whole mesh (0, not NaN, for a band without masked nodes) */
int gwen_masked_l1_workspace_bytes(int64_t n, size_t* bytes_out_host);
int gwen_masked_l1_fwd(const void* y, const void* target, const uint8_t* mask, int64_t batch, int64_t n,
                       int64_t feat, int dtype, const float* count_override, float* loss_and_scale,
                       void* ws, size_t ws_bytes, void* stream);
int gwen_masked_l1_bwd(const void* y, const void* target, const uint8_t* mask,
                       const float* loss_and_scale, const float* dloss, int64_t batch, int64_t n,
                       int64_t feat, int dtype, void* dy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Halo exchange helpers for the row-band mesh partition (no reference counterpart; SURVEY.md
 * section 8(e)): gather / scatter whole feature rows by index so the send and receive buffers
 * are contiguous for ncclSend / ncclRecv.
 *   pack:   buf[b, j, :] = x[b, idx[j], :]        unpack: x[b, idx[j], :] = buf[b, j, :]
 */
int gwen_rows_gather(const void* x, const int32_t* idx, void* buf, int64_t batch, int64_t n_idx,
                     int64_t feat, int64_t ldx, int64_t x_bstride, int dtype, void* stream);
int gwen_rows_scatter(const void* buf, const int32_t* idx, void* x, int64_t batch, int64_t n_idx,
                      int64_t feat, int64_t ldx, int64_t x_bstride, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Full-neighbour subgraph extraction: what
 * NeighborLoader(data, num_neighbors=[-1] * hops, batch_size, shuffle=False) does on the CPU before every
 * model call of the reference loops (src/gwen/models_gnn.py:351-356, :434-439) through torch_geometric 2.3.1
 * -> torch_sparse 0.6.17 neighbor_sample<replace=false, directed=true> (not in the reference tree; contract
 * restated in oracle/neighbor_oracle.py).
 *   colptr int32[n + 1], row int32[e] : CSC of the graph (in-neighbours of node w = row[colptr[w] .. colptr[w+1]));
 *       this is gwen_graph_build with flags = 0 (destination-sorted, stable in edge_index order -- for the
 *       (row, col)-sorted edge lists the reference builds that is PyG's to_csc order)
 *   perm int64[e] (nullable) : CSC slot -> position in the caller's edge_index (gwen_graph_build's perm)
 *   seeds int64[n_seeds] : the batch's input nodes, distinct
 *   node_out int64[n]   : seeds first, then newly reached nodes in discovery order (n_id)
 *   row_out / col_out int64[e] : relabelled edge_index of the batch (row = source, col = destination), one edge
 *       per (frontier node, in-neighbour) in visiting order;  edge_out int64[e] : e_id (perm[slot], or the slot)
 *   counts_out int32[3] device : {nodes, edges, bad seeds (out of range or repeated)}
 * Nothing synchronises the host; the caller reads counts_out when it needs the sizes. */
int gwen_neighbor_workspace_bytes(int64_t n, int64_t e, size_t* bytes_out_host);
int gwen_neighbor_sample_full(const int32_t* colptr, const int32_t* row, const int64_t* perm, int64_t n,
                              int64_t e, const int64_t* seeds, int64_t n_seeds, int32_t hops,
                              int64_t* node_out, int64_t* row_out, int64_t* col_out, int64_t* edge_out,
                              int32_t* counts_out, void* ws, size_t ws_bytes, void* stream);
/* The same for the complete directed graph on n nodes sorted by (row, col) (erdos_renyi_graph(n, 1),
 * src/gwen/utils.py:176) and seeds seed_start .. seed_start + n_seeds - 1, hops >= 2: closed form, ONE launch.
 * node_out int64[n]; edge_index_out int64[2, n (n - 1)] (row 0 = source, row 1 = destination);
 * eid_out int64[n (n - 1)] or NULL. */
int gwen_neighbor_complete(int64_t n, int64_t seed_start, int64_t n_seeds, int64_t* node_out,
                           int64_t* edge_index_out, int64_t* eid_out, void* stream);
/* dst[j, :] = src[idx[j], :] for rows of row_bytes bytes (x[n_id], target_mask[n_id]: the node-attribute
 * filtering of PyG's filter_data). */
int gwen_gather_rows_bytes(const void* src, const int64_t* idx, void* dst, int64_t rows, int64_t row_bytes,
                           int64_t src_rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Masked meshes (an H x W 8-neighbour mesh with a set of nodes cut out: land/sea masks, non-rectangular
 * domains).  A cut-out node keeps only the self loop GCN normalisation adds (out = x); a valid node's
 * degree counts its valid neighbours.  With dis zeroed at the cut-out nodes the aggregation of the valid
 * nodes is gwen_grid_stencil_fwd unchanged, and gwen_rows_self_fwd rewrites the cut-out rows.
 * gwen_mesh_mask_detect: is the CSR of gwen_graph_build (flags = GWEN_GRAPH_ADD_SELF_LOOPS) such a graph
 * for this h x w?  valid_out uint8[n] (1 = node has an in-edge besides its self loop);
 * status int32[2] device: [0] = nodes whose in-neighbour set is NOT exactly {valid mesh neighbours}
 * (0 means yes), [1] = number of cut-out nodes. */
int gwen_mesh_mask_detect(const int32_t* rowptr, const int32_t* src, int64_t n, int64_t h, int64_t w,
                          uint8_t* valid_out, int32_t* status, void* stream);
/* out[b, idx[j], :] = epi(x[b, idx[j], :] + bias) for the n_idx listed rows (fp32 add, then ReLU). */
int gwen_rows_self_fwd(const void* x, void* out, const int32_t* idx, int64_t n_idx, int64_t batch,
                       int64_t feat, int64_t ldx, int64_t x_bstride, int64_t ldo, int64_t o_bstride,
                       int dtype, const float* bias, int epilogue, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step of the reference training loop (configs_train_gnn.optimizer.step(), src/gwen/models_gnn.py:373,
 * with torch.optim.Adam(model.parameters(), lr = config["lr"] * 10), src/gwen/train_gnn.py:111): torch.optim.Adam's
 * update (amsgrad = False) for `count` fp32 tensors in one launch per 32 tensors.
 *   params / grads / exp_avg / exp_avg_sq : HOST arrays of `count` DEVICE pointers (fp32, numel[i] elements each)
 *   step : 1-based step count of this update (bias corrections 1 - beta^step are formed on the host in fp64) */
int gwen_adam_step(int32_t count, void* const* params, const void* const* grads, void* const* exp_avg,
                   void* const* exp_avg_sq, const int64_t* numel, int64_t step, double lr, double beta1,
                   double beta2, double eps, double weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWEN_B200_H_ */
