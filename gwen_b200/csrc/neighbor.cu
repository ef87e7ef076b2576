// Full-neighbour subgraph extraction: the step BEFORE the layer stack on every iteration of the
// reference loops (NeighborLoader(data, num_neighbors=[-1, -1], batch_size, shuffle=False),
// src/gwen/models_gnn.py:351-356 and :434-439).  The reference runs it on the CPU through
// torch_geometric 2.3.1 -> torch_sparse 0.6.17 `neighbor_sample` (neither is in the reference tree;
// the contract restated here is written out in oracle/neighbor_oracle.py):
//
//   node list  : the seeds in batch order, then every newly reached node in DISCOVERY order, hop by hop --
//                frontier nodes in list order, each frontier node's in-neighbours in CSC order;
//   edges      : one per (frontier node, in-neighbour) in exactly that visiting order,
//                row = local id of the in-neighbour (source), col = local id of the frontier node,
//                edge = CSC slot (the caller maps it through the CSC permutation to e_id);
//   every node is a frontier node at most once (the hop after it was discovered).
//
// A hash map + sequential loop on the host becomes, per hop, a fixed sequence of data-parallel passes:
//   degrees of the frontier -> exclusive scan (edge slots) -> atomicMin of the slot index per not yet
//   numbered source (order-independent, hence deterministic) -> the slots that won are the discoveries:
//   flag + exclusive scan = discovery rank -> number the new nodes -> emit (row, col, edge).
// All counts stay on the device (kernels are grid-stride over device-side counters, scans run over the
// host-known upper bounds), so a batch is extracted without a host round trip until the caller reads the
// two totals.
//
// Complete graphs (what GWEN actually builds, utils.py:176) have a closed form and take ONE launch:
// every batch is the whole graph renumbered [seeds | the other nodes ascending].
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gwen {
namespace {

constexpr int kNbThreads = 256;

struct NbCounters {   // device-side state of one extraction
  int32_t begin, end;       // frontier = samples[begin, end)
  int32_t n_edges;          // edges emitted by the finished hops
  int32_t hop_edges;        // edge slots of the current hop
  int32_t hop_new;          // nodes discovered in the current hop
  int32_t bad_seed;         // seeds outside [0, n) or repeated
};

struct NbWs {
  size_t off_local, off_first, off_deg, off_offs, off_flag, off_rank, off_cnt, off_cub, cub_bytes, bytes;
};

cudaError_t plan_nb(int64_t n, int64_t e, NbWs* p) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1), ee = static_cast<size_t>(e > 0 ? e : 1);
  p->off_local = take(nn * 4);
  p->off_first = take(nn * 4);
  p->off_deg = take((nn + 1) * 4);
  p->off_offs = take((nn + 1) * 4);
  p->off_flag = take((ee + 1) * 4);
  p->off_rank = take((ee + 1) * 4);
  p->off_cnt = take(sizeof(NbCounters));
  size_t b1 = 0, b2 = 0;
  cudaError_t err = cub::DeviceScan::ExclusiveSum(nullptr, b1, static_cast<int32_t*>(nullptr),
                                                  static_cast<int32_t*>(nullptr), static_cast<int>(nn + 1));
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::ExclusiveSum(nullptr, b2, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(ee + 1));
  if (err != cudaSuccess) return err;
  p->cub_bytes = b1 > b2 ? b1 : b2;
  p->off_cub = take(p->cub_bytes);
  p->bytes = off;
  return cudaSuccess;
}

__global__ void k_nb_seeds(const int64_t* __restrict__ seeds, int64_t n_seeds, int64_t n,
                           int32_t* __restrict__ local_of, int64_t* __restrict__ node_out,
                           NbCounters* __restrict__ c) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i == 0) {
    c->begin = 0;
    c->end = static_cast<int32_t>(n_seeds);
    c->n_edges = 0;
    c->hop_edges = 0;
    c->hop_new = 0;
  }
  if (i >= n_seeds) return;
  const int64_t v = seeds[i];
  if (v < 0 || v >= n) {
    atomicAdd(&c->bad_seed, 1);
    return;
  }
  node_out[i] = v;
  if (atomicCAS(&local_of[v], -1, static_cast<int32_t>(i)) != -1) atomicAdd(&c->bad_seed, 1);
}

// deg[j] = in-degree of frontier node j (0 beyond the frontier, up to the scan bound)
__global__ void k_nb_deg(const int32_t* __restrict__ colptr, const int64_t* __restrict__ node_out,
                         const NbCounters* __restrict__ c, int64_t bound, int32_t* __restrict__ deg) {
  const int32_t cnt = c->end - c->begin;
  for (int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; j < bound;
       j += int64_t(gridDim.x) * blockDim.x) {
    int32_t d = 0;
    if (j < cnt) {
      const int64_t w = node_out[c->begin + j];
      d = colptr[w + 1] - colptr[w];
    }
    deg[j] = d;
  }
}

// frontier index of edge slot s: the last j with offs[j] <= s
__device__ __forceinline__ int32_t slot_owner(const int32_t* __restrict__ offs, int32_t cnt, int32_t s) {
  int32_t lo = 0, hi = cnt;  // offs[cnt] = total > s
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (offs[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_nb_mark(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                          const int64_t* __restrict__ node_out, const int32_t* __restrict__ offs,
                          const int32_t* __restrict__ local_of, int32_t* __restrict__ first_pos,
                          NbCounters* __restrict__ c) {
  const int32_t cnt = c->end - c->begin;
  const int32_t total = offs[cnt];
  if (blockIdx.x == 0 && threadIdx.x == 0) c->hop_edges = total;
  for (int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; s < total;
       s += int64_t(gridDim.x) * blockDim.x) {
    const int32_t j = slot_owner(offs, cnt, static_cast<int32_t>(s));
    const int64_t w = node_out[c->begin + j];
    const int32_t v = row[colptr[w] + (static_cast<int32_t>(s) - offs[j])];
    if (local_of[v] < 0) atomicMin(&first_pos[v], static_cast<int32_t>(s));
  }
}

// flag[s] = 1 where slot s is the FIRST slot (in visiting order) that reaches a not yet numbered node
__global__ void k_nb_flag(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                          const int64_t* __restrict__ node_out, const int32_t* __restrict__ offs,
                          const int32_t* __restrict__ local_of, const int32_t* __restrict__ first_pos,
                          const NbCounters* __restrict__ c, int64_t bound, int32_t* __restrict__ flag) {
  const int32_t cnt = c->end - c->begin;
  const int32_t total = c->hop_edges;
  for (int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; s < bound;
       s += int64_t(gridDim.x) * blockDim.x) {
    int32_t f = 0;
    if (s < total) {
      const int32_t j = slot_owner(offs, cnt, static_cast<int32_t>(s));
      const int64_t w = node_out[c->begin + j];
      const int32_t v = row[colptr[w] + (static_cast<int32_t>(s) - offs[j])];
      f = (local_of[v] < 0 && first_pos[v] == static_cast<int32_t>(s)) ? 1 : 0;
    }
    flag[s] = f;
  }
}

__global__ void k_nb_assign(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                            const int32_t* __restrict__ offs, const int32_t* __restrict__ flag,
                            const int32_t* __restrict__ rank, int32_t* __restrict__ local_of,
                            int64_t* __restrict__ node_out, NbCounters* __restrict__ c) {
  const int32_t cnt = c->end - c->begin;
  const int32_t total = c->hop_edges;
  if (blockIdx.x == 0 && threadIdx.x == 0) c->hop_new = rank[total];   // exclusive scan: rank[total] = #flags
  for (int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; s < total;
       s += int64_t(gridDim.x) * blockDim.x) {
    if (!flag[s]) continue;
    const int32_t j = slot_owner(offs, cnt, static_cast<int32_t>(s));
    const int64_t w = node_out[c->begin + j];
    const int32_t v = row[colptr[w] + (static_cast<int32_t>(s) - offs[j])];
    const int32_t id = c->end + rank[s];
    node_out[id] = v;
    local_of[v] = id;
  }
}

__global__ void k_nb_emit(const int32_t* __restrict__ colptr, const int32_t* __restrict__ row,
                          const int64_t* __restrict__ node_out, const int32_t* __restrict__ offs,
                          const int32_t* __restrict__ local_of, const NbCounters* __restrict__ c,
                          const int64_t* __restrict__ perm, int64_t* __restrict__ row_out,
                          int64_t* __restrict__ col_out, int64_t* __restrict__ edge_out) {
  const int32_t cnt = c->end - c->begin;
  const int32_t total = c->hop_edges;
  const int64_t base = c->n_edges;
  for (int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; s < total;
       s += int64_t(gridDim.x) * blockDim.x) {
    const int32_t j = slot_owner(offs, cnt, static_cast<int32_t>(s));
    const int64_t w = node_out[c->begin + j];
    const int32_t slot = colptr[w] + (static_cast<int32_t>(s) - offs[j]);
    row_out[base + s] = local_of[row[slot]];
    col_out[base + s] = c->begin + j;
    edge_out[base + s] = perm ? perm[slot] : int64_t(slot);   // e_id: position in the caller's edge_index
  }
}

__global__ void k_nb_advance(NbCounters* __restrict__ c, int32_t* __restrict__ counts_out) {
  c->begin = c->end;
  c->end += c->hop_new;
  c->n_edges += c->hop_edges;
  c->hop_edges = 0;
  c->hop_new = 0;
  counts_out[0] = c->end;       // nodes so far
  counts_out[1] = c->n_edges;   // edges so far
  counts_out[2] = c->bad_seed;
}

// ---- complete graph, closed form ---------------------------------------------------------------
// local -> global: l < bs ? s0 + l : (l - bs < s0 ? l - bs : l)           (seeds first, then the rest ascending)
// global -> local: s0 <= v < s0 + bs ? v - s0 : (v < s0 ? bs + v : v)
__global__ void k_nb_complete(int64_t n, int64_t s0, int64_t bs, int64_t* __restrict__ node_out,
                              int64_t* __restrict__ ei_out, int64_t* __restrict__ eid_out) {
  const int64_t e = n * (n - 1);
  const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (k < n) node_out[k] = k < bs ? s0 + k : (k - bs < s0 ? k - bs : k);
  if (k >= e) return;
  // visiting order: frontier node of local id i (hop 1: the seeds, hop 2: everybody else), its in-neighbours
  // ascending by global id (the CSC of the (row, col)-sorted complete graph), self excluded
  const int64_t i = k / (n - 1), t = k % (n - 1);
  const int64_t w = i < bs ? s0 + i : (i - bs < s0 ? i - bs : i);
  const int64_t v = t + (t >= w);
  const int64_t lv = (v >= s0 && v < s0 + bs) ? v - s0 : (v < s0 ? bs + v : v);
  ei_out[k] = lv;         // row: source
  ei_out[e + k] = i;      // col: the frontier (destination) node
  // e_id: position in the ORIGINAL (row, col)-sorted edge list of the edge (v -> w)
  if (eid_out) eid_out[k] = v * (n - 1) + (w - (w > v));
}

// dst[j, :] = src[idx[j], :] for rows of row_bytes bytes (16-byte vector path when everything is aligned)
__global__ void k_gather_rows_bytes(const unsigned char* __restrict__ src, const int64_t* __restrict__ idx,
                                    unsigned char* __restrict__ dst, int64_t rows, int64_t row_bytes,
                                    int64_t src_rows, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t j = warp; j < rows; j += nwarps) {
    const int64_t r = idx[j];
    if (r < 0 || r >= src_rows) continue;
    const unsigned char* s = src + r * row_bytes;
    unsigned char* d = dst + j * row_bytes;
    if (vec) {
      const uint4* s4 = reinterpret_cast<const uint4*>(s);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      for (int64_t q = lane; q < row_bytes / 16; q += 32) d4[q] = s4[q];
    } else {
      for (int64_t q = lane; q < row_bytes; q += 32) d[q] = s[q];
    }
  }
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_neighbor_workspace_bytes(int64_t n, int64_t e, size_t* out) {
  GWEN_CHECK_ARG(out != nullptr, "bytes_out_host is null");
  GWEN_CHECK_ARG(n >= 0 && e >= 0 && n < INT32_MAX - 1 && e < INT32_MAX - 1, "graph too large for int32 slots");
  NbWs p;
  GWEN_CUDA(plan_nb(n, e, &p));
  *out = p.bytes;
  return GWEN_OK;
}

extern "C" int gwen_neighbor_sample_full(const int32_t* colptr, const int32_t* row, const int64_t* perm,
                                         int64_t n, int64_t e, const int64_t* seeds, int64_t n_seeds, int32_t hops,
                                         int64_t* node_out, int64_t* row_out, int64_t* col_out,
                                         int64_t* edge_out, int32_t* counts_out, void* ws, size_t ws_bytes,
                                         void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GWEN_CHECK_ARG(n >= 0 && e >= 0 && n < INT32_MAX - 1 && e < INT32_MAX - 1, "graph too large for int32 slots");
  GWEN_CHECK_ARG(n_seeds >= 0 && n_seeds <= n && hops >= 0 && hops <= 64, "bad seed count / hop count");
  GWEN_CHECK_ARG(colptr && counts_out && ws && node_out, "null pointer");
  GWEN_CHECK_ARG(e == 0 || (row && row_out && col_out && edge_out), "null edge pointer");
  GWEN_CHECK_ARG(n_seeds == 0 || seeds, "seeds is null");
  NbWs p;
  GWEN_CUDA(plan_nb(n, e, &p));
  if (ws_bytes < p.bytes) return set_err(GWEN_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes, p.bytes);
  char* base = static_cast<char*>(ws);
  int32_t* local_of = reinterpret_cast<int32_t*>(base + p.off_local);
  int32_t* first_pos = reinterpret_cast<int32_t*>(base + p.off_first);
  int32_t* deg = reinterpret_cast<int32_t*>(base + p.off_deg);
  int32_t* offs = reinterpret_cast<int32_t*>(base + p.off_offs);
  int32_t* flag = reinterpret_cast<int32_t*>(base + p.off_flag);
  int32_t* rank = reinterpret_cast<int32_t*>(base + p.off_rank);
  NbCounters* cnt = reinterpret_cast<NbCounters*>(base + p.off_cnt);
  void* cub_ws = base + p.off_cub;
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  GWEN_CUDA(cudaMemsetAsync(local_of, 0xFF, nn * 4, st));       // -1
  GWEN_CUDA(cudaMemsetAsync(first_pos, 0x7F, nn * 4, st));      // 0x7F7F7F7F > any slot index
  GWEN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(NbCounters), st));
  const int grid_cap = std::max(1, sm_count() * 8);
  auto grid_for = [&](int64_t items) {
    return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, kNbThreads), grid_cap)));
  };
  k_nb_seeds<<<grid_for(std::max<int64_t>(n_seeds, 1)), kNbThreads, 0, st>>>(seeds, n_seeds, n, local_of, node_out, cnt);
  GWEN_LAUNCH_CHECK("k_nb_seeds");
  for (int h = 0; h < hops && e > 0; ++h) {
    k_nb_deg<<<grid_for(n + 1), kNbThreads, 0, st>>>(colptr, node_out, cnt, n + 1, deg);
    GWEN_LAUNCH_CHECK("k_nb_deg");
    size_t cb = p.cub_bytes;
    GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cb, deg, offs, static_cast<int>(n + 1), st));
    k_nb_mark<<<grid_for(e), kNbThreads, 0, st>>>(colptr, row, node_out, offs, local_of, first_pos, cnt);
    GWEN_LAUNCH_CHECK("k_nb_mark");
    k_nb_flag<<<grid_for(e + 1), kNbThreads, 0, st>>>(colptr, row, node_out, offs, local_of, first_pos, cnt, e + 1, flag);
    GWEN_LAUNCH_CHECK("k_nb_flag");
    cb = p.cub_bytes;
    GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cb, flag, rank, static_cast<int>(e + 1), st));
    k_nb_assign<<<grid_for(e), kNbThreads, 0, st>>>(colptr, row, offs, flag, rank, local_of, node_out, cnt);
    GWEN_LAUNCH_CHECK("k_nb_assign");
    k_nb_emit<<<grid_for(e), kNbThreads, 0, st>>>(colptr, row, node_out, offs, local_of, cnt, perm, row_out, col_out, edge_out);
    GWEN_LAUNCH_CHECK("k_nb_emit");
    k_nb_advance<<<1, 1, 0, st>>>(cnt, counts_out);
    GWEN_LAUNCH_CHECK("k_nb_advance");
  }
  if (hops == 0 || e == 0) {
    k_nb_advance<<<1, 1, 0, st>>>(cnt, counts_out);
    GWEN_LAUNCH_CHECK("k_nb_advance");
  }
  return GWEN_OK;
}

extern "C" int gwen_neighbor_complete(int64_t n, int64_t seed_start, int64_t n_seeds, int64_t* node_out,
                                      int64_t* edge_index_out, int64_t* eid_out, void* stream_) {
  GWEN_CHECK_ARG(n >= 1 && seed_start >= 0 && n_seeds >= 1 && seed_start + n_seeds <= n, "bad seed range");
  GWEN_CHECK_ARG(node_out && (n == 1 || edge_index_out), "null pointer");
  const int64_t items = std::max(n, n * (n - 1));
  k_nb_complete<<<static_cast<unsigned>(ceil_div(items, kNbThreads)), kNbThreads, 0,
                  static_cast<cudaStream_t>(stream_)>>>(n, seed_start, n_seeds, node_out, edge_index_out, eid_out);
  GWEN_LAUNCH_CHECK("k_nb_complete");
  return GWEN_OK;
}

extern "C" int gwen_gather_rows_bytes(const void* src, const int64_t* idx, void* dst, int64_t rows,
                                      int64_t row_bytes, int64_t src_rows, void* stream_) {
  GWEN_CHECK_ARG(rows >= 0 && row_bytes >= 0 && src_rows >= 0, "negative size");
  if (rows == 0 || row_bytes == 0) return GWEN_OK;
  GWEN_CHECK_ARG(src && idx && dst, "null pointer");
  const int vec = (row_bytes % 16 == 0) && aligned16(src) && aligned16(dst);
  const int64_t blocks = std::min<int64_t>(ceil_div(rows * 32, kNbThreads), int64_t(sm_count()) * 16);
  k_gather_rows_bytes<<<static_cast<unsigned>(std::max<int64_t>(blocks, 1)), kNbThreads, 0,
                        static_cast<cudaStream_t>(stream_)>>>(static_cast<const unsigned char*>(src), idx,
                                                              static_cast<unsigned char*>(dst), rows, row_bytes,
                                                              src_rows, vec);
  GWEN_LAUNCH_CHECK("k_gather_rows_bytes");
  return GWEN_OK;
}
