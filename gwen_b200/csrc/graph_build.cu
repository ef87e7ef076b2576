// K0: one-time graph preprocessor (integer work, HBM-bound) and the two graph builders whose
// edge order is part of the parity contract.  See include/gwen_b200.h for the contract and
// SURVEY.md Appendix A.2-A.4 / B for the semantics restated here.
//
// Pipeline (all on the caller's stream, no host sync):
//   prepare : key = destination (or N for dropped self loops / bad indices), val = position in
//             the concatenated list [input edges | appended self loops]; keep flags
//   scan    : exclusive sum of keep flags -> position of every surviving edge in edge_index'
//   sort    : STABLE LSB radix sort by key over ceil(log2(N+1)) bits (cub::DeviceRadixSort)
//   rowptr  : segment boundaries from the sorted keys (handles empty segments)
//   fill    : src / perm per CSR slot;  dis : fp64 1/sqrt(deg) rounded once;  w : per-slot weight
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gwen {
namespace {

constexpr int kThreads = 256;

struct GraphWs {
  int64_t total;  // E + (self ? N : 0)
  size_t off_keys_in, off_keys_out, off_vals_in, off_vals_out, off_pos, off_loop, off_cub, cub_bytes, bytes;
};

int key_bits(int64_t n) {
  int b = 1;
  while ((int64_t(1) << b) <= n) ++b;
  return b;
}

cudaError_t plan_ws(int64_t n, int64_t e, uint32_t flags, GraphWs* p) {
  const bool self = flags & GWEN_GRAPH_ADD_SELF_LOOPS;
  p->total = e + (self ? n : 0);
  size_t t = static_cast<size_t>(p->total > 0 ? p->total : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p->off_keys_in = take(t * 4);
  p->off_keys_out = take(t * 4);
  p->off_vals_in = take(t * 4);
  p->off_vals_out = take(t * 4);
  p->off_pos = take((static_cast<size_t>(e) + 1) * 4);
  p->off_loop = take(static_cast<size_t>(n > 0 ? n : 1));  // had_loop[n]: node had a self loop in the input
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t err = cub::DeviceRadixSort::SortPairs(
      nullptr, sort_bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
      static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
      static_cast<int>(t), 0, key_bits(n));
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(e + 1));
  if (err != cudaSuccess) return err;
  p->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  p->off_cub = take(p->cub_bytes);
  p->bytes = off;
  return cudaSuccess;
}

__global__ void k_prepare(const int64_t* __restrict__ ei, int64_t e, int64_t n, int64_t total,
                          bool self, bool transpose, int32_t* __restrict__ keys,
                          int32_t* __restrict__ vals, int32_t* __restrict__ pos,
                          unsigned char* __restrict__ had_loop, int32_t* __restrict__ status) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i == 0) pos[e] = 0;
  if (i >= total) return;
  if (i < e) {
    int64_t r = ei[i], c = ei[e + i];
    bool ok = r >= 0 && r < n && c >= 0 && c < n;
    if (!ok) atomicAdd(&status[0], 1);
    bool keep = ok && (!self || r != c);
    if (ok && self && r == c) had_loop[r] = 1;  // add_remaining_self_loops keeps an existing loop's weight (1)
    int64_t dst = transpose ? r : c;
    keys[i] = keep ? static_cast<int32_t>(dst) : static_cast<int32_t>(n);
    pos[i] = keep ? 1 : 0;
  } else {
    keys[i] = static_cast<int32_t>(i - e);
  }
  vals[i] = static_cast<int32_t>(i);
}

// rowptr[d] = first slot whose key >= d.  One thread per slot writes the (possibly empty) run
// of boundaries between its predecessor's key and its own.
__global__ void k_rowptr(const int32_t* __restrict__ keys, int64_t total, int64_t n,
                         int32_t* __restrict__ rowptr) {
  int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (s >= total) return;
  int32_t k = keys[s];
  int32_t kp = s ? keys[s - 1] : -1;
  for (int64_t d = int64_t(kp) + 1; d <= k && d <= n; ++d) rowptr[d] = static_cast<int32_t>(s);
  if (s == total - 1 && k < n)
    for (int64_t d = int64_t(k) + 1; d <= n; ++d) rowptr[d] = static_cast<int32_t>(total);
}

__global__ void k_fill(const int64_t* __restrict__ ei, int64_t e, int64_t n, int64_t total,
                       bool transpose, const int32_t* __restrict__ keys,
                       const int32_t* __restrict__ vals, const int32_t* __restrict__ pos,
                       int32_t* __restrict__ src, int64_t* __restrict__ perm) {
  int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (s >= total) return;
  if (keys[s] >= n) return;
  int32_t v = vals[s];
  if (v < e) {
    src[s] = static_cast<int32_t>(transpose ? ei[e + v] : ei[v]);
    if (perm) perm[s] = pos[v];
  } else {
    src[s] = static_cast<int32_t>(v - e);
    if (perm) perm[s] = int64_t(pos[e]) + (v - e);
  }
}

// deg = sum of edge weights into i: 1 per surviving edge; the appended self loop weighs `loop_fill`
// (2 for improved=True) unless the input already held a self loop of i, whose weight 1 is kept
// (PyG add_remaining_self_loops: loop_attr[existing] = edge_attr[existing]).
__global__ void k_dis(const int32_t* __restrict__ rowptr, int64_t n, int extra_deg,
                      const unsigned char* __restrict__ had_loop, float* __restrict__ dis,
                      int32_t* __restrict__ status) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i == 0) status[1] = rowptr[n];
  if (i >= n) return;
  int32_t cnt = rowptr[i + 1] - rowptr[i];
  double deg = double(cnt > 0 ? cnt + (had_loop[i] ? 0 : extra_deg) : 0);
  dis[i] = deg > 0.0 ? static_cast<float>(1.0 / sqrt(deg)) : 0.0f;
}

// w[s] = (dis[src] * fill) * dis[dst]: the two fp32 multiplies of `dis[row] * ew * dis[col]`.
__global__ void k_weights(const int32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                          const int32_t* __restrict__ src, const float* __restrict__ dis,
                          int64_t e, int64_t n, int64_t total, float loop_fill,
                          const unsigned char* __restrict__ had_loop, float* __restrict__ w) {
  int64_t s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (s >= total) return;
  int32_t k = keys[s];
  if (k >= n) return;
  float fill = (vals[s] >= e && !had_loop[k]) ? loop_fill : 1.0f;
  w[s] = __fmul_rn(__fmul_rn(dis[src[s]], fill), dis[k]);
}

// ---- builders -------------------------------------------------------------------------------
__host__ __device__ inline int64_t cnt_upto(int64_t c, int64_t w) {
  // sum_{c' < c} (1 + [c' > 0] + [c' < w-1])
  int64_t a = c > 0 ? c - 1 : 0;
  int64_t b = c < w - 1 ? c : w - 1;
  if (b < 0) b = 0;
  return c + a + b;
}

__global__ void k_grid_edges(int64_t h, int64_t w, int64_t e, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= h * w) return;
  int64_t r = i / w, c = i % w;
  int64_t per_row = cnt_upto(w, w);
  int64_t nc = 1 + (c > 0) + (c < w - 1);
  int64_t off = cnt_upto(r, h) * per_row + (1 + (r > 0) + (r < h - 1)) * cnt_upto(c, w);
  (void)nc;
  for (int dr = -1; dr <= 1; ++dr) {
    int64_t rr = r + dr;
    if (rr < 0 || rr >= h) continue;
    for (int dc = -1; dc <= 1; ++dc) {
      int64_t cc = c + dc;
      if (cc < 0 || cc >= w) continue;
      out[off] = i;
      out[e + off] = rr * w + cc;
      ++off;
    }
  }
}

__global__ void k_complete_edges(int64_t n, int64_t e, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= e) return;
  int64_t r = i / (n - 1), j = i % (n - 1);
  out[i] = r;
  out[e + i] = j + (j >= r);
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_graph_workspace_bytes(int64_t n, int64_t e, uint32_t flags, size_t* out) {
  GWEN_CHECK_ARG(out != nullptr, "bytes_out_host is null");
  GWEN_CHECK_ARG(n >= 0 && e >= 0, "negative size");
  GWEN_CHECK_ARG(n < INT32_MAX - 1 && e + n < INT32_MAX, "graph too large for int32 CSR");
  GraphWs p;
  GWEN_CUDA(plan_ws(n, e, flags, &p));
  *out = p.bytes;
  return GWEN_OK;
}

extern "C" int gwen_graph_build(const int64_t* edge_index, int64_t e, int64_t n, uint32_t flags,
                                int32_t* rowptr, int32_t* src, int64_t* perm, float* dis,
                                float* w, int32_t* status, void* ws, size_t ws_bytes,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWEN_CHECK_ARG(n >= 0 && e >= 0, "negative size");
  GWEN_CHECK_ARG(n < INT32_MAX - 1 && e + n < INT32_MAX, "graph too large for int32 CSR");
  GWEN_CHECK_ARG(rowptr && dis && status && ws, "null output pointer");
  GWEN_CHECK_ARG(e == 0 || edge_index, "edge_index is null");
  const bool self = flags & GWEN_GRAPH_ADD_SELF_LOOPS;
  const bool improved = flags & GWEN_GRAPH_IMPROVED;
  const bool transpose = flags & GWEN_GRAPH_TRANSPOSE;
  GraphWs p;
  GWEN_CUDA(plan_ws(n, e, flags, &p));
  if (ws_bytes < p.bytes)
    return set_err(GWEN_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes, p.bytes);
  GWEN_CHECK_ARG(p.total == 0 || src, "src is null");
  char* base = static_cast<char*>(ws);
  int32_t* keys_in = reinterpret_cast<int32_t*>(base + p.off_keys_in);
  int32_t* keys_out = reinterpret_cast<int32_t*>(base + p.off_keys_out);
  int32_t* vals_in = reinterpret_cast<int32_t*>(base + p.off_vals_in);
  int32_t* vals_out = reinterpret_cast<int32_t*>(base + p.off_vals_out);
  int32_t* pos = reinterpret_cast<int32_t*>(base + p.off_pos);
  unsigned char* had_loop = reinterpret_cast<unsigned char*>(base + p.off_loop);
  void* cub_ws = base + p.off_cub;
  GWEN_CUDA(cudaMemsetAsync(had_loop, 0, static_cast<size_t>(n > 0 ? n : 1), stream));

  GWEN_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), stream));
  GWEN_CUDA(cudaMemsetAsync(rowptr, 0, (n + 1) * sizeof(int32_t), stream));
  const int64_t total = p.total;
  const unsigned grid_t = static_cast<unsigned>(ceil_div(total > 0 ? total : 1, kThreads));
  k_prepare<<<grid_t, kThreads, 0, stream>>>(edge_index, e, n, total, self, transpose, keys_in,
                                             vals_in, pos, had_loop, status);
  GWEN_LAUNCH_CHECK("k_prepare");
  size_t cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, pos, pos, static_cast<int>(e + 1),
                                          stream));
  if (total > 0) {
    cub_bytes = p.cub_bytes;
    GWEN_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, vals_in,
                                              vals_out, static_cast<int>(total), 0, key_bits(n),
                                              stream));
    k_rowptr<<<grid_t, kThreads, 0, stream>>>(keys_out, total, n, rowptr);
    GWEN_LAUNCH_CHECK("k_rowptr");
    k_fill<<<grid_t, kThreads, 0, stream>>>(edge_index, e, n, total, transpose, keys_out,
                                            vals_out, pos, src, perm);
    GWEN_LAUNCH_CHECK("k_fill");
  }
  const unsigned grid_n = static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kThreads));
  if (!transpose) {
    k_dis<<<grid_n, kThreads, 0, stream>>>(rowptr, n, (self && improved) ? 1 : 0, had_loop, dis, status);
    GWEN_LAUNCH_CHECK("k_dis");
  } else {
    // dis is an INPUT here (the forward graph's in-degree normalisation); only record E'.
    GWEN_CUDA(cudaMemcpyAsync(status + 1, rowptr + n, sizeof(int32_t), cudaMemcpyDeviceToDevice,
                              stream));
  }
  if (w && total > 0) {
    k_weights<<<grid_t, kThreads, 0, stream>>>(keys_out, vals_out, src, dis, e, n, total,
                                               improved ? 2.0f : 1.0f, had_loop, w);
    GWEN_LAUNCH_CHECK("k_weights");
  }
  return GWEN_OK;
}

extern "C" int64_t gwen_grid_edge_count(int64_t h, int64_t w) {
  if (h <= 0 || w <= 0) return 0;
  return cnt_upto(h, h) * cnt_upto(w, w);
}

extern "C" int gwen_grid_edges(int64_t h, int64_t w, int64_t* out, void* stream_) {
  GWEN_CHECK_ARG(h > 0 && w > 0 && out, "bad grid arguments");
  int64_t e = gwen_grid_edge_count(h, w);
  k_grid_edges<<<static_cast<unsigned>(ceil_div(h * w, kThreads)), kThreads, 0,
                 static_cast<cudaStream_t>(stream_)>>>(h, w, e, out);
  GWEN_LAUNCH_CHECK("k_grid_edges");
  return GWEN_OK;
}

extern "C" int gwen_complete_edges(int64_t n, int64_t* out, void* stream_) {
  GWEN_CHECK_ARG(n >= 0, "negative node count");
  int64_t e = n * (n - 1);
  if (e <= 0) return GWEN_OK;
  GWEN_CHECK_ARG(out, "edge_index_out is null");
  k_complete_edges<<<static_cast<unsigned>(ceil_div(e, kThreads)), kThreads, 0,
                     static_cast<cudaStream_t>(stream_)>>>(n, e, out);
  GWEN_LAUNCH_CHECK("k_complete_edges");
  return GWEN_OK;
}
