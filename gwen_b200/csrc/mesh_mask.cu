// Masked meshes: an H x W 8-neighbour mesh from which a set of nodes has been cut out (land/sea masks,
// non-rectangular model domains).  The cut-out nodes stay in the node list but keep no edge except the
// self loop GCN normalisation adds, so for them  out = x  (deg = 1), and a valid node's degree counts its
// VALID neighbours only.  With  dis_m[i] = dis[i]  for valid nodes and 0 for cut-out ones,
//     out[i] = dis_m[i] * sum_{3x3 window} dis_m[j] x[j]            (valid i: exactly the CSR sum)
// is the separable stencil of k_grid_stencil with another dis array, and the cut-out rows are
// rewritten by one small pass  out[i] = epi(x[i] + bias).  So a masked mesh keeps the mesh fast path
// (one TMA box per tile, no index traffic) instead of dropping to the general CSR kernel.
//
// gwen_mesh_mask_detect decides whether a CSR (K0 output, self loops added) IS such a graph for a
// given H x W: every in-edge joins 8-neighbours, no duplicates, and a node's in-neighbours are exactly
// its valid mesh neighbours, where valid = "has at least one in-edge besides the self loop".
#include "common.cuh"

namespace gwen {
namespace {

constexpr int kMmThreads = 256;

__global__ void k_mm_valid(const int32_t* __restrict__ rowptr, int64_t n, uint8_t* __restrict__ valid,
                           int32_t* __restrict__ status) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int v = (rowptr[i + 1] - rowptr[i]) > 1;
  valid[i] = static_cast<uint8_t>(v);
  if (!v) atomicAdd(&status[1], 1);
}

__global__ void k_mm_check(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int64_t n,
                           int64_t h, int64_t w, const uint8_t* __restrict__ valid,
                           int32_t* __restrict__ status) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int64_t r = i / w, c = i % w;
  uint32_t seen = 0;
  int bad = 0, loops = 0;
  for (int32_t s = rowptr[i]; s < rowptr[i + 1]; ++s) {
    const int64_t j = src[s];
    if (j == i) {            // the appended self loop (K0 drops input self loops and appends one)
      ++loops;
      continue;
    }
    const int64_t dr = j / w - r, dc = j % w - c;
    if (dr < -1 || dr > 1 || dc < -1 || dc > 1 || !valid[j]) {
      bad = 1;
      continue;
    }
    const uint32_t bit = 1u << ((dr + 1) * 3 + (dc + 1));
    if (seen & bit) bad = 1;   // duplicate edge
    seen |= bit;
  }
  if (loops != 1) bad = 1;
  // the in-neighbours must be ALL valid mesh neighbours (a valid node's are; a cut-out node has none)
  uint32_t want = 0;
  if (valid[i]) {
    for (int dr = -1; dr <= 1; ++dr)
      for (int dc = -1; dc <= 1; ++dc) {
        if (!dr && !dc) continue;
        const int64_t rr = r + dr, cc = c + dc;
        if (rr < 0 || rr >= h || cc < 0 || cc >= w) continue;
        if (valid[rr * w + cc]) want |= 1u << ((dr + 1) * 3 + (dc + 1));
      }
  }
  if (want != seen) bad = 1;
  if (bad) atomicAdd(&status[0], 1);
}

// out[b, idx[j], :] = epi(x[b, idx[j], :] + bias): the self-loop-only rows of a masked mesh
template <typename T>
__global__ void k_rows_self(const T* __restrict__ x, T* __restrict__ out, const int32_t* __restrict__ idx,
                            int64_t n_idx, int64_t batch, int64_t feat, int64_t ldx, int64_t x_bstride,
                            int64_t ldo, int64_t o_bstride, const float* __restrict__ bias, int relu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t u = warp; u < n_idx * batch; u += nwarps) {
    const int64_t b = u / n_idx, i = idx[u % n_idx];
    const T* xr = x + b * x_bstride + i * ldx;
    T* orow = out + b * o_bstride + i * ldo;
    for (int64_t k = lane; k < feat; k += 32) {
      float v = to_f32(xr[k]);
      if (bias) v = __fadd_rn(v, bias[k]);
      if (relu) v = fmaxf(v, 0.0f);
      orow[k] = from_f32<T>(v);
    }
  }
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_mesh_mask_detect(const int32_t* rowptr, const int32_t* src, int64_t n, int64_t h,
                                     int64_t w, uint8_t* valid_out, int32_t* status, void* stream_) {
  GWEN_CHECK_ARG(n >= 0 && h > 0 && w > 0 && h * w == n, "h * w must equal the node count");
  GWEN_CHECK_ARG(rowptr && valid_out && status && (n == 0 || src), "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  GWEN_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  if (n == 0) return GWEN_OK;
  const unsigned grid = static_cast<unsigned>(ceil_div(n, kMmThreads));
  k_mm_valid<<<grid, kMmThreads, 0, st>>>(rowptr, n, valid_out, status);
  GWEN_LAUNCH_CHECK("k_mm_valid");
  k_mm_check<<<grid, kMmThreads, 0, st>>>(rowptr, src, n, h, w, valid_out, status);
  GWEN_LAUNCH_CHECK("k_mm_check");
  return GWEN_OK;
}

extern "C" int gwen_rows_self_fwd(const void* x, void* out, const int32_t* idx, int64_t n_idx, int64_t batch,
                                  int64_t feat, int64_t ldx, int64_t x_bstride, int64_t ldo,
                                  int64_t o_bstride, int dtype, const float* bias, int epilogue,
                                  void* stream_) {
  GWEN_CHECK_ARG(n_idx >= 0 && batch >= 0 && feat >= 0, "negative size");
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  if (n_idx * batch * feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && out && idx, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int64_t blocks = std::min<int64_t>(ceil_div(n_idx * batch * 32, kMmThreads), int64_t(sm_count()) * 16);
  const unsigned grid = static_cast<unsigned>(std::max<int64_t>(blocks, 1));
  const int relu = (epilogue & GWEN_EPI_RELU) ? 1 : 0;
  if (dtype == GWEN_F32)
    k_rows_self<float><<<grid, kMmThreads, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(out), idx,
                                                    n_idx, batch, feat, ldx, x_bstride, ldo, o_bstride, bias, relu);
  else
    k_rows_self<__nv_bfloat16><<<grid, kMmThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                            static_cast<__nv_bfloat16*>(out), idx, n_idx, batch,
                                                            feat, ldx, x_bstride, ldo, o_bstride, bias, relu);
  GWEN_LAUNCH_CHECK("k_rows_self");
  return GWEN_OK;
}
