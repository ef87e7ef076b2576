// Caller-side loss of the reference training loop (src/gwen/models_gnn.py:261-265, used at :370):
//   loss_func(output, target, mask) = L1Loss()(output[mask], target[mask])
//                                   = sum_{b, n: mask[n], c} |output - target| / (B * count(mask) * C)
// The reference's boolean-mask indexing launches nonzero + two gathers and synchronises the host
// (the size of output[mask] is data dependent).  Here: one pass for the value (per-block partial
// sums in a fixed order, then a fixed-order final sum -> deterministic), one pass for the gradient
// dL/doutput = sign(output - target) * mask * dloss / (B * count * C); no host synchronisation, so the
// training step can be enqueued (or graph-captured) without a round trip.
#include "common.cuh"

namespace gwen {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossRows = 64;  // node rows per block

// part[g] = {sum |y - t| over the masked rows of row group g (kLossRows rows), number of masked rows}.
// One WARP per row group: the lanes read the group's mask bytes at once (two rows each), and only for a masked row
// -- rare in the reference's use: one node in 125 -- the warp sums |y - t| over batch x feat (lanes strided over the
// features, fixed-order shuffle tree) and lane 0 adds the row sums in row order: deterministic, no shared memory,
// no block barrier.  (First version: a 256-thread block per group walking the mask bytes one dependent load at a
// time and an 8-step block reduction for mostly empty groups: 116 us at the cfg 5 member shape.)
template <typename T>
__global__ void __launch_bounds__(kLossThreads) k_l1_partial(const T* __restrict__ y,
                                                             const T* __restrict__ t,
                                                             const uint8_t* __restrict__ mask,
                                                             int64_t batch, int64_t n, int64_t feat,
                                                             int64_t groups, float2* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t g = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (g >= groups) return;
  const int64_t r0 = g * kLossRows;
  static_assert(kLossRows == 64, "two mask bytes per lane");
  const int64_t ra = r0 + lane, rb = r0 + 32 + lane;
  const unsigned ma = __ballot_sync(0xffffffffu, ra < n && mask[ra] != 0);
  const unsigned mb = __ballot_sync(0xffffffffu, rb < n && mask[rb] != 0);
  float acc = 0.0f;
  for (int half = 0; half < 2; ++half) {
    unsigned bits = half ? mb : ma;
    while (bits) {
      const int j = __ffs(bits) - 1;
      bits &= bits - 1;
      const int64_t r = r0 + 32 * half + j;
      float s = 0.0f;
      for (int64_t b = 0; b < batch; ++b) {
        const T* yr = y + (b * n + r) * feat;
        const T* tr = t + (b * n + r) * feat;
        for (int64_t c = lane; c < feat; c += 32) s += fabsf(to_f32(yr[c]) - to_f32(tr[c]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      acc += s;
    }
  }
  if (lane == 0) part[g] = make_float2(acc, float(__popc(ma) + __popc(mb)));
}

// out[0] = loss, out[1] = 1 / (B * count * C) (the scale the backward pass uses)
__global__ void __launch_bounds__(kLossThreads) k_l1_final(const float2* __restrict__ part, int blocks,
                                                           int64_t batch, int64_t feat,
                                                           const float* __restrict__ count_override,
                                                           float* __restrict__ out) {
  __shared__ double s_sum[kLossThreads], s_cnt[kLossThreads];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < blocks; i += kLossThreads) {
    a += part[i].x;
    c += part[i].y;
  }
  s_sum[threadIdx.x] = a;
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  for (int s = kLossThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + s];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double denom = double(batch) * (count_override ? double(count_override[0]) : s_cnt[0]) * double(feat);
    out[0] = float(s_sum[0] / denom);   // 0 / 0 = NaN for an empty mask, like mean() of an empty tensor
    out[1] = float(1.0 / denom);
  }
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads) k_l1_bwd(const T* __restrict__ y, const T* __restrict__ t,
                                                         const uint8_t* __restrict__ mask,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ dloss, int64_t batch,
                                                         int64_t n, int64_t feat, T* __restrict__ dy) {
  const float g = scale[1] * (dloss ? dloss[0] : 1.0f);
  const int64_t total = batch * n * feat;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = (i / feat) % n;
    float v = 0.0f;
    if (mask[r]) {
      const float d = to_f32(y[i]) - to_f32(t[i]);
      v = d > 0.0f ? g : (d < 0.0f ? -g : 0.0f);   // sign(0) = 0, as torch's L1 backward
    }
    dy[i] = from_f32<T>(v);
  }
}

// The same gradient sixteen bytes per thread (feat a multiple of the vector width, dense rows): one mask byte per
// vector, y and t are only read on masked rows, unmasked rows are one 16-byte store of zeros -- the scalar kernel
// (a 64-bit division and a 2-byte store per element) took 160 us for the 115 MB of dy at the cfg 5 member shape.
template <typename T>
__global__ void __launch_bounds__(kLossThreads) k_l1_bwd_vec(const T* __restrict__ y, const T* __restrict__ t,
                                                             const uint8_t* __restrict__ mask,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ dloss, int64_t n,
                                                             int64_t vecs_per_row, int64_t total_vecs,
                                                             T* __restrict__ dy) {
  constexpr int VN = 16 / sizeof(T);
  const float g = scale[1] * (dloss ? dloss[0] : 1.0f);
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < total_vecs;
       v += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = (v / vecs_per_row) % n;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (mask[r]) {
      const uint4 a = *reinterpret_cast<const uint4*>(y + v * VN);
      const uint4 b = *reinterpret_cast<const uint4*>(t + v * VN);
      const T* ya = reinterpret_cast<const T*>(&a);
      const T* tb = reinterpret_cast<const T*>(&b);
      T* oo = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float d = to_f32(ya[k]) - to_f32(tb[k]);
        oo[k] = from_f32<T>(d > 0.0f ? g : (d < 0.0f ? -g : 0.0f));
      }
    }
    *reinterpret_cast<uint4*>(dy + v * VN) = o;
  }
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_masked_l1_workspace_bytes(int64_t n, size_t* out) {
  GWEN_CHECK_ARG(out && n >= 0, "bad arguments");
  *out = static_cast<size_t>(ceil_div(std::max<int64_t>(n, 1), kLossRows)) * sizeof(float2) + 256;
  return GWEN_OK;
}

extern "C" int gwen_masked_l1_fwd(const void* y, const void* target, const uint8_t* mask, int64_t batch,
                                  int64_t n, int64_t feat, int dtype, const float* count_override,
                                  float* loss_and_scale, void* ws, size_t ws_bytes, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && n >= 0 && feat >= 0, "negative size");
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  GWEN_CHECK_ARG(loss_and_scale && ws && (n == 0 || (y && target && mask)), "null pointer");
  const int blocks = static_cast<int>(ceil_div(std::max<int64_t>(n, 1), kLossRows));
  if (ws_bytes < size_t(blocks) * sizeof(float2)) return set_err(GWEN_E_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float2* part = static_cast<float2*>(ws);
  const unsigned pblocks = static_cast<unsigned>(ceil_div(int64_t(blocks) * 32, kLossThreads));   // a warp per row group
  if (dtype == GWEN_F32)
    k_l1_partial<float><<<pblocks, kLossThreads, 0, st>>>(static_cast<const float*>(y),
                                                          static_cast<const float*>(target), mask, batch, n,
                                                          feat, blocks, part);
  else
    k_l1_partial<__nv_bfloat16><<<pblocks, kLossThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(target), mask, batch, n,
        feat, blocks, part);
  GWEN_LAUNCH_CHECK("k_l1_partial");
  k_l1_final<<<1, kLossThreads, 0, st>>>(part, blocks, batch, feat, count_override, loss_and_scale);
  GWEN_LAUNCH_CHECK("k_l1_final");
  return GWEN_OK;
}

extern "C" int gwen_masked_l1_bwd(const void* y, const void* target, const uint8_t* mask,
                                  const float* loss_and_scale, const float* dloss, int64_t batch,
                                  int64_t n, int64_t feat, int dtype, void* dy, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && n >= 0 && feat >= 0, "negative size");
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  if (batch * n * feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(y && target && mask && loss_and_scale && dy, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vn = dtype == GWEN_F32 ? 4 : 8;
  if (feat % vn == 0 && aligned16(y) && aligned16(target) && aligned16(dy)) {
    const int64_t vpr = feat / vn, total_vecs = batch * n * vpr;
    const unsigned vblocks = static_cast<unsigned>(
        std::min<int64_t>(ceil_div(total_vecs, kLossThreads), int64_t(sm_count()) * 16));
    if (dtype == GWEN_F32)
      k_l1_bwd_vec<float><<<vblocks, kLossThreads, 0, st>>>(static_cast<const float*>(y),
                                                            static_cast<const float*>(target), mask, loss_and_scale,
                                                            dloss, n, vpr, total_vecs, static_cast<float*>(dy));
    else
      k_l1_bwd_vec<__nv_bfloat16><<<vblocks, kLossThreads, 0, st>>>(
          static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(target), mask, loss_and_scale, dloss,
          n, vpr, total_vecs, static_cast<__nv_bfloat16*>(dy));
    GWEN_LAUNCH_CHECK("k_l1_bwd_vec");
    return GWEN_OK;
  }
  const unsigned blocks = static_cast<unsigned>(
      std::min<int64_t>(ceil_div(batch * n * feat, kLossThreads), int64_t(sm_count()) * 16));
  if (dtype == GWEN_F32)
    k_l1_bwd<float><<<blocks, kLossThreads, 0, st>>>(static_cast<const float*>(y),
                                                     static_cast<const float*>(target), mask,
                                                     loss_and_scale, dloss, batch, n, feat,
                                                     static_cast<float*>(dy));
  else
    k_l1_bwd<__nv_bfloat16><<<blocks, kLossThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(target), mask,
        loss_and_scale, dloss, batch, n, feat, static_cast<__nv_bfloat16*>(dy));
  GWEN_LAUNCH_CHECK("k_l1_bwd");
  return GWEN_OK;
}
