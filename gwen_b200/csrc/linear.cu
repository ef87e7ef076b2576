// K2 (CUDA-core path): dense projections y = x W^T (+bias, relu) and the two backward GEMMs, as a
// register-blocked FFMA GEMM with exact fp32 accumulation.  This is the path for fp32 (the
// reference runs fp32 Linear on SIMT cuBLAS: allow_tf32 defaults to False) and for shapes the
// tcgen05 kernel in linear_tc.cu does not take (ragged K, tiny M).  bf16 problems with
// tensor-core-friendly shapes are routed to linear_tc.cu by gwen_linear_fwd.
//
// One kernel serves forward, dgrad and wgrad through operand layout flags:
//   C[i, j] = sum_r A(i, r) * B(j, r)
//   A_RMAJOR: A(i, r) = A[i*lda + r]  (reduction index contiguous)   else A[r*lda + i]
//   B_RMAJOR: B(j, r) = B[j*ldb + r]                                  else B[r*ldb + j]
//   forward: A = x  (R-major), B = W  (R-major), reduce over K
//   dgrad  : A = dy (R-major), B = W  (col-major in r = n: B[r*ldw + j]), reduce over N_out
//   wgrad  : A = dy (A[r*lddy + i]), B = x (B[r*ldx + j]), reduce over M, split across CTAs
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace gwen {

// linear_tc.cu
int linear_tc_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                        const void* x, const void* w, const void* y);
int linear_tc_fwd_bf16(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                       int64_t ldx, int64_t ldw, int64_t ldy, const float* bias, int relu,
                       cudaStream_t st);

int linear_tc_dgrad_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldw,
                              int64_t lddx, const void* dy, const void* w, const void* dx);
int linear_tc_dgrad_bf16(const void* dy, const void* w, void* dx, int64_t m, int64_t k_in,
                         int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, cudaStream_t st,
                         const void* mask = nullptr, int64_t ldm = 0);

int linear_tc_batched_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                                int64_t x_bstride, int64_t y_bstride, const void* x, const void* w,
                                const void* y, int b_mn);
int linear_tc_batched_bf16(const void* x, const void* w, void* y, int64_t batch, int64_t m, int64_t k,
                           int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int64_t x_bstride,
                           int64_t y_bstride, const float* bias, int relu, int b_mn, cudaStream_t st,
                           const void* mask = nullptr, int64_t ldm = 0, int64_t m_bstride = 0);
// linear_tf32x3.cu
int linear_tf32x3_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                            const void* x, const void* w, const void* y);
size_t linear_tf32x3_workspace_bytes(int64_t m, int64_t k, int64_t n_out);
int linear_tf32x3_fwd(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                      int64_t ldy, const float* bias, int relu, void* ws, size_t ws_bytes,
                      cudaStream_t st);
int linear_tf32x3_dgrad(const void* dy, const void* w, void* dx, int64_t m, int64_t k_in, int64_t n_out,
                        int64_t lddx, void* ws, size_t ws_bytes, cudaStream_t st);
// linear_wgrad_tc.cu
int linear_tc_wgrad_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldx,
                              const void* dy, const void* x);
int linear_tc_wgrad_splits(int64_t m, int64_t k_in, int64_t n_out, int with_db = 0);
int linear_tc_wgrad_bf16(const void* dy, const void* x, float* part, int64_t m, int64_t k_in,
                         int64_t n_out, int64_t lddy, int64_t ldx, cudaStream_t st, float* dbpart = nullptr);

// linear_wgrad_tf32x3.cu
int linear_wgrad_tf32x3_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldx,
                                  const void* dy, const void* x);
int linear_wgrad_tf32x3_splits(int64_t m, int64_t k_in, int64_t n_out);
int linear_wgrad_tf32x3(const void* dy, const void* x, float* part, int64_t m, int64_t k_in, int64_t n_out,
                        int64_t lddy, int64_t ldx, cudaStream_t st);

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int kGemmThreads = 256;

struct GemmArgs {
  const void* a;
  const void* b;
  void* c;         // T (or float when OUT_F32)
  const float* bias;
  int64_t m, n, r;  // C is m x n, reduction length r
  int64_t lda, ldb, ldc;
  int64_t r_per_split;  // reduction slice per blockIdx.z
  int64_t c_split_stride;
  int relu;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, int64_t stride, int valid, bool vec, float* f) {
  // 4 consecutive (stride 1) elements starting at p, `valid` of them in range.
  if (vec && valid == 4) {
    if constexpr (sizeof(T) == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p));
      f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
      f[0] = __uint_as_float(v.x << 16);
      f[1] = __uint_as_float(v.x & 0xffff0000u);
      f[2] = __uint_as_float(v.y << 16);
      f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = i < valid ? to_f32(p[i * stride]) : 0.0f;
  }
}

// Stage one operand tile (128 outer x 16 reduction) into registers.  RMAJOR: each thread reads
// 2 x 4 elements along r; otherwise 2 x 4 elements along the outer index.
template <typename T, bool RMAJOR>
__device__ __forceinline__ void fetch_tile(const T* base, int64_t ld, int64_t outer0,
                                           int64_t outer_n, int64_t r0, int64_t r_end, bool vec,
                                           float (&reg)[2][4]) {
  const int tid = threadIdx.x;
  if constexpr (RMAJOR) {
    const int rq = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t o = outer0 + (tid >> 2) + h * 64;
      const int64_t r = r0 + rq;
      int valid = o < outer_n ? static_cast<int>(max(int64_t(0), min(int64_t(4), r_end - r))) : 0;
      load4<T>(base + o * ld + r, 1, valid, vec, reg[h]);
    }
  } else {
    const int oq = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = r0 + (tid >> 5) + h * 8;
      const int64_t o = outer0 + oq;
      int valid = r < r_end ? static_cast<int>(max(int64_t(0), min(int64_t(4), outer_n - o))) : 0;
      load4<T>(base + r * ld + o, 1, valid, vec, reg[h]);
    }
  }
}

template <bool RMAJOR>
__device__ __forceinline__ void stash_tile(float (*s)[BM + 4], const float (&reg)[2][4]) {
  const int tid = threadIdx.x;
  if constexpr (RMAJOR) {
    const int rq = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = (tid >> 2) + h * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[rq + i][o] = reg[h][i];
    }
  } else {
    const int oq = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = (tid >> 5) + h * 8;
      *reinterpret_cast<float4*>(&s[r][oq]) = make_float4(reg[h][0], reg[h][1], reg[h][2], reg[h][3]);
    }
  }
}

template <typename T, bool A_RMAJOR, bool B_RMAJOR, bool OUT_F32>
__global__ void __launch_bounds__(kGemmThreads) k_gemm(GemmArgs g, bool vec_a, bool vec_b) {
  __shared__ __align__(16) float sa[2][BK][BM + 4];
  __shared__ __align__(16) float sb[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // 1-D tile index, N tiles fastest: CTAs that run together share the A tile through L2.
  const int64_t n_tiles = (g.n + BN - 1) / BN;
  const int64_t m0 = (int64_t(blockIdx.x) / n_tiles) * BM, n0 = (int64_t(blockIdx.x) % n_tiles) * BN;
  const int64_t r_beg = int64_t(blockIdx.y) * g.r_per_split;
  const int64_t r_end = min(g.r, r_beg + g.r_per_split);
  const T* A = static_cast<const T*>(g.a);
  const T* B = static_cast<const T*>(g.b);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float ra[2][4], rb[2][4];
  fetch_tile<T, A_RMAJOR>(A, g.lda, m0, g.m, r_beg, r_end, vec_a, ra);
  fetch_tile<T, B_RMAJOR>(B, g.ldb, n0, g.n, r_beg, r_end, vec_b, rb);
  stash_tile<A_RMAJOR>(sa[0], ra);
  stash_tile<B_RMAJOR>(sb[0], rb);
  __syncthreads();
  int buf = 0;
  for (int64_t r0 = r_beg; r0 < r_end; r0 += BK) {
    const bool more = r0 + BK < r_end;
    if (more) {
      fetch_tile<T, A_RMAJOR>(A, g.lda, m0, g.m, r0 + BK, r_end, vec_a, ra);
      fetch_tile<T, B_RMAJOR>(B, g.ldb, n0, g.n, r0 + BK, r_end, vec_b, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sa[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sa[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sb[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sb[buf][kk][64 + tx * 4]);
      const float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[TN] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash_tile<A_RMAJOR>(sa[buf ^ 1], ra);
      stash_tile<B_RMAJOR>(sb[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }
  // epilogue
  using OutT = typename std::conditional<OUT_F32, float, T>::type;
  OutT* C = static_cast<OutT*>(g.c) + int64_t(blockIdx.y) * g.c_split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.m) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.n) continue;
      float v = acc[i][j];
      if (g.bias) v += __ldg(g.bias + n);
      if (g.relu) v = fmaxf(v, 0.0f);
      if constexpr (OUT_F32) C[m * g.ldc + n] = v; else C[m * g.ldc + n] = from_f32<T>(v);
    }
  }
}

template <typename T>
bool vec_ok(const void* p, int64_t ld) {
  const int a = sizeof(T) == 4 ? 16 : 8;
  return (reinterpret_cast<uintptr_t>(p) % a) == 0 && ld % 4 == 0;
}

template <typename T, bool AR, bool BR, bool OUT_F32>
int launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
  const int64_t tiles = ceil_div(g.n, BN) * ceil_div(g.m, BM);
  if (tiles > INT32_MAX || splits > 65535) return set_err(GWEN_E_NOSUPPORT, "GEMM grid too large");
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(splits));
  k_gemm<T, AR, BR, OUT_F32><<<grid, kGemmThreads, 0, st>>>(g, vec_ok<T>(g.a, g.lda),
                                                            vec_ok<T>(g.b, g.ldb));
  GWEN_LAUNCH_CHECK("k_gemm");
  return GWEN_OK;
}

// fixed-order reduction of wgrad split partials: out[i] = sum_s part[s][i], s ascending
__global__ void k_reduce_splits(const float* __restrict__ part, int splits, int64_t n_elems,
                                int64_t stride, float* __restrict__ out, int64_t ld_out,
                                int64_t cols) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n_elems) return;
  float s = 0.0f;
  for (int k = 0; k < splits; ++k) s += part[k * stride + i];
  out[(i / cols) * ld_out + (i % cols)] = s;
}

// Partial-sum block for long reductions: block (x, g) sums splits [g * group, (g + 1) * group) of
// `part` into tmp[g], so a reduction over thousands of partial rows (the bias gradient at COSMO-1E size:
// ~3500 x feat) is two short fixed-order passes instead of `feat` threads looping over all of them.
__global__ void k_reduce_groups(const float* __restrict__ part, int splits, int group, int64_t n_elems,
                                float* __restrict__ tmp) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n_elems) return;
  const int s0 = blockIdx.y * group, s1 = min(splits, s0 + group);
  float s = 0.0f;
  for (int k = s0; k < s1; ++k) s += part[int64_t(k) * n_elems + i];
  tmp[int64_t(blockIdx.y) * n_elems + i] = s;
}

constexpr int kReduceGroup = 64;
inline int64_t reduce_groups(int64_t splits) { return splits > 2 * kReduceGroup ? ceil_div(splits, kReduceGroup) : 0; }

// out[i] = sum over splits of part[k][i], fixed order; `tmp` (reduce_groups(splits) x n_elems floats) is
// used when the reduction is long.
int reduce_vector(const float* part, int64_t splits, int64_t n_elems, float* out, float* tmp, cudaStream_t st) {
  const int64_t groups = reduce_groups(splits);
  const unsigned bx = static_cast<unsigned>(ceil_div(n_elems, 256));
  if (groups > 0 && tmp) {
    k_reduce_groups<<<dim3(bx, static_cast<unsigned>(groups)), 256, 0, st>>>(
        part, static_cast<int>(splits), kReduceGroup, n_elems, tmp);
    GWEN_LAUNCH_CHECK("k_reduce_groups");
    k_reduce_splits<<<bx, 256, 0, st>>>(tmp, static_cast<int>(groups), n_elems, n_elems, out, n_elems, n_elems);
  } else {
    k_reduce_splits<<<bx, 256, 0, st>>>(part, static_cast<int>(splits), n_elems, n_elems, out, n_elems, n_elems);
  }
  GWEN_LAUNCH_CHECK("k_reduce_splits");
  return GWEN_OK;
}

int wgrad_splits(int64_t m, int64_t k, int64_t n_out) {
  const int64_t tiles = ceil_div(n_out, BM) * ceil_div(k, BN);
  int64_t s = ceil_div(int64_t(sm_count()) * 2, tiles);
  s = std::min<int64_t>(s, ceil_div(m, 4 * BK));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(s, 512)));
}

template <typename T>
__global__ void k_relu_bwd(const T* __restrict__ y, T* __restrict__ dy, int64_t rows, int64_t feat,
                           int64_t ldy, int64_t lddy) {
  const int64_t total = rows * feat;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / feat, c = i % feat;
    if (!(to_f32(y[r * ldy + c]) > 0.0f)) dy[r * lddy + c] = from_f32<T>(0.0f);
  }
}

// db partials: block (bx, by) sums rows [by*chunk, (by+1)*chunk) of columns bx*32..+31 in a fixed
// order: 8 row-lanes stride the chunk sequentially, then a fixed 8-way tree in shared memory.
template <typename T>
__global__ void __launch_bounds__(256) k_bias_partial(const T* __restrict__ dy, int64_t rows,
                                                      int64_t feat, int64_t lddy, int64_t chunk,
                                                      float* __restrict__ part) {
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = int64_t(blockIdx.x) * 32 + cx;
  const int64_t r0 = int64_t(blockIdx.y) * chunk, r1 = min(rows, r0 + chunk);
  float acc = 0.0f;
  if (c < feat)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += to_f32(dy[r * lddy + c]);
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < feat) {
    float t = ((s[0][cx] + s[1][cx]) + (s[2][cx] + s[3][cx])) +
              ((s[4][cx] + s[5][cx]) + (s[6][cx] + s[7][cx]));
    part[int64_t(blockIdx.y) * feat + c] = t;
  }
}

constexpr int64_t kBiasChunk = 2048;
// rows per block of the fused mask + bias pass: small chunks = enough 16-byte loads in flight per SM
// (438 blocks of 2048 rows left the kernel at 4.7 TB/s)
constexpr int64_t kFusedChunk = 512;

// Fused epilogue backward: dz = dy * (y > 0) (written out of place; skipped when y == nullptr) and
// the per-chunk column sums of dz for the bias gradient, in ONE pass over 16-byte vectors.
// Thread t owns column group t % groups (groups = feat / VN divides 256) and the rows
// r0 + t / groups, + 256 / groups, ...: sums in a fixed order, then a fixed-order cross-thread sum
// in shared memory -> deterministic.
template <typename T>
__global__ void __launch_bounds__(256) k_relu_bias_bwd(const T* __restrict__ y,
                                                       const T* __restrict__ dy, T* __restrict__ dz,
                                                       int64_t rows, int groups, int64_t chunk,
                                                       float* __restrict__ part) {
  constexpr int VN = 16 / sizeof(T);
  __shared__ float s[256][VN + 1];
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups, rstep = 256 / groups;
  const int64_t r0 = int64_t(blockIdx.x) * chunk, r1 = min(rows, r0 + chunk);
  const int64_t feat = int64_t(groups) * VN;
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.0f;
  for (int64_t r = r0 + rl; r < r1; r += rstep) {
    const int64_t off = r * feat + int64_t(cg) * VN;
    uint4 d = __ldg(reinterpret_cast<const uint4*>(dy + off));
    if (y) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(y + off));
      if constexpr (sizeof(T) == 4) {
        if (!(__uint_as_float(v.x) > 0.0f)) d.x = 0u;
        if (!(__uint_as_float(v.y) > 0.0f)) d.y = 0u;
        if (!(__uint_as_float(v.z) > 0.0f)) d.z = 0u;
        if (!(__uint_as_float(v.w) > 0.0f)) d.w = 0u;
      } else {
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
        uint32_t dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 > 0  <=>  sign clear and magnitude non-zero (not NaN-safe, as relu)
          const uint32_t lo = vv[i] & 0xffffu, hi = vv[i] >> 16;
          const uint32_t mlo = (lo != 0u && lo < 0x8000u) ? 0xffffu : 0u;
          const uint32_t mhi = (hi != 0u && hi < 0x8000u) ? 0xffff0000u : 0u;
          dd[i] &= (mlo | mhi);
        }
        d = make_uint4(dd[0], dd[1], dd[2], dd[3]);
      }
      *reinterpret_cast<uint4*>(dz + off) = d;
    }
    if (part) {
      if constexpr (sizeof(T) == 4) {
        acc[0] += __uint_as_float(d.x); acc[1] += __uint_as_float(d.y);
        acc[2] += __uint_as_float(d.z); acc[3] += __uint_as_float(d.w);
      } else {
        const uint32_t dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += __uint_as_float(dd[i] << 16);
          acc[2 * i + 1] += __uint_as_float(dd[i] & 0xffff0000u);
        }
      }
    }
  }
  if (!part) return;
#pragma unroll
  for (int k = 0; k < VN; ++k) s[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < groups) {
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      float t = 0.0f;
      for (int j = 0; j < rstep; ++j) t += s[j * groups + threadIdx.x][k];
      part[int64_t(blockIdx.x) * feat + int64_t(threadIdx.x) * VN + k] = t;
    }
  }
}

}  // namespace
}  // namespace gwen

using namespace gwen;

static int check_dtype(int dtype) {
  if (dtype == GWEN_F32 || dtype == GWEN_BF16) return GWEN_OK;
  return set_err(GWEN_E_DTYPE, "unknown dtype %d", dtype);
}

extern "C" int gwen_linear_fwd(const void* x, const void* weight, void* y, int64_t m, int64_t k,
                               int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int dtype,
                               const float* bias, int epilogue, void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (m == 0 || n_out == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && weight && y, "null pointer");
  GWEN_CHECK_ARG(ldx >= k && ldw >= k && ldy >= n_out, "row pitch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int relu = (epilogue & GWEN_EPI_RELU) ? 1 : 0;
  if (dtype == GWEN_BF16 && linear_tc_supported(m, k, n_out, ldx, ldw, ldy, x, weight, y))
    return linear_tc_fwd_bf16(x, weight, y, m, k, n_out, ldx, ldw, ldy, bias, relu, st);
  GemmArgs g{x, weight, y, bias, m, n_out, k, ldx, ldw, ldy, k, 0, relu};
  return dtype == GWEN_F32 ? launch_gemm<float, true, true, false>(g, 1, st)
                           : launch_gemm<__nv_bfloat16, true, true, false>(g, 1, st);
}

extern "C" int gwen_linear_fwd_workspace_bytes(int64_t m, int64_t k, int64_t n_out, int dtype,
                                               size_t* out) {
  GWEN_CHECK_ARG(out && m >= 0 && k >= 0 && n_out >= 0, "bad arguments");
  *out = (dtype == GWEN_F32 && m >= 4096 && k >= 32 && k % 4 == 0 && n_out % 64 == 0)
             ? linear_tf32x3_workspace_bytes(m, k, n_out) : 0;   // the split copy of W
  return GWEN_OK;
}

extern "C" int gwen_linear_fwd_ws(const void* x, const void* weight, void* y, int64_t m, int64_t k,
                                  int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int dtype,
                                  const float* bias, int epilogue, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (dtype == GWEN_F32 && ws && m > 0 && x && weight && y &&
      linear_tf32x3_supported(m, k, n_out, ldx, ldw, ldy, x, weight, y) &&
      ws_bytes >= linear_tf32x3_workspace_bytes(m, k, n_out))
    return linear_tf32x3_fwd(x, weight, y, m, k, n_out, ldy, bias, (epilogue & GWEN_EPI_RELU) ? 1 : 0, ws,
                             ws_bytes, static_cast<cudaStream_t>(stream));
  return gwen_linear_fwd(x, weight, y, m, k, n_out, ldx, ldw, ldy, dtype, bias, epilogue, stream);
}

extern "C" int gwen_linear_bwd_data(const void* dy, const void* weight, void* dx, int64_t m,
                                    int64_t k, int64_t n_out, int64_t lddy, int64_t ldw,
                                    int64_t lddx, int dtype, void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (m == 0 || k == 0) return GWEN_OK;
  GWEN_CHECK_ARG(dy && weight && dx, "null pointer");
  GWEN_CHECK_ARG(lddy >= n_out && ldw >= k && lddx >= k, "row pitch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GWEN_BF16 && linear_tc_dgrad_supported(m, k, n_out, lddy, ldw, lddx, dy, weight, dx))
    return linear_tc_dgrad_bf16(dy, weight, dx, m, k, n_out, lddy, ldw, lddx, st);
  // dx[m, kk] = sum_n dy[m, n] * W[n, kk]:  A = dy (R-major), B(j=kk, r=n) = W[r*ldw + j]
  GemmArgs g{dy, weight, dx, nullptr, m, k, n_out, lddy, ldw, lddx, n_out, 0, 0};
  return dtype == GWEN_F32 ? launch_gemm<float, true, false, false>(g, 1, st)
                           : launch_gemm<__nv_bfloat16, true, false, false>(g, 1, st);
}

// dx = (dy W) * (mask > 0): the dgrad with the previous layer's ReLU backward folded into the GEMM epilogue
// (tcgen05 CTA-pair kernel only; GWEN_E_NOSUPPORT otherwise -- the caller then runs gwen_linear_bwd_data +
// gwen_relu_bias_bwd, which give the same bits)
extern "C" int gwen_linear_bwd_data_masked(const void* dy, const void* weight, void* dx, const void* mask,
                                           int64_t m, int64_t k, int64_t n_out, int64_t lddy, int64_t ldw,
                                           int64_t lddx, int64_t ldmask, int dtype, void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  GWEN_CHECK_ARG(dy && weight && dx && mask, "null pointer");
  GWEN_CHECK_ARG(lddy >= n_out && ldw >= k && lddx >= k && ldmask >= k, "row pitch too small");
  if (dtype != GWEN_BF16 || m == 0 || k == 0 || ldmask % 8 || (reinterpret_cast<uintptr_t>(mask) & 15u) ||
      !linear_tc_dgrad_supported(m, k, n_out, lddy, ldw, lddx, dy, weight, dx))
    return set_err(GWEN_E_NOSUPPORT, "masked dgrad runs on the bf16 tcgen05 pair kernel only");
  return linear_tc_dgrad_bf16(dy, weight, dx, m, k, n_out, lddy, ldw, lddx, static_cast<cudaStream_t>(stream), mask,
                              ldmask);
}

// The same over `batch` slices in ONE launch (batch-strided dy / dx / mask, e.g. the owned rows of a band buffer)
extern "C" int gwen_linear_batched_bwd_data_masked(const void* dy, const void* weight, void* dx, const void* mask,
                                                   int64_t batch, int64_t m, int64_t k, int64_t n_out, int64_t lddy,
                                                   int64_t ldw, int64_t lddx, int64_t ldmask, int64_t dy_bstride,
                                                   int64_t dx_bstride, int64_t mask_bstride, int dtype, void* stream) {
  GWEN_CHECK_ARG(batch >= 1 && m >= 0 && k >= 0 && n_out >= 0, "bad sizes");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  GWEN_CHECK_ARG(dy && weight && dx && mask, "null pointer");
  GWEN_CHECK_ARG(lddy >= n_out && ldw >= k && lddx >= k && ldmask >= k, "row pitch too small");
  if (dtype != GWEN_BF16 || m == 0 || k == 0 || ldmask % 8 || mask_bstride % 8 || (reinterpret_cast<uintptr_t>(mask) & 15u) ||
      !linear_tc_batched_supported(m, n_out, k, lddy, ldw, lddx, dy_bstride, dx_bstride, dy, weight, dx, 1))
    return set_err(GWEN_E_NOSUPPORT, "masked dgrad runs on the bf16 tcgen05 pair kernel only");
  return linear_tc_batched_bf16(dy, weight, dx, batch, m, n_out, k, lddy, ldw, lddx, dy_bstride, dx_bstride, nullptr, 0,
                                1, static_cast<cudaStream_t>(stream), mask, ldmask, mask_bstride);
}

extern "C" int gwen_linear_batched_fwd(const void* x, const void* weight, void* y, int64_t batch, int64_t m,
                                       int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                                       int64_t x_bstride, int64_t y_bstride, int dtype, const float* bias,
                                       int epilogue, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (batch == 0 || m == 0 || n_out == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && weight && y, "null pointer");
  if (dtype == GWEN_BF16 && batch > 1 &&
      linear_tc_batched_supported(m, k, n_out, ldx, ldw, ldy, x_bstride, y_bstride, x, weight, y, 0))
    return linear_tc_batched_bf16(x, weight, y, batch, m, k, n_out, ldx, ldw, ldy, x_bstride, y_bstride, bias,
                                  (epilogue & GWEN_EPI_RELU) ? 1 : 0, 0, static_cast<cudaStream_t>(stream));
  const int64_t esz = dtype == GWEN_F32 ? 4 : 2;
  for (int64_t b = 0; b < batch; ++b) {
    int rc = gwen_linear_fwd(static_cast<const char*>(x) + b * x_bstride * esz, weight,
                             static_cast<char*>(y) + b * y_bstride * esz, m, k, n_out, ldx, ldw, ldy, dtype,
                             bias, epilogue, stream);
    if (rc != GWEN_OK) return rc;
  }
  return GWEN_OK;
}

extern "C" int gwen_linear_batched_bwd_data(const void* dy, const void* weight, void* dx, int64_t batch,
                                            int64_t m, int64_t k, int64_t n_out, int64_t lddy, int64_t ldw,
                                            int64_t lddx, int64_t dy_bstride, int64_t dx_bstride, int dtype,
                                            void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (batch == 0 || m == 0 || k == 0) return GWEN_OK;
  GWEN_CHECK_ARG(dy && weight && dx, "null pointer");
  if (dtype == GWEN_BF16 && batch > 1 &&
      linear_tc_batched_supported(m, n_out, k, lddy, ldw, lddx, dy_bstride, dx_bstride, dy, weight, dx, 1))
    return linear_tc_batched_bf16(dy, weight, dx, batch, m, n_out, k, lddy, ldw, lddx, dy_bstride, dx_bstride,
                                  nullptr, 0, 1, static_cast<cudaStream_t>(stream));
  const int64_t esz = dtype == GWEN_F32 ? 4 : 2;
  for (int64_t b = 0; b < batch; ++b) {
    int rc = gwen_linear_bwd_data(static_cast<const char*>(dy) + b * dy_bstride * esz, weight,
                                  static_cast<char*>(dx) + b * dx_bstride * esz, m, k, n_out, lddy, ldw, lddx,
                                  dtype, stream);
    if (rc != GWEN_OK) return rc;
  }
  return GWEN_OK;
}

extern "C" int gwen_linear_bwd_data_workspace_bytes(int64_t m, int64_t k, int64_t n_out, int dtype,
                                                    size_t* out) {
  GWEN_CHECK_ARG(out && m >= 0 && k >= 0 && n_out >= 0, "bad arguments");
  *out = (dtype == GWEN_F32 && m >= 4096 && n_out >= 32 && n_out % 4 == 0 && k % 64 == 0)
             ? linear_tf32x3_workspace_bytes(m, n_out, k) : 0;   // split copy of W^T
  return GWEN_OK;
}

extern "C" int gwen_linear_bwd_data_ws(const void* dy, const void* weight, void* dx, int64_t m, int64_t k,
                                       int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, int dtype,
                                       void* ws, size_t ws_bytes, void* stream) {
  // fp32: dx = dy (W^T)^T through the 3xTF32 forward kernel (reduction over n_out, k output columns)
  if (dtype == GWEN_F32 && ws && m > 0 && dy && weight && dx && ldw == k &&
      linear_tf32x3_supported(m, n_out, k, lddy, n_out, lddx, dy, weight, dx) &&
      ws_bytes >= linear_tf32x3_workspace_bytes(m, n_out, k))
    return linear_tf32x3_dgrad(dy, weight, dx, m, k, n_out, lddx, ws, ws_bytes,
                               static_cast<cudaStream_t>(stream));
  return gwen_linear_bwd_data(dy, weight, dx, m, k, n_out, lddy, ldw, lddx, dtype, stream);
}

extern "C" int gwen_linear_bwd_weight_workspace_bytes(int64_t m, int64_t k, int64_t n_out,
                                                      size_t* out) {
  GWEN_CHECK_ARG(out && m >= 0 && k >= 0 && n_out >= 0, "bad arguments");
  int splits = wgrad_splits(m, k, n_out);
  if (k % 64 == 0 && n_out % 64 == 0 && m >= 256) splits = std::max(splits, linear_tc_wgrad_splits(m, k, n_out));
  if (k % 4 == 0 && n_out % 4 == 0 && m >= 4096) splits = std::max(splits, linear_wgrad_tf32x3_splits(m, k, n_out));
  *out = static_cast<size_t>(splits) * n_out * k * sizeof(float) + 256;
  return GWEN_OK;
}

// dW AND db = column sums of dy from the same pass (bf16 tcgen05 wgrad only: ones-column MMA, k_wgrad_tc<kDb>)
extern "C" int gwen_linear_bwd_weight_bias_workspace_bytes(int64_t m, int64_t k, int64_t n_out, size_t* out) {
  GWEN_CHECK_ARG(out && m >= 0 && k >= 0 && n_out >= 0, "bad arguments");
  const int splits = (k % 64 == 0 && n_out % 64 == 0 && m >= 256) ? linear_tc_wgrad_splits(m, k, n_out, 1) : 1;
  *out = static_cast<size_t>(splits) * n_out * (k + 1) * sizeof(float) + 512;
  return GWEN_OK;
}

extern "C" int gwen_linear_bwd_weight_bias(const void* dy, const void* x, float* dw, float* db, int64_t m, int64_t k,
                                           int64_t n_out, int64_t lddy, int64_t ldx, int64_t lddw, int dtype,
                                           void* ws, size_t ws_bytes, void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  GWEN_CHECK_ARG(dy && x && dw && db && ws, "null pointer");
  GWEN_CHECK_ARG(lddy >= n_out && ldx >= k && lddw >= k, "row pitch too small");
  // K tiles of 256 columns fill tensor memory (2 x 256 accumulator columns): the bias sums would force 128-column
  // tiles, and that costs more than the pass it saves (measured, 512 -> 1024 at M = 896 292: 787 -> 1528 us)
  if (dtype != GWEN_BF16 || k % 256 == 0 || !linear_tc_wgrad_supported(m, k, n_out, lddy, ldx, dy, x))
    return set_err(GWEN_E_NOSUPPORT, "wgrad with bias sums runs on the bf16 tcgen05 kernel with K tiles <= 128 only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int splits = linear_tc_wgrad_splits(m, k, n_out, 1);
  const size_t wbytes = align_up(static_cast<size_t>(splits) * n_out * k * sizeof(float), 256);
  const size_t need = wbytes + static_cast<size_t>(splits) * n_out * sizeof(float);
  if (ws_bytes < need) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  float* part = static_cast<float*>(ws);
  float* dbpart = reinterpret_cast<float*>(static_cast<char*>(ws) + wbytes);
  int rc = linear_tc_wgrad_bf16(dy, x, part, m, k, n_out, lddy, ldx, st, dbpart);
  if (rc != GWEN_OK) return rc;
  const int64_t n_elems = n_out * k;
  k_reduce_splits<<<static_cast<unsigned>(ceil_div(n_elems, 256)), 256, 0, st>>>(part, splits, n_elems, n_elems, dw,
                                                                               lddw, k);
  GWEN_LAUNCH_CHECK("k_reduce_splits");
  k_reduce_splits<<<static_cast<unsigned>(ceil_div(n_out, 256)), 256, 0, st>>>(dbpart, splits, n_out, n_out, db, n_out,
                                                                             n_out);
  GWEN_LAUNCH_CHECK("k_reduce_splits (db)");
  return GWEN_OK;
}

extern "C" int gwen_linear_bwd_weight(const void* dy, const void* x, float* dw, int64_t m,
                                      int64_t k, int64_t n_out, int64_t lddy, int64_t ldx,
                                      int64_t lddw, int dtype, void* ws, size_t ws_bytes,
                                      void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k >= 0 && n_out >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (k == 0 || n_out == 0) return GWEN_OK;
  GWEN_CHECK_ARG(dw && ws, "null pointer");
  GWEN_CHECK_ARG(m == 0 || (dy && x), "null pointer");
  GWEN_CHECK_ARG(lddy >= n_out && ldx >= k && lddw >= k, "row pitch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_elems = n_out * k;
  if (dtype == GWEN_BF16 && linear_tc_wgrad_supported(m, k, n_out, lddy, ldx, dy, x)) {
    const int tsplits = linear_tc_wgrad_splits(m, k, n_out);
    const size_t tneed = static_cast<size_t>(tsplits) * n_out * k * sizeof(float);
    if (ws_bytes < tneed) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, tneed);
    int rc = linear_tc_wgrad_bf16(dy, x, static_cast<float*>(ws), m, k, n_out, lddy, ldx, st);
    if (rc != GWEN_OK) return rc;
    k_reduce_splits<<<static_cast<unsigned>(ceil_div(n_elems, 256)), 256, 0, st>>>(
        static_cast<const float*>(ws), tsplits, n_elems, n_elems, dw, lddw, k);
    GWEN_LAUNCH_CHECK("k_reduce_splits");
    return GWEN_OK;
  }
  if (dtype == GWEN_F32 && m > 0 && linear_wgrad_tf32x3_supported(m, k, n_out, lddy, ldx, dy, x)) {
    const int fsplits = linear_wgrad_tf32x3_splits(m, k, n_out);
    const size_t fneed = static_cast<size_t>(fsplits) * n_out * k * sizeof(float);
    if (ws_bytes < fneed) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, fneed);
    int rc = linear_wgrad_tf32x3(dy, x, static_cast<float*>(ws), m, k, n_out, lddy, ldx, st);
    if (rc != GWEN_OK) return rc;
    k_reduce_splits<<<static_cast<unsigned>(ceil_div(n_elems, 256)), 256, 0, st>>>(
        static_cast<const float*>(ws), fsplits, n_elems, n_elems, dw, lddw, k);
    GWEN_LAUNCH_CHECK("k_reduce_splits");
    return GWEN_OK;
  }
  const int splits = wgrad_splits(m, k, n_out);
  const size_t need = static_cast<size_t>(splits) * n_out * k * sizeof(float);
  if (ws_bytes < need) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  // dw[n, kk] = sum_m dy[m, n] * x[m, kk]:  A(i=n, r=m) = dy[r*lddy + i], B(j=kk, r=m) = x[r*ldx + j]
  int64_t per = ceil_div(ceil_div(m, splits), BK) * BK;
  if (per == 0) per = BK;
  GemmArgs g{dy, x, ws, nullptr, n_out, k, m, lddy, ldx, k, per, n_out * k, 0};
  int rc = dtype == GWEN_F32 ? launch_gemm<float, false, false, true>(g, splits, st)
                             : launch_gemm<__nv_bfloat16, false, false, true>(g, splits, st);
  if (rc != GWEN_OK) return rc;
  k_reduce_splits<<<static_cast<unsigned>(ceil_div(n_elems, 256)), 256, 0, st>>>(
      static_cast<const float*>(ws), splits, n_elems, n_elems, dw, lddw, k);
  GWEN_LAUNCH_CHECK("k_reduce_splits");
  return GWEN_OK;
}

extern "C" int gwen_relu_bwd(const void* y, void* dy, int64_t rows, int64_t feat, int64_t ldy,
                             int64_t lddy, int dtype, void* stream) {
  GWEN_CHECK_ARG(rows >= 0 && feat >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (rows * feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(y && dy, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks =
      static_cast<unsigned>(std::min<int64_t>(ceil_div(rows * feat, 256), int64_t(sm_count()) * 16));
  if (dtype == GWEN_F32)
    k_relu_bwd<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(y),
                                              static_cast<float*>(dy), rows, feat, ldy, lddy);
  else
    k_relu_bwd<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(y),
                                                      static_cast<__nv_bfloat16*>(dy), rows, feat,
                                                      ldy, lddy);
  GWEN_LAUNCH_CHECK("k_relu_bwd");
  return GWEN_OK;
}

extern "C" int gwen_bias_grad_workspace_bytes(int64_t rows, int64_t feat, size_t* out) {
  GWEN_CHECK_ARG(out && rows >= 0 && feat >= 0, "bad arguments");
  const int64_t chunks = std::max<int64_t>(1, ceil_div(rows, kFusedChunk));
  *out = static_cast<size_t>(chunks + reduce_groups(chunks)) * feat * sizeof(float) + 256;
  return GWEN_OK;
}

extern "C" int gwen_bias_grad(const void* dy, float* db, int64_t rows, int64_t feat, int64_t lddy,
                              int dtype, void* ws, size_t ws_bytes, void* stream) {
  GWEN_CHECK_ARG(rows >= 0 && feat >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(db && ws && (rows == 0 || dy), "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunks = std::max<int64_t>(1, ceil_div(rows, kBiasChunk));
  const size_t need = static_cast<size_t>(chunks) * feat * sizeof(float);
  if (ws_bytes < need) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  if (chunks > 65535) return set_err(GWEN_E_NOSUPPORT, "too many rows for bias_grad");
  dim3 grid(static_cast<unsigned>(ceil_div(feat, 32)), static_cast<unsigned>(chunks));
  float* part = static_cast<float*>(ws);
  if (dtype == GWEN_F32)
    k_bias_partial<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dy), rows, feat, lddy,
                                                kBiasChunk, part);
  else
    k_bias_partial<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy),
                                                        rows, feat, lddy, kBiasChunk, part);
  GWEN_LAUNCH_CHECK("k_bias_partial");
  float* tmp = ws_bytes >= size_t(chunks + reduce_groups(chunks)) * feat * sizeof(float) ? part + chunks * feat : nullptr;
  return reduce_vector(part, chunks, feat, db, tmp, st);
}

extern "C" int gwen_relu_bias_bwd(const void* y, const void* dy, void* dz, float* db, int64_t rows,
                                  int64_t feat, int dtype, void* ws, size_t ws_bytes, void* stream) {
  GWEN_CHECK_ARG(rows >= 0 && feat >= 0, "negative size");
  if (check_dtype(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG((rows == 0 || dy) && (!y || dz) && (!db || ws), "null pointer");
  const int vn = dtype == GWEN_F32 ? 4 : 8;
  const int64_t groups = feat / vn;
  if (feat % vn || groups > 256 || 256 % groups || !aligned16(dy) || !aligned16(y) || !aligned16(dz))
    return set_err(GWEN_E_NOSUPPORT, "fused relu/bias backward needs feat = %d * (a divisor of 256)", vn);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunks = std::max<int64_t>(1, ceil_div(rows, kFusedChunk));
  if (db) {
    const size_t need = static_cast<size_t>(chunks) * feat * sizeof(float);
    if (ws_bytes < need) return set_err(GWEN_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  }
  float* part = db ? static_cast<float*>(ws) : nullptr;
  if (!y && !db) return GWEN_OK;
  if (dtype == GWEN_F32)
    k_relu_bias_bwd<float><<<static_cast<unsigned>(chunks), 256, 0, st>>>(
        static_cast<const float*>(y), static_cast<const float*>(dy), static_cast<float*>(dz), rows,
        static_cast<int>(groups), kFusedChunk, part);
  else
    k_relu_bias_bwd<__nv_bfloat16><<<static_cast<unsigned>(chunks), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(dy),
        static_cast<__nv_bfloat16*>(dz), rows, static_cast<int>(groups), kFusedChunk, part);
  GWEN_LAUNCH_CHECK("k_relu_bias_bwd");
  if (db) {
    float* tmp = ws_bytes >= size_t(chunks + reduce_groups(chunks)) * feat * sizeof(float) ? part + chunks * feat : nullptr;
    return reduce_vector(part, chunks, feat, db, tmp, st);
  }
  return GWEN_OK;
}
