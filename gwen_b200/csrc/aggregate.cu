// K1: deterministic segment-reduce aggregation  out[i] = epi(sum_s w[s] * x[src[s]] + bias).
// HBM-bound: per layer call the compulsory traffic is one read of every source row, one write
// of every destination row and 8 bytes per message (DESIGN.md section "K1").
//
// Two kernels:
//   k_agg_rows   sub-warp per destination row, 128-bit read-only loads straight from L2/L1,
//                4 messages in flight per lane.  Serves any graph (also the complete-graph
//                regime of the shipped GWEN model) and is the fall-back of the tiled kernel.
//   k_agg_tiled  destination rows in tiles; the tile's distinct source-row slabs are staged once
//                in shared memory by cp.async.bulk (UBLKCP) completing on an mbarrier, then
//                warps reduce out of shared memory with LDS.128.  L2->SM traffic drops from
//                deg x to (distinct sources / tile rows) x; 2+ CTAs per SM overlap one CTA's
//                staging with another's reduction.
// Both accumulate in fp32 in CSR order with unfused mul.rn/add.rn (bitwise == CPU scatter_add_).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gwen {
namespace {

// ---- 16-byte vectors of T ---------------------------------------------------------------------
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x);
    f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z);
    f[3] = __uint_as_float(r.w);
  }
  __device__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& r, float* f) {
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(u[i] << 16);
      f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  __device__ static uint4 pack(const float* f) {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      u[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    return make_uint4(u[0], u[1], u[2], u[3]);
  }
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(saddr));
  return r;
}

__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

template <int N>
__device__ __forceinline__ void axpy_exact(float* acc, float w, const float* v) {
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(w, v[k]));
}

template <typename T, int N>
__device__ __forceinline__ uint4 finish(float* acc, const float* __restrict__ bias, int64_t col,
                                        int relu) {
  if (bias) {
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], __ldg(bias + col + k));
  }
  if (relu) {
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = fmaxf(acc[k], 0.0f);
  }
  return Vec16<T>::pack(acc);
}

struct AggArgs {
  const int32_t* rowptr;
  const int32_t* src;
  const float* w;
  const int32_t* order;
  const void* x;
  void* out;
  const float* bias;
  int64_t batch, n_dst, feat, ldx, x_bstride, ldo, o_bstride;
  int relu;
};

// ---- k_agg_rows -------------------------------------------------------------------------------
// LPR lanes share one destination row; each lane owns CH 16-byte column chunks, LPR chunks apart.
template <typename T, int LPR, int CH>
__global__ void __launch_bounds__(256) k_agg_rows(AggArgs a, int64_t row_groups, int col_blocks) {
  constexpr int VN = Vec16<T>::N;
  constexpr int RPW = 32 / LPR;  // destination rows per warp
  constexpr int U = 4;           // messages in flight
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int64_t items = a.batch * col_blocks * row_groups;
  if (warp >= items) return;
  const int64_t rg = warp % row_groups;
  const int cb = static_cast<int>((warp / row_groups) % col_blocks);
  const int64_t b = warp / (row_groups * col_blocks);
  const int64_t pos = rg * RPW + sub;
  if (pos >= a.n_dst) return;
  const int64_t d = a.order ? a.order[pos] : pos;

  int64_t col[CH];
  bool on[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    col[c] = (int64_t(cb) * CH * LPR + c * LPR + l) * VN;
    on[c] = col[c] < a.feat;
  }
  const T* xb = static_cast<const T*>(a.x) + b * a.x_bstride;
  float acc[CH][VN];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[c][k] = 0.0f;

  const int beg = __ldg(a.rowptr + d), end = __ldg(a.rowptr + d + 1);
  int e = beg;
  for (; e + U <= end; e += U) {
    int s[U];
    float ww[U];
    uint4 v[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] = __ldg(a.src + e + u);
      ww[u] = __ldg(a.w + e + u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (on[c]) v[u][c] = ldg_nc_v4(xb + int64_t(s[u]) * a.ldx + col[c]);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (on[c]) {
          float f[VN];
          Vec16<T>::unpack(v[u][c], f);
          axpy_exact<VN>(acc[c], ww[u], f);
        }
  }
  for (; e < end; ++e) {
    const int s = __ldg(a.src + e);
    const float ww = __ldg(a.w + e);
    uint4 v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (on[c]) v[c] = ldg_nc_v4(xb + int64_t(s) * a.ldx + col[c]);
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (on[c]) {
        float f[VN];
        Vec16<T>::unpack(v[c], f);
        axpy_exact<VN>(acc[c], ww, f);
      }
  }
  T* ob = static_cast<T*>(a.out) + b * a.o_bstride + d * a.ldo;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    if (on[c]) stg_v4(ob + col[c], finish<T, VN>(acc[c], a.bias, col[c], a.relu));
}

// Scalar fall-back for feature counts / pitches that are not 16-byte multiples: warp per row.
template <typename T>
__global__ void __launch_bounds__(256) k_agg_rows_scalar(AggArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (warp >= a.batch * a.n_dst) return;
  const int64_t b = warp / a.n_dst, pos = warp % a.n_dst;
  const int64_t d = a.order ? a.order[pos] : pos;
  const T* xb = static_cast<const T*>(a.x) + b * a.x_bstride;
  T* ob = static_cast<T*>(a.out) + b * a.o_bstride + d * a.ldo;
  const int beg = a.rowptr[d], end = a.rowptr[d + 1];
  for (int64_t c = lane; c < a.feat; c += 32) {
    float acc = 0.0f;
    for (int e = beg; e < end; ++e)
      acc = __fadd_rn(acc, __fmul_rn(a.w[e], to_f32(xb[int64_t(a.src[e]) * a.ldx + c])));
    if (a.bias) acc = __fadd_rn(acc, a.bias[c]);
    if (a.relu) acc = fmaxf(acc, 0.0f);
    ob[c] = from_f32<T>(acc);
  }
}

template <typename T, int LPR, int CH>
int launch_rows(const AggArgs& a, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  constexpr int RPW = 32 / LPR;
  const int64_t row_groups = ceil_div(a.n_dst, RPW);
  const int col_blocks = static_cast<int>(ceil_div(a.feat, int64_t(LPR) * CH * VN));
  const int64_t warps = a.batch * col_blocks * row_groups;
  const int64_t blocks = ceil_div(warps, 8);
  if (blocks > INT32_MAX) return set_err(GWEN_E_NOSUPPORT, "aggregate grid too large");
  k_agg_rows<T, LPR, CH><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, row_groups, col_blocks);
  GWEN_LAUNCH_CHECK("k_agg_rows");
  return GWEN_OK;
}

template <typename T>
int dispatch_rows(const AggArgs& a, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const bool vec_ok = a.feat % VN == 0 && a.ldx % VN == 0 && a.ldo % VN == 0 &&
                      a.x_bstride % VN == 0 && a.o_bstride % VN == 0 && aligned16(a.x) &&
                      aligned16(a.out);
  if (!vec_ok) {
    const int64_t blocks = ceil_div(a.batch * a.n_dst, 8);
    if (blocks > INT32_MAX) return set_err(GWEN_E_NOSUPPORT, "aggregate grid too large");
    k_agg_rows_scalar<T><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
    GWEN_LAUNCH_CHECK("k_agg_rows_scalar");
    return GWEN_OK;
  }
  const int64_t nv = a.feat / VN;  // 16-byte chunks per row
  if (nv <= 8) return launch_rows<T, 8, 1>(a, st);
  if (nv <= 16) return launch_rows<T, 16, 1>(a, st);
  if (nv <= 32) return launch_rows<T, 32, 1>(a, st);
  if (nv <= 64) return launch_rows<T, 32, 2>(a, st);
  return launch_rows<T, 32, 4>(a, st);
}

// ---- k_agg_tiled ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// One bulk asynchronous copy global -> shared, completing `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

struct TiledArgs {
  const int32_t* rowptr;
  const int32_t* order;
  const int32_t* tile_ptr;
  const int32_t* tsrc_ptr;
  const int32_t* tsrc;
  const uint64_t* msg;
  const void* x;
  void* out;
  const float* bias;
  int64_t batch, feat, ldx, x_bstride, ldo, o_bstride;
  int num_tiles, slabs, slab_elems, relu;
};

// Work item = (batch b, feature slab, tile).  A slab is CH*32 16-byte chunks wide at most.
template <typename T, int CH>
__global__ void __launch_bounds__(256) k_agg_tiled(TiledArgs a) {
  constexpr int VN = Vec16<T>::N;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_storage;
  const uint32_t bar = smem_u32(&bar_storage);
  const uint32_t stage = smem_u32(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint32_t parity = 0;
  const int64_t items = a.batch * a.slabs * int64_t(a.num_tiles);
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int t = static_cast<int>(it % a.num_tiles);
    const int slab = static_cast<int>((it / a.num_tiles) % a.slabs);
    const int64_t b = it / (int64_t(a.num_tiles) * a.slabs);
    const int64_t f0 = int64_t(slab) * a.slab_elems;
    const int cols = static_cast<int>(imin64(a.slab_elems, a.feat - f0));
    const uint32_t row_bytes = cols * sizeof(T);
    const int s_beg = __ldg(a.tsrc_ptr + t), s_end = __ldg(a.tsrc_ptr + t + 1);
    const T* xb = static_cast<const T*>(a.x) + b * a.x_bstride + f0;
    // stage: one bulk copy per distinct source row slab
    if (threadIdx.x == 0) mbar_expect_tx(bar, uint32_t(s_end - s_beg) * row_bytes);
    for (int i = s_beg + threadIdx.x; i < s_end; i += blockDim.x)
      bulk_g2s(stage + uint32_t(i - s_beg) * row_bytes, xb + int64_t(__ldg(a.tsrc + i)) * a.ldx,
               row_bytes, bar);
    mbar_wait(bar, parity);
    parity ^= 1;
    // reduce: warp per destination row of the tile
    const int p_beg = __ldg(a.tile_ptr + t), p_end = __ldg(a.tile_ptr + t + 1);
    const int nchunk = cols / VN;  // 16-byte chunks in this slab
    for (int p = p_beg + warp; p < p_end; p += nwarps) {
      const int64_t d = a.order ? __ldg(a.order + p) : p;
      const int beg = __ldg(a.rowptr + d), end = __ldg(a.rowptr + d + 1);
      float acc[CH][VN];
      bool on[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        on[c] = c * 32 + lane < nchunk;
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[c][k] = 0.0f;
      }
      const uint32_t lane_off = lane * 16;
#pragma unroll 3
      for (int e = beg; e < end; ++e) {
        const uint64_t m = __ldg(reinterpret_cast<const unsigned long long*>(a.msg) + e);
        const uint32_t li = static_cast<uint32_t>(m);
        const float ww = __uint_as_float(static_cast<uint32_t>(m >> 32));
        const uint32_t base = stage + li * row_bytes + lane_off;
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (on[c]) {
            float f[VN];
            Vec16<T>::unpack(lds_v4(base + c * 512), f);
            axpy_exact<VN>(acc[c], ww, f);
          }
      }
      T* ob = static_cast<T*>(a.out) + b * a.o_bstride + d * a.ldo + f0;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (on[c]) {
          const int64_t cc = int64_t(c * 32 + lane) * VN;
          stg_v4(ob + cc, finish<T, VN>(acc[c], a.bias, f0 + cc, a.relu));
        }
    }
    __syncthreads();  // every warp is done with the stage before the next item overwrites it
  }
}

template <typename T, int CH>
int launch_tiled(const TiledArgs& a, size_t smem_bytes, cudaStream_t st) {
  auto kern = k_agg_tiled<T, CH>;
  GWEN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem_bytes)));
  int occ = 0;
  GWEN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem_bytes));
  if (occ < 1) return set_err(GWEN_E_NOSUPPORT, "tiled aggregate does not fit on an SM");
  const int64_t items = a.batch * a.slabs * int64_t(a.num_tiles);
  const int64_t grid = std::min<int64_t>(items, int64_t(sm_count()) * occ);
  kern<<<static_cast<unsigned>(grid), 256, smem_bytes, st>>>(a);
  GWEN_LAUNCH_CHECK("k_agg_tiled");
  return GWEN_OK;
}

// ---- tile plan --------------------------------------------------------------------------------
constexpr int kThreads = 256;

struct PlanWs {
  size_t off_keys_in, off_keys_out, off_vals_in, off_vals_out, off_rank, off_cub, cub_bytes, bytes;
  int bits;
};

cudaError_t plan_ws(int64_t n_dst, int64_t m, int64_t num_tiles, PlanWs* p) {
  (void)n_dst;
  size_t t = static_cast<size_t>(m > 0 ? m : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p->off_keys_in = take(t * 8);
  p->off_keys_out = take(t * 8);
  p->off_vals_in = take(t * 4);
  p->off_vals_out = take(t * 4);
  p->off_rank = take(t * 4);
  int tb = 1;
  while ((int64_t(1) << tb) < num_tiles + 1) ++tb;
  p->bits = 32 + tb;
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t err = cub::DeviceRadixSort::SortPairs(
      nullptr, sort_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
      static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(t),
      0, p->bits);
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(t));
  if (err != cudaSuccess) return err;
  p->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  p->off_cub = take(p->cub_bytes);
  p->bytes = off;
  return cudaSuccess;
}

// thread per position p of the processing order: key = (tile << 32) | src for each message
__global__ void k_plan_keys(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                            const int32_t* __restrict__ order,
                            const int32_t* __restrict__ tile_ptr, int num_tiles, int64_t n_dst,
                            uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p >= n_dst) return;
  int lo = 0, hi = num_tiles;  // last tile with tile_ptr[t] <= p
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile_ptr[mid] <= p) lo = mid; else hi = mid;
  }
  const int64_t d = order ? order[p] : p;
  for (int s = rowptr[d]; s < rowptr[d + 1]; ++s) {
    keys[s] = (uint64_t(uint32_t(lo)) << 32) | uint32_t(src[s]);
    vals[s] = s;
  }
}

__global__ void k_plan_heads(const uint64_t* __restrict__ keys, int64_t m,
                             int32_t* __restrict__ rank) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  rank[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void k_plan_tsrc(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rank,
                            int64_t m, int num_tiles, int32_t* __restrict__ tsrc_ptr,
                            int32_t* __restrict__ tsrc, int32_t* __restrict__ status) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  const uint64_t k = keys[i];
  const int t = static_cast<int>(k >> 32);
  const int u = rank[i] - 1;
  const bool head = i == 0 || k != keys[i - 1];
  if (head) tsrc[u] = static_cast<int32_t>(uint32_t(k));
  const int tp = i ? static_cast<int>(keys[i - 1] >> 32) : -1;
  for (int tt = tp + 1; tt <= t; ++tt) tsrc_ptr[tt] = u;
  if (i == m - 1) {
    for (int tt = t + 1; tt <= num_tiles; ++tt) tsrc_ptr[tt] = u + 1;
    status[0] = u + 1;
  }
}

__global__ void k_plan_msg(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                           const int32_t* __restrict__ rank,
                           const int32_t* __restrict__ tsrc_ptr, const float* __restrict__ w,
                           int64_t m, uint64_t* __restrict__ msg) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  const int t = static_cast<int>(keys[i] >> 32);
  const uint32_t li = static_cast<uint32_t>(rank[i] - 1 - tsrc_ptr[t]);
  const int32_t slot = vals[i];
  msg[slot] = (uint64_t(__float_as_uint(w[slot])) << 32) | li;
}

__global__ void k_plan_max(const int32_t* __restrict__ tsrc_ptr, int num_tiles,
                           int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  atomicMax(&status[1], tsrc_ptr[t + 1] - tsrc_ptr[t]);
}

__global__ void k_uniform_tiles(int64_t n_dst, int tile_rows, int64_t num_tiles,
                                int32_t* __restrict__ tile_ptr) {
  int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t > num_tiles) return;
  tile_ptr[t] = static_cast<int32_t>(imin64(t * tile_rows, n_dst));
}

__global__ void k_grid_tiles(int64_t h, int64_t w, int th, int tw, int32_t* __restrict__ order,
                             int32_t* __restrict__ tile_ptr) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t tcols = (w + tw - 1) / tw, trows = (h + th - 1) / th;
  if (i <= trows * tcols) {  // tile_ptr: nodes before tile (ti, tj)
    if (i == trows * tcols) {
      tile_ptr[i] = static_cast<int32_t>(h * w);
    } else {
      const int64_t ti = i / tcols, tj = i % tcols;
      const int64_t rows_here = imin64(th, h - ti * th);
      tile_ptr[i] = static_cast<int32_t>(ti * th * w + rows_here * tj * tw);
    }
  }
  if (i >= h * w) return;
  const int64_t r = i / w, c = i % w;
  const int64_t ti = r / th, tj = c / tw;
  const int64_t rows_here = imin64(th, h - ti * th);
  const int64_t cols_here = imin64(tw, w - tj * tw);
  const int64_t p = ti * th * w + rows_here * tj * tw + (r - ti * th) * cols_here + (c - tj * tw);
  order[p] = static_cast<int32_t>(i);
}

// ---- row gather / scatter (halo pack / unpack) ------------------------------------------------
template <typename T, bool kGather>
__global__ void __launch_bounds__(256) k_rows_move(const T* __restrict__ from, T* __restrict__ to,
                                                   const int32_t* __restrict__ idx, int64_t batch,
                                                   int64_t n_idx, int64_t feat, int64_t ldx,
                                                   int64_t x_bstride, bool vec) {
  // `from`/`to`: for gather from = x, to = buf; for scatter from = buf, to = x.
  constexpr int VN = Vec16<T>::N;
  const int64_t per_row = vec ? feat / VN : feat;
  const int64_t total = batch * n_idx * per_row;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t c = i % per_row, j = (i / per_row) % n_idx, b = i / (per_row * n_idx);
    const int64_t xoff = b * x_bstride + int64_t(idx[j]) * ldx;
    const int64_t boff = (b * n_idx + j) * feat;
    if (vec) {
      const uint4* s = reinterpret_cast<const uint4*>(from + (kGather ? xoff : boff)) + c;
      uint4* d = reinterpret_cast<uint4*>(to + (kGather ? boff : xoff)) + c;
      *d = *s;
    } else {
      to[(kGather ? boff : xoff) + c] = from[(kGather ? xoff : boff) + c];
    }
  }
}

template <typename T, bool kGather>
int rows_move(const void* from, void* to, const int32_t* idx, int64_t batch, int64_t n_idx,
              int64_t feat, int64_t ldx, int64_t x_bstride, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  if (batch * n_idx * feat == 0) return GWEN_OK;
  const bool vec = feat % VN == 0 && ldx % VN == 0 && x_bstride % VN == 0 && aligned16(from) &&
                   aligned16(to);
  const int64_t total = batch * n_idx * (vec ? feat / VN : feat);
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), int64_t(sm_count()) * 8);
  k_rows_move<T, kGather><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
      static_cast<const T*>(from), static_cast<T*>(to), idx, batch, n_idx, feat, ldx, x_bstride,
      vec);
  GWEN_LAUNCH_CHECK("k_rows_move");
  return GWEN_OK;
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_aggregate_fwd(const int32_t* rowptr, const int32_t* src, const float* w,
                                  const int32_t* order, const void* x, void* out, int64_t batch,
                                  int64_t n_dst, int64_t n_src, int64_t feat, int64_t ldx,
                                  int64_t x_bstride, int64_t ldo, int64_t o_bstride, int dtype,
                                  const float* bias, int epilogue, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && n_dst >= 0 && n_src >= 0 && feat >= 0, "negative size");
  if (batch == 0 || n_dst == 0 || feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(rowptr && src && w && x && out, "null pointer");
  GWEN_CHECK_ARG(ldx >= feat && ldo >= feat, "row pitch smaller than feature count");
  AggArgs a{rowptr, src, w, order, x, out, bias, batch, n_dst, feat, ldx, x_bstride, ldo,
            o_bstride, (epilogue & GWEN_EPI_RELU) ? 1 : 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GWEN_F32) return dispatch_rows<float>(a, st);
  if (dtype == GWEN_BF16) return dispatch_rows<__nv_bfloat16>(a, st);
  return set_err(GWEN_E_DTYPE, "unknown dtype %d", dtype);
}

extern "C" int gwen_tile_plan_workspace_bytes(int64_t n_dst, int64_t m, int64_t num_tiles,
                                              size_t* out) {
  GWEN_CHECK_ARG(out && n_dst >= 0 && m >= 0 && num_tiles >= 0, "bad arguments");
  GWEN_CHECK_ARG(m < INT32_MAX, "too many messages");
  PlanWs p;
  GWEN_CUDA(plan_ws(n_dst, m, num_tiles, &p));
  *out = p.bytes;
  return GWEN_OK;
}

extern "C" int gwen_tile_plan_build(const int32_t* rowptr, const int32_t* src, const float* w,
                                    const int32_t* order, const int32_t* tile_ptr,
                                    int64_t num_tiles, int64_t n_dst, int64_t m,
                                    int32_t* tsrc_ptr, int32_t* tsrc, uint64_t* msg,
                                    int32_t* status, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GWEN_CHECK_ARG(rowptr && src && w && tile_ptr && tsrc_ptr && tsrc && msg && status && ws,
                 "null pointer");
  GWEN_CHECK_ARG(num_tiles > 0 && num_tiles < INT32_MAX && n_dst > 0 && m > 0 && m < INT32_MAX,
                 "bad sizes");
  PlanWs p;
  GWEN_CUDA(plan_ws(n_dst, m, num_tiles, &p));
  if (ws_bytes < p.bytes)
    return set_err(GWEN_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes, p.bytes);
  char* base = static_cast<char*>(ws);
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(base + p.off_keys_in);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(base + p.off_keys_out);
  int32_t* vals_in = reinterpret_cast<int32_t*>(base + p.off_vals_in);
  int32_t* vals_out = reinterpret_cast<int32_t*>(base + p.off_vals_out);
  int32_t* rank = reinterpret_cast<int32_t*>(base + p.off_rank);
  void* cub_ws = base + p.off_cub;
  GWEN_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  const unsigned grid_n = static_cast<unsigned>(ceil_div(n_dst, kThreads));
  const unsigned grid_m = static_cast<unsigned>(ceil_div(m, kThreads));
  k_plan_keys<<<grid_n, kThreads, 0, st>>>(rowptr, src, order, tile_ptr,
                                           static_cast<int>(num_tiles), n_dst, keys_in, vals_in);
  GWEN_LAUNCH_CHECK("k_plan_keys");
  size_t cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, vals_in,
                                            vals_out, static_cast<int>(m), 0, p.bits, st));
  k_plan_heads<<<grid_m, kThreads, 0, st>>>(keys_out, m, rank);
  GWEN_LAUNCH_CHECK("k_plan_heads");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes, rank, rank, static_cast<int>(m), st));
  k_plan_tsrc<<<grid_m, kThreads, 0, st>>>(keys_out, rank, m, static_cast<int>(num_tiles),
                                           tsrc_ptr, tsrc, status);
  GWEN_LAUNCH_CHECK("k_plan_tsrc");
  k_plan_msg<<<grid_m, kThreads, 0, st>>>(keys_out, vals_out, rank, tsrc_ptr, w, m, msg);
  GWEN_LAUNCH_CHECK("k_plan_msg");
  k_plan_max<<<static_cast<unsigned>(ceil_div(num_tiles, kThreads)), kThreads, 0, st>>>(
      tsrc_ptr, static_cast<int>(num_tiles), status);
  GWEN_LAUNCH_CHECK("k_plan_max");
  return GWEN_OK;
}

extern "C" int gwen_uniform_tiles(int64_t n_dst, int32_t tile_rows, int32_t* tile_ptr,
                                  void* stream) {
  GWEN_CHECK_ARG(n_dst >= 0 && tile_rows > 0 && tile_ptr, "bad arguments");
  const int64_t nt = ceil_div(n_dst, tile_rows);
  k_uniform_tiles<<<static_cast<unsigned>(ceil_div(nt + 1, kThreads)), kThreads, 0,
                    static_cast<cudaStream_t>(stream)>>>(n_dst, tile_rows, nt, tile_ptr);
  GWEN_LAUNCH_CHECK("k_uniform_tiles");
  return GWEN_OK;
}

extern "C" int gwen_grid_tiles(int64_t h, int64_t w, int32_t th, int32_t tw, int32_t* order,
                               int32_t* tile_ptr, void* stream) {
  GWEN_CHECK_ARG(h > 0 && w > 0 && th > 0 && tw > 0 && order && tile_ptr, "bad arguments");
  GWEN_CHECK_ARG(h * w < INT32_MAX, "grid too large");
  const int64_t tiles = ceil_div(h, th) * ceil_div(w, tw);
  const int64_t threads = std::max<int64_t>(h * w, tiles + 1);
  k_grid_tiles<<<static_cast<unsigned>(ceil_div(threads, kThreads)), kThreads, 0,
                 static_cast<cudaStream_t>(stream)>>>(h, w, th, tw, order, tile_ptr);
  GWEN_LAUNCH_CHECK("k_grid_tiles");
  return GWEN_OK;
}

extern "C" int gwen_aggregate_tiled_fwd(const gwen_tile_plan* plan, const int32_t* rowptr,
                                        const void* x, void* out, int64_t batch, int64_t n_src,
                                        int64_t feat, int64_t ldx, int64_t x_bstride, int64_t ldo,
                                        int64_t o_bstride, int dtype, const float* bias,
                                        int epilogue, int32_t slab_elems, void* stream) {
  GWEN_CHECK_ARG(plan && rowptr && x && out, "null pointer");
  GWEN_CHECK_ARG(batch >= 0 && feat >= 0 && n_src >= 0, "negative size");
  if (batch == 0 || feat == 0 || plan->n_dst == 0) return GWEN_OK;
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  GWEN_CHECK_ARG(plan->tile_ptr && plan->tsrc_ptr && plan->tsrc && plan->msg, "incomplete plan");
  const int esz = dtype == GWEN_F32 ? 4 : 2;
  const int vn = 16 / esz;
  if (feat % vn || ldx % vn || ldo % vn || x_bstride % vn || o_bstride % vn || !aligned16(x) ||
      !aligned16(out))
    return set_err(GWEN_E_ALIGN, "tiled aggregate needs 16-byte aligned rows (feat %% %d == 0)", vn);
  // slab: whole 16-byte chunks, at most 64 per row (CH <= 2), and the stage must fit in smem.
  const int64_t max_stage = 200 * 1024;
  int64_t slab = slab_elems > 0 ? slab_elems : 32 * vn;
  slab = std::min<int64_t>(slab, feat);
  slab = std::min<int64_t>(slab, 64 * vn);
  slab -= slab % vn;
  while (slab > vn && int64_t(plan->max_tile_src) * slab * esz > max_stage) slab -= vn;
  if (slab <= 0 || int64_t(plan->max_tile_src) * slab * esz > max_stage)
    return set_err(GWEN_E_NOSUPPORT, "tile with %d sources does not fit in shared memory",
                   plan->max_tile_src);
  TiledArgs a{rowptr, plan->order, plan->tile_ptr, plan->tsrc_ptr, plan->tsrc, plan->msg, x, out,
              bias, batch, feat, ldx, x_bstride, ldo, o_bstride, plan->num_tiles,
              static_cast<int>(ceil_div(feat, slab)), static_cast<int>(slab),
              (epilogue & GWEN_EPI_RELU) ? 1 : 0};
  const size_t smem = static_cast<size_t>(plan->max_tile_src) * slab * esz;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool two = slab > 32 * vn;
  if (dtype == GWEN_F32)
    return two ? launch_tiled<float, 2>(a, smem, st) : launch_tiled<float, 1>(a, smem, st);
  return two ? launch_tiled<__nv_bfloat16, 2>(a, smem, st)
             : launch_tiled<__nv_bfloat16, 1>(a, smem, st);
}

static int dtype_ok(int dtype) {
  if (dtype == GWEN_F32 || dtype == GWEN_BF16) return GWEN_OK;
  return set_err(GWEN_E_DTYPE, "unknown dtype %d", dtype);
}

extern "C" int gwen_rows_gather(const void* x, const int32_t* idx, void* buf, int64_t batch,
                                int64_t n_idx, int64_t feat, int64_t ldx, int64_t x_bstride,
                                int dtype, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && n_idx >= 0 && feat >= 0, "negative size");
  if (dtype_ok(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (batch * n_idx * feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && idx && buf, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == GWEN_F32
             ? rows_move<float, true>(x, buf, idx, batch, n_idx, feat, ldx, x_bstride, st)
             : rows_move<__nv_bfloat16, true>(x, buf, idx, batch, n_idx, feat, ldx, x_bstride, st);
}

extern "C" int gwen_rows_scatter(const void* buf, const int32_t* idx, void* x, int64_t batch,
                                 int64_t n_idx, int64_t feat, int64_t ldx, int64_t x_bstride,
                                 int dtype, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && n_idx >= 0 && feat >= 0, "negative size");
  if (dtype_ok(dtype) != GWEN_OK) return GWEN_E_DTYPE;
  if (batch * n_idx * feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && idx && buf, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == GWEN_F32
             ? rows_move<float, false>(buf, x, idx, batch, n_idx, feat, ldx, x_bstride, st)
             : rows_move<__nv_bfloat16, false>(buf, x, idx, batch, n_idx, feat, ldx, x_bstride,
                                               st);
}
