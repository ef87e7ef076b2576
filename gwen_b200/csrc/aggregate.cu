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
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "tma.cuh"

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

// fp32 data: unfused mul.rn + add.rn, so the sum equals the CPU scatter_add_ order bitwise.
// bf16 data: the result is rounded to bf16 anyway (tolerance-checked), so one FFMA per element.
template <typename T, int N>
__device__ __forceinline__ void axpy_exact(float* acc, float w, const float* v) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if constexpr (sizeof(T) == 4) acc[k] = __fadd_rn(acc[k], __fmul_rn(w, v[k]));
    else acc[k] = fmaf(w, v[k], acc[k]);
  }
}

template <typename T, int N>
__device__ __forceinline__ uint4 finish(float* acc, const float* __restrict__ bias, int64_t col,
                                        int relu) {
  if (bias) {
    float bv[N];
    if constexpr (N == 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
    } else {
#pragma unroll
      for (int k = 0; k < N; k += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col + k));
        bv[k] = b4.x; bv[k + 1] = b4.y; bv[k + 2] = b4.z; bv[k + 3] = b4.w;
      }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], bv[k]);
  }
  if (relu) {
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = fmaxf(acc[k], 0.0f);
  }
  return Vec16<T>::pack(acc);
}

// Same epilogue with the bias already in registers (hoisted out of the row loop).
template <typename T, int N>
__device__ __forceinline__ uint4 finish_reg(float* acc, const float* bv, bool has_bias, int relu) {
  if (has_bias) {
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = __fadd_rn(acc[k], bv[k]);
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
          axpy_exact<T, VN>(acc[c], ww[u], f);
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
        axpy_exact<T, VN>(acc[c], ww, f);
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
struct TiledArgs {
  const int32_t* tile_ptr;
  const int32_t* run_ptr;
  const int32_t* run_start;
  const int4* trec;       // per position p: {dst row, message offset inside the tile, degree, 0}
  const uint64_t* tmsg;   // messages in processing order: low 32 = staged row, high 32 = weight
  const int32_t* tmsg_base;
  void* out;
  const float* bias;
  int64_t batch, feat, ldo, o_bstride;
  int num_tiles, tile_begin, slabs, slab_elems, run_len, relu;
  uint32_t stage_bytes;  // one data stage = max_tile_runs * run_len * slab bytes
  uint32_t rec_bytes;    // one metadata buffer: records, then messages
  uint32_t meta_bytes;
  int num_stages;
  // gather mode (plans whose runs are single rows: locality tiles of an arbitrarily numbered graph):
  // the producer warp copies row slabs with 16-byte cp.async instead of one TMA box per run
  int gather;            // 0: TMA boxes per run; 1: 16-byte cp.async per row slab; 2: one bulk copy per row slab
  int gather_warps;      // producer warps of gather mode 1
  const void* x;
  int64_t ldx, x_bstride;
  uint32_t runs_off;     // offset of the tile's run list inside a metadata buffer (gather mode)
};

__device__ __forceinline__ int4 lds_i4(uint32_t saddr) {
  int4 r;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(saddr));
  return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t saddr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
  return r;
}
// Plain bulk copy global -> shared (contiguous bytes, multiple of 16), completing on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

__device__ __forceinline__ int lds_s32(uint32_t saddr) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(saddr));
  return r;
}
// 16 bytes global -> shared past L1 (LDGSTS.128); completion is collected by cp_async_mbar_arrive
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the executing thread's arrival on `bar` happens when all of its earlier cp.async copies have landed;
// .noinc: the arrival is one of the barrier's expected count
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Up to U messages of one destination row out of shared memory: message words (staged row, weight), then the
// source slabs (LDS.128), then the fp32 accumulation in message order.
template <typename T, int U>
__device__ __forceinline__ void reduce_batch(float* acc, uint32_t mp, uint32_t stage, uint32_t slab_bytes, int n) {
  constexpr int VN = Vec16<T>::N;
  uint2 m[U];
  uint4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n) m[u] = lds_u2(mp + u * 8u);
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n) v[u] = lds_v4(stage + m[u].x * slab_bytes);
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n) {
      float f[VN];
      Vec16<T>::unpack(v[u], f);
      axpy_exact<T, VN>(acc, __uint_as_float(m[u].y), f);
    }
}

// Persistent, warp-specialised CTA; work item = (tile, batch b, feature slab), tile-major.
// Shared memory holds NS data stages and two metadata buffers (row records + messages of a tile).
//   producer (last warp, one lane): for every item waits for its stage to be empty, then issues
//     one TMA tensor copy per run of consecutive source rows (completing on full[stage]); when
//     the item starts a new tile it first brings the tile's records and message list with two
//     plain bulk copies (meta_full).
//   consumers (all other warps): wait full[stage]; each sub-warp of LPR lanes reduces destination
//     rows out of shared memory -- record -> messages (broadcast LDS.64) -> source slabs
//     (LDS.128) -> fp32 accumulate in CSR order -> STG.128 -- then one lane per warp arrives on
//     empty[stage].  No block-wide barrier: a fast warp runs up to NS-1 items ahead.
template <typename T, int LPR, int U>  // U = messages in flight per row
__global__ void __launch_bounds__(512, 1) k_agg_tiled(const __grid_constant__ CUtensorMap xmap,
                                                      TiledArgs a) {
  constexpr int VN = Vec16<T>::N;
  constexpr int RPW = 32 / LPR;
  constexpr int MAXS = 8;   // most data stages
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAXS], empty_bar[MAXS], mfull_bar[2], mempty_bar[2];
  const int ns = a.num_stages;
  const uint32_t stage0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t meta0 = stage0 + uint32_t(ns) * a.stage_bytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int npw = a.gather == 1 ? a.gather_warps : 1;  // producer warps
  const int ncw = nwarps - npw;                         // consumer warps
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    for (int i = 0; i < ns; ++i) {
      mbar_init(smem_u32(&full_bar[i]), a.gather == 1 ? 32 * npw : 1);
      mbar_init(smem_u32(&empty_bar[i]), ncw);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&mfull_bar[i]), 1);
      mbar_init(smem_u32(&mempty_bar[i]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t slab_bytes = a.slab_elems * sizeof(T);
  const uint32_t run_bytes = a.run_len * slab_bytes;
  const int per_tile = static_cast<int>(a.batch) * a.slabs;
  const int my_tiles = a.num_tiles > int(blockIdx.x)
                           ? (a.num_tiles - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
  const int64_t n_items = int64_t(my_tiles) * per_tile;

  if (warp >= ncw && a.gather) {
    // ===== producers, gather mode: single source rows instead of TMA boxes =====
    // Per tile the run list (one source row id per staged row) arrives with the records and messages (producer
    // warp 0); per item the tile's row slabs are copied either
    //   gather == 1: with 16-byte cp.async (LDGSTS), CPR lanes per row slab, rows dealt round-robin to the
    //                producer warps; every lane arrives on full[stage] through cp.async.mbarrier.arrive.noinc
    //                (the barrier expects 32 x producer warps arrivals), or
    //   gather == 2: with one cp.async.bulk per row slab (UBLKCP), a row per lane, completing on full[stage]
    //                by byte count (one producer warp).
    const int pw = warp - ncw;
    const int cpr = int(slab_bytes >> 4);          // 16-byte chunks per row slab: 8, 16 or 32
    const int rpp = 32 / cpr;                      // rows per pass of one warp
    const int lrow = lane / cpr, lchunk = lane % cpr;
    const char* xbase = static_cast<const char*>(a.x);
    const uint32_t pitch = uint32_t(a.ldx * int64_t(sizeof(T)));   // < 4 GB (checked by the launcher)
    int st = 0;
    uint32_t st_round = 0;
    int seq = 0, rem = 0, b = 0, slab = 0, r0 = 0, r1 = 0;
    for (int64_t it = 0; it < n_items; ++it) {
      const int t = a.tile_begin + int(blockIdx.x) + seq * int(gridDim.x);
      const int mb_i = seq & 1;
      if (rem == 0) {
        r0 = __ldg(a.run_ptr + t);
        r1 = __ldg(a.run_ptr + t + 1);
      }
      const int r0a = r0 & ~3;
      if (rem == 0) {  // first item of a tile: its records, messages and run list
        if (pw == 0 && lane == 0) {
          if (seq >= 2) mbar_wait(smem_u32(&mempty_bar[mb_i]), uint32_t((seq >> 1) - 1) & 1u);
          const int p0 = __ldg(a.tile_ptr + t), p1 = __ldg(a.tile_ptr + t + 1);
          const int m0 = __ldg(a.tmsg_base + t) & ~1, m1 = __ldg(a.tmsg_base + t + 1);
          const uint32_t rb = uint32_t(p1 - p0) * 16u, mb = uint32_t((m1 - m0 + 1) & ~1) * 8u;
          const uint32_t lb = uint32_t((r1 - r0a + 3) & ~3) * 4u;
          const uint32_t bar = smem_u32(&mfull_bar[mb_i]);
          const uint32_t dst = meta0 + uint32_t(mb_i) * a.meta_bytes;
          mbar_expect_tx(bar, rb + mb + lb);
          bulk_g2s(dst, a.trec + p0, rb, bar);
          if (mb) bulk_g2s(dst + a.rec_bytes, a.tmsg + m0, mb, bar);
          if (lb) bulk_g2s(dst + a.runs_off, a.run_start + r0a, lb, bar);
        }
        mbar_wait(smem_u32(&mfull_bar[mb_i]), uint32_t(seq >> 1) & 1u);
      }
      if (st_round > 0) mbar_wait(smem_u32(&empty_bar[st]), (st_round - 1) & 1u);
      const uint32_t runs = meta0 + uint32_t(mb_i) * a.meta_bytes + a.runs_off + uint32_t(r0 - r0a) * 4u;
      const uint32_t sbase = stage0 + uint32_t(st) * a.stage_bytes;
      const int n_rows = r1 - r0;
      const int64_t col0 = int64_t(slab) * a.slab_elems;
      if (a.gather == 2) {
        const int64_t left = (a.feat - col0) * int64_t(sizeof(T));
        const uint32_t row_bytes = left < int64_t(slab_bytes) ? uint32_t(left) : slab_bytes;
        const uint32_t bar = smem_u32(&full_bar[st]);
        if (lane == 0) mbar_expect_tx(bar, uint32_t(n_rows) * row_bytes);
        __syncwarp();
        const char* xb = xbase + (int64_t(b) * a.x_bstride + col0) * int64_t(sizeof(T));
        for (int i = lane; i < n_rows; i += 32) {
          const uint32_t row = uint32_t(lds_s32(runs + uint32_t(i) * 4u));
          bulk_g2s(sbase + uint32_t(i) * slab_bytes, xb + uint64_t(row) * pitch, row_bytes, bar);
        }
      } else {
        const int64_t col = col0 + int64_t(lchunk) * VN;
        const char* xb = xbase + (int64_t(b) * a.x_bstride + col) * int64_t(sizeof(T));
        const uint32_t dst = sbase + uint32_t(lchunk) * 16u;
        const int step = rpp * npw;
        if (col < a.feat) {
          constexpr int UN = 4;
          int i0 = pw * rpp + lrow;
          for (; i0 + (UN - 1) * step < n_rows; i0 += UN * step) {   // full groups: no per-copy predicate
            uint32_t rows[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) rows[u] = uint32_t(lds_s32(runs + uint32_t(i0 + u * step) * 4u));
#pragma unroll
            for (int u = 0; u < UN; ++u)
              cp_async16(dst + uint32_t(i0 + u * step) * slab_bytes, xb + uint64_t(rows[u]) * pitch);
          }
          for (; i0 < n_rows; i0 += step)
            cp_async16(dst + uint32_t(i0) * slab_bytes, xb + uint64_t(uint32_t(lds_s32(runs + uint32_t(i0) * 4u))) * pitch);
        }
        cp_async_mbar_arrive(smem_u32(&full_bar[st]));
      }
      if (++st == ns) { st = 0; ++st_round; }
      if (++slab == a.slabs) { slab = 0; ++b; }
      if (++rem == per_tile) { rem = 0; b = 0; ++seq; }
    }
    return;
  }
  if (warp == ncw) {
    // ===== producer =====
    if (lane == 0) {
      int st = 0;
      uint32_t st_round = 0;  // how many times the stage ring wrapped
      int seq = 0, rem = 0, b = 0, slab = 0;
      for (int64_t it = 0; it < n_items; ++it) {
        const int t = a.tile_begin + int(blockIdx.x) + seq * int(gridDim.x);
        if (rem == 0) {  // first item of a tile: its records and messages
          const int mb_i = seq & 1;
          if (seq >= 2) mbar_wait(smem_u32(&mempty_bar[mb_i]), uint32_t((seq >> 1) - 1) & 1u);
          const int p0 = __ldg(a.tile_ptr + t), p1 = __ldg(a.tile_ptr + t + 1);
          const int m0 = __ldg(a.tmsg_base + t) & ~1, m1 = __ldg(a.tmsg_base + t + 1);
          const uint32_t rb = uint32_t(p1 - p0) * 16u, mb = uint32_t((m1 - m0 + 1) & ~1) * 8u;
          const uint32_t bar = smem_u32(&mfull_bar[mb_i]);
          const uint32_t dst = meta0 + uint32_t(mb_i) * a.meta_bytes;
          mbar_expect_tx(bar, rb + mb);
          bulk_g2s(dst, a.trec + p0, rb, bar);
          if (mb) bulk_g2s(dst + a.rec_bytes, a.tmsg + m0, mb, bar);
        }
        if (st_round > 0) mbar_wait(smem_u32(&empty_bar[st]), (st_round - 1) & 1u);
        const int r0 = __ldg(a.run_ptr + t), r1 = __ldg(a.run_ptr + t + 1);
        const uint32_t bar = smem_u32(&full_bar[st]);
        const uint32_t dst = stage0 + uint32_t(st) * a.stage_bytes;
        mbar_expect_tx(bar, uint32_t(r1 - r0) * run_bytes);
        for (int r = r0; r < r1; ++r)
          tma_load_3d(dst + uint32_t(r - r0) * run_bytes, &xmap, slab * a.slab_elems,
                      __ldg(a.run_start + r), b, bar);
        if (++st == ns) { st = 0; ++st_round; }
        if (++slab == a.slabs) { slab = 0; ++b; }
        if (++rem == per_tile) { rem = 0; b = 0; ++seq; }
      }
    }
    return;
  }

  // ===== consumers =====
  const int sub = lane / LPR, l = lane % LPR;
  int st = 0;
  uint32_t st_round = 0;
  int seq = 0, rem = 0, b = 0, slab = 0, n_rows = 0;   // item = (tile seq, batch b, slab), kept incrementally
  for (int64_t it = 0; it < n_items; ++it) {
    const int t = a.tile_begin + int(blockIdx.x) + seq * int(gridDim.x);
    const int64_t f0 = int64_t(slab) * a.slab_elems;
    const int64_t col = f0 + int64_t(l) * VN;
    const bool on = col < a.feat;
    const uint32_t stage = stage0 + uint32_t(st) * a.stage_bytes + uint32_t(l) * 16u;
    const uint32_t recs = meta0 + uint32_t(seq & 1) * a.meta_bytes;
    const uint32_t msgs = recs + a.rec_bytes;
    if (rem == 0) mbar_wait(smem_u32(&mfull_bar[seq & 1]), uint32_t(seq >> 1) & 1u);
    mbar_wait(smem_u32(&full_bar[st]), st_round & 1u);
    if (rem == 0) n_rows = __ldg(a.tile_ptr + t + 1) - __ldg(a.tile_ptr + t);
    T* ob = static_cast<T*>(a.out) + b * a.o_bstride + col;
    float bv[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) bv[k] = (a.bias && on) ? __ldg(a.bias + col + k) : 0.0f;
    for (int p = warp * RPW + sub; p < n_rows; p += ncw * RPW) {
      const int4 rec = lds_i4(recs + uint32_t(p) * 16u);  // {dst, message offset, degree, -}
      float acc[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = 0.0f;
      // (measured: a separate unpredicated path for full batches of U messages is slower, 139 -> 145 us at cfg 2)
      for (int e0 = 0; e0 < rec.z; e0 += U)
        reduce_batch<T, U>(acc, msgs + uint32_t(rec.y + e0) * 8u, stage, slab_bytes, min(U, rec.z - e0));
      if (on)
        stg_v4(ob + int64_t(rec.x) * a.ldo, finish_reg<T, VN>(acc, bv, a.bias != nullptr, a.relu));
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&empty_bar[st]));
      if (rem == per_tile - 1) mbar_arrive(smem_u32(&mempty_bar[seq & 1]));
    }
    if (++st == ns) { st = 0; ++st_round; }
    if (++slab == a.slabs) { slab = 0; ++b; }
    if (++rem == per_tile) { rem = 0; b = 0; ++seq; }
  }
}

// Developer tuning knobs (read once): GWEN_TILED_THREADS (multiple of 32, 64..512),
// GWEN_TILED_STAGES (2..8), GWEN_TILED_U (5 or 9 messages in flight per row).
inline int env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v) return dflt;
  int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
inline int tiled_threads_hint() {
  static int v = env_int("GWEN_TILED_THREADS", 512, 64, 512) / 32 * 32;
  return v;
}
// gather mode of locality plans: GWEN_GATHER_MODE 1 = cp.async per 16 bytes, 2 = one bulk copy per row slab;
// GWEN_GATHER_WARPS = producer warps of mode 1
inline int gather_mode_hint() {
  static int v = env_int("GWEN_GATHER_MODE", 1, 1, 2);
  return v;
}
inline int gather_warps_hint() {
  static int v = env_int("GWEN_GATHER_WARPS", 2, 1, 4);
  return v;
}
inline int num_stages_hint() {
  static int v = env_int("GWEN_TILED_STAGES", 4, 2, 8);
  return v;
}

template <typename T, int LPR, int U>
int launch_tiled_u(const CUtensorMap& map, const TiledArgs& a, size_t smem_bytes, int threads,
                   cudaStream_t st) {
  auto kern = k_agg_tiled<T, LPR, U>;
  GWEN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem_bytes)));
  const int grid = std::min(a.num_tiles, std::max(1, sm_count() - sm_reserve()));
  kern<<<grid, threads, smem_bytes, st>>>(map, a);
  GWEN_LAUNCH_CHECK("k_agg_tiled");
  return GWEN_OK;
}

template <typename T, int LPR>
int launch_tiled(const CUtensorMap& map, const TiledArgs& a, size_t smem_bytes, int threads,
                 cudaStream_t st) {
  static const int u = env_int("GWEN_TILED_U", 9, 1, 9);
  return u >= 9 ? launch_tiled_u<T, LPR, 9>(map, a, smem_bytes, threads, st)
                : launch_tiled_u<T, LPR, 5>(map, a, smem_bytes, threads, st);
}

// ---- tile plan --------------------------------------------------------------------------------
constexpr int kThreads = 256;

struct PlanWs {
  size_t off_keys_in, off_keys_out, off_vals_in, off_vals_out, off_rank, off_usrc, off_uptr,
      off_ulocal, off_rtmp, off_nruns, off_pos, off_cub, cub_bytes, bytes;
  int bits;
};

cudaError_t plan_ws(int64_t n_dst, int64_t m, int64_t num_tiles, PlanWs* p) {
  size_t t = static_cast<size_t>(m > 0 ? m : 1);
  size_t nt = static_cast<size_t>(num_tiles + 1);
  size_t nd = static_cast<size_t>(n_dst + 1);
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
  p->off_usrc = take(t * 4);    // distinct (tile, source) list
  p->off_uptr = take(nt * 4);   // its per-tile offsets
  p->off_ulocal = take(t * 4);  // stage row of each distinct source
  p->off_rtmp = take(t * 4);    // run starts before compaction
  p->off_nruns = take(nt * 4);
  p->off_pos = take(nd * 4);    // message offset of every position of the processing order
  int tb = 1;
  while ((int64_t(1) << tb) < num_tiles + 1) ++tb;
  p->bits = 32 + tb;
  size_t sort_bytes = 0, scan_bytes = 0, scan2_bytes = 0;
  cudaError_t err = cub::DeviceRadixSort::SortPairs(
      nullptr, sort_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
      static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(t),
      0, p->bits);
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(t));
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr),
                                      static_cast<int>(std::max(nt, nd)));
  if (err != cudaSuccess) return err;
  p->cub_bytes = std::max(sort_bytes, std::max(scan_bytes, scan2_bytes));
  p->off_cub = take(p->cub_bytes);
  p->bytes = off;
  return cudaSuccess;
}

// degree of every position of the processing order (then exclusive-scanned into pos)
__global__ void k_plan_deg(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ order,
                           int64_t n_dst, int32_t* __restrict__ pos) {
  int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p > n_dst) return;
  if (p == n_dst) {
    pos[p] = 0;
    return;
  }
  const int64_t d = order ? order[p] : p;
  pos[p] = rowptr[d + 1] - rowptr[d];
}

// thread per position p: row record, sort keys (tile << 32 | src) and the weight half of the
// processing-order message list.
__global__ void k_plan_keys(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                            const float* __restrict__ w, const int32_t* __restrict__ order,
                            const int32_t* __restrict__ tile_ptr,
                            const int32_t* __restrict__ pos, int num_tiles, int64_t n_dst,
                            uint64_t* __restrict__ keys, int32_t* __restrict__ vals,
                            int4* __restrict__ trec, uint64_t* __restrict__ tmsg,
                            int32_t* __restrict__ tmsg_base, int32_t* __restrict__ status) {
  int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p <= num_tiles) {
    const int tp = tile_ptr[p];
    tmsg_base[p] = pos[tp];
    if (p < num_tiles) {
      const int tp1 = tile_ptr[p + 1];
      atomicMax(&status[3], pos[tp1] - (pos[tp] & ~1));  // messages staged per tile
      atomicMax(&status[4], tp1 - tp);                    // rows per tile
    }
  }
  if (p >= n_dst) return;
  int lo = 0, hi = num_tiles;  // last tile with tile_ptr[t] <= p
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile_ptr[mid] <= p) lo = mid; else hi = mid;
  }
  const int64_t d = order ? order[p] : p;
  const int beg = rowptr[d], end = rowptr[d + 1];
  const int q0 = pos[p];
  trec[p] = make_int4(static_cast<int>(d), q0 - (pos[tile_ptr[lo]] & ~1), end - beg, 0);
  for (int s = beg; s < end; ++s) {
    const int q = q0 + (s - beg);
    keys[q] = (uint64_t(uint32_t(lo)) << 32) | uint32_t(src[s]);
    vals[q] = q;
    tmsg[q] = uint64_t(__float_as_uint(w[s])) << 32;
  }
}

__global__ void k_plan_heads(const uint64_t* __restrict__ keys, int64_t m,
                             int32_t* __restrict__ rank) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  rank[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// distinct (tile, source) pairs: usrc[u] and per-tile offsets uptr[t]
__global__ void k_plan_unique(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rank,
                              int64_t m, int num_tiles, int32_t* __restrict__ uptr,
                              int32_t* __restrict__ usrc) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  const uint64_t k = keys[i];
  const int t = static_cast<int>(k >> 32);
  const int u = rank[i] - 1;
  const bool head = i == 0 || k != keys[i - 1];
  if (head) usrc[u] = static_cast<int32_t>(uint32_t(k));
  const int tp = i ? static_cast<int>(keys[i - 1] >> 32) : -1;
  for (int tt = tp + 1; tt <= t; ++tt) uptr[tt] = u;
  if (i == m - 1)
    for (int tt = t + 1; tt <= num_tiles; ++tt) uptr[tt] = u + 1;
}

// thread per tile: cover the tile's ascending distinct sources greedily with windows of run_len
// consecutive row ids; a source's stage row = window index * run_len + offset in the window.
__global__ void k_plan_runs(const int32_t* __restrict__ uptr, const int32_t* __restrict__ usrc,
                            int num_tiles, int run_len, int32_t* __restrict__ ulocal,
                            int32_t* __restrict__ rtmp, int32_t* __restrict__ nruns,
                            int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > num_tiles) return;
  if (t == num_tiles) {
    nruns[t] = 0;
    return;
  }
  const int u0 = uptr[t], u1 = uptr[t + 1];
  int runs = 0;
  int64_t cur = INT64_MIN / 2;
  for (int u = u0; u < u1; ++u) {
    const int s = usrc[u];
    if (s >= cur + run_len) {
      cur = s;
      rtmp[u0 + runs] = s;
      ++runs;
    }
    ulocal[u] = (runs - 1) * run_len + static_cast<int>(s - cur);
  }
  nruns[t] = runs;
  atomicMax(&status[1], runs);
  atomicAdd(&status[2], u1 - u0);
}

__global__ void k_plan_compact(const int32_t* __restrict__ uptr, const int32_t* __restrict__ rtmp,
                               const int32_t* __restrict__ run_ptr, int num_tiles,
                               int32_t* __restrict__ run_start, int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  const int r0 = run_ptr[t], r1 = run_ptr[t + 1], u0 = uptr[t];
  for (int r = r0; r < r1; ++r) run_start[r] = rtmp[u0 + (r - r0)];
  if (t == num_tiles - 1) status[0] = r1;
}

__global__ void k_plan_msg(const int32_t* __restrict__ vals, const int32_t* __restrict__ rank,
                           const int32_t* __restrict__ ulocal, int64_t m,
                           uint64_t* __restrict__ tmsg) {
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  tmsg[vals[i]] |= uint64_t(uint32_t(ulocal[rank[i] - 1]));
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
                                    int64_t num_tiles, int64_t n_dst, int64_t m, int32_t run_len,
                                    int32_t* run_ptr, int32_t* run_start, void* trec,
                                    uint64_t* tmsg, int32_t* tmsg_base, int32_t* status, void* ws,
                                    size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GWEN_CHECK_ARG(rowptr && src && w && tile_ptr && run_ptr && run_start && trec && tmsg &&
                     tmsg_base && status && ws,
                 "null pointer");
  GWEN_CHECK_ARG(num_tiles > 0 && num_tiles < INT32_MAX && n_dst > 0 && m > 0 && m < INT32_MAX,
                 "bad sizes");
  GWEN_CHECK_ARG(run_len >= 1 && run_len <= 256, "run_len must be in [1, 256]");
  GWEN_CHECK_ARG(aligned16(trec) && aligned16(tmsg), "trec / tmsg must be 16-byte aligned");
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
  int32_t* usrc = reinterpret_cast<int32_t*>(base + p.off_usrc);
  int32_t* uptr = reinterpret_cast<int32_t*>(base + p.off_uptr);
  int32_t* ulocal = reinterpret_cast<int32_t*>(base + p.off_ulocal);
  int32_t* rtmp = reinterpret_cast<int32_t*>(base + p.off_rtmp);
  int32_t* nruns = reinterpret_cast<int32_t*>(base + p.off_nruns);
  int32_t* pos = reinterpret_cast<int32_t*>(base + p.off_pos);
  void* cub_ws = base + p.off_cub;
  const int nt = static_cast<int>(num_tiles);
  GWEN_CUDA(cudaMemsetAsync(status, 0, 8 * sizeof(int32_t), st));
  GWEN_CUDA(cudaMemsetAsync(tmsg + m, 0, 2 * sizeof(uint64_t), st));  // tail padding
  const unsigned grid_n =
      static_cast<unsigned>(ceil_div(std::max<int64_t>(n_dst, num_tiles) + 1, kThreads));
  const unsigned grid_m = static_cast<unsigned>(ceil_div(m, kThreads));
  const unsigned grid_t = static_cast<unsigned>(ceil_div(num_tiles + 1, kThreads));
  k_plan_deg<<<grid_n, kThreads, 0, st>>>(rowptr, order, n_dst, pos);
  GWEN_LAUNCH_CHECK("k_plan_deg");
  size_t cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, pos, pos,
                                          static_cast<int>(n_dst + 1), st));
  k_plan_keys<<<grid_n, kThreads, 0, st>>>(rowptr, src, w, order, tile_ptr, pos, nt, n_dst,
                                           keys_in, vals_in, static_cast<int4*>(trec), tmsg,
                                           tmsg_base, status);
  GWEN_LAUNCH_CHECK("k_plan_keys");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, vals_in,
                                            vals_out, static_cast<int>(m), 0, p.bits, st));
  k_plan_heads<<<grid_m, kThreads, 0, st>>>(keys_out, m, rank);
  GWEN_LAUNCH_CHECK("k_plan_heads");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes, rank, rank, static_cast<int>(m), st));
  k_plan_unique<<<grid_m, kThreads, 0, st>>>(keys_out, rank, m, nt, uptr, usrc);
  GWEN_LAUNCH_CHECK("k_plan_unique");
  k_plan_runs<<<grid_t, kThreads, 0, st>>>(uptr, usrc, nt, run_len, ulocal, rtmp, nruns, status);
  GWEN_LAUNCH_CHECK("k_plan_runs");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, nruns, run_ptr, nt + 1, st));
  k_plan_compact<<<grid_t, kThreads, 0, st>>>(uptr, rtmp, run_ptr, nt, run_start, status);
  GWEN_LAUNCH_CHECK("k_plan_compact");
  k_plan_msg<<<grid_m, kThreads, 0, st>>>(vals_out, rank, ulocal, m, tmsg);
  GWEN_LAUNCH_CHECK("k_plan_msg");
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

extern "C" int gwen_aggregate_tiled_fwd(const gwen_tile_plan* plan, const void* x, void* out,
                                        int64_t batch, int64_t n_src, int64_t feat, int64_t ldx,
                                        int64_t x_bstride, int64_t ldo, int64_t o_bstride,
                                        int dtype, const float* bias, int epilogue,
                                        int32_t slab_elems, int32_t tile_begin,
                                        int32_t tile_count, void* stream) {
  GWEN_CHECK_ARG(plan && x && out, "null pointer");
  GWEN_CHECK_ARG(tile_begin >= 0 && tile_count >= 0 && tile_begin + tile_count <= plan->num_tiles,
                 "tile range outside the plan");
  if (tile_count == 0) {
    if (tile_begin != 0) return GWEN_OK;  // empty range
    tile_count = plan->num_tiles;
  }
  GWEN_CHECK_ARG(batch >= 0 && feat >= 0 && n_src >= 0, "negative size");
  if (batch == 0 || feat == 0 || plan->n_dst == 0) return GWEN_OK;
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  GWEN_CHECK_ARG(plan->tile_ptr && plan->run_ptr && plan->run_start && plan->trec && plan->tmsg &&
                     plan->tmsg_base && plan->run_len > 0 && plan->max_tile_runs > 0 &&
                     plan->max_tile_rows > 0,
                 "incomplete plan");
  const int esz = dtype == GWEN_F32 ? 4 : 2;
  const int vn = 16 / esz;
  if (feat % vn || ldx % vn || ldo % vn || x_bstride % vn || o_bstride % vn || !aligned16(x) ||
      !aligned16(out))
    return set_err(GWEN_E_ALIGN, "tiled aggregate needs 16-byte aligned rows (feat %% %d == 0)", vn);
  // shared memory: NS >= 2 data stages (max_tile_runs * run_len rows of one slab each) + two
  // metadata buffers (row records + messages of a tile).  slab = 8, 16 or 32 sixteen-byte
  // chunks; slab_elems is an upper bound that is halved until two stages fit.
  const int64_t rows = int64_t(plan->max_tile_runs) * plan->run_len;
  const size_t rec_bytes = align_up(size_t(plan->max_tile_rows) * 16, 128);
  // gather mode: single-row runs copied by cp.async; the tile's run list travels with the metadata
  // (run_start then needs 4 readable entries past run_ptr[num_tiles]: GWEN_PLAN_GATHER promises them)
  const bool gather = plan->run_len == 1 && (plan->reserved & GWEN_PLAN_GATHER);
  GWEN_CHECK_ARG(!gather || ldx * esz < (int64_t(1) << 32), "row pitch must be below 4 GB");
  const size_t runs_off = rec_bytes + align_up(size_t(plan->max_tile_msgs + 2) * 8, 128);
  const size_t meta_bytes = runs_off + (gather ? align_up(size_t(plan->max_tile_runs + 8) * 4, 128) : 0);
  const size_t smem_cap = 226 * 1024;
  auto smem_for = [&](int lpr_, int ns_) {
    return size_t(ns_) * size_t(rows) * lpr_ * 16 + 2 * meta_bytes + 256;
  };
  int lpr = slab_elems > 0 ? static_cast<int>(slab_elems / vn) : 32;
  if (lpr != 8 && lpr != 16 && lpr != 32)
    return set_err(GWEN_E_BADARG, "slab_elems must be %d, %d or %d", 8 * vn, 16 * vn, 32 * vn);
  while (lpr > 8 && (lpr * vn / 2 >= feat || smem_for(lpr, 2) > smem_cap)) lpr /= 2;
  if (smem_for(lpr, 2) > smem_cap)
    return set_err(GWEN_E_NOSUPPORT,
                   "tile (%lld staged rows x %d B, %zu B metadata) does not fit twice in smem",
                   (long long)rows, lpr * 16, meta_bytes);
  int ns = 2;
  while (ns < 8 && ns < num_stages_hint() && smem_for(lpr, ns + 1) <= smem_cap) ++ns;
  const size_t stage_bytes = static_cast<size_t>(rows) * lpr * 16;
  const size_t smem = smem_for(lpr, ns);
  const int slab = lpr * vn;
  CUtensorMap map;
  int rc = make_tensor_map_3d(&map, x, dtype, feat, n_src, batch, ldx, x_bstride, slab,
                              plan->run_len, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != GWEN_OK) return rc;
  TiledArgs a{plan->tile_ptr, plan->run_ptr, plan->run_start,
              static_cast<const int4*>(plan->trec), plan->tmsg, plan->tmsg_base, out, bias, batch,
              feat, ldo, o_bstride, tile_count, tile_begin, static_cast<int>(ceil_div(feat, slab)), slab,
              plan->run_len, (epilogue & GWEN_EPI_RELU) ? 1 : 0,
              static_cast<uint32_t>(stage_bytes), static_cast<uint32_t>(rec_bytes),
              static_cast<uint32_t>(meta_bytes), ns, gather ? gather_mode_hint() : 0, gather_warps_hint(), x, ldx,
              x_bstride,
              static_cast<uint32_t>(runs_off)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = tiled_threads_hint();
  if (dtype == GWEN_F32) {
    if (lpr == 32) return launch_tiled<float, 32>(map, a, smem, threads, st);
    if (lpr == 16) return launch_tiled<float, 16>(map, a, smem, threads, st);
    return launch_tiled<float, 8>(map, a, smem, threads, st);
  }
  if (lpr == 32) return launch_tiled<__nv_bfloat16, 32>(map, a, smem, threads, st);
  if (lpr == 16) return launch_tiled<__nv_bfloat16, 16>(map, a, smem, threads, st);
  return launch_tiled<__nv_bfloat16, 8>(map, a, smem, threads, st);
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
