// K2 backward (weight gradient) on the tensor cores:  dW[N_out, K_in] = dY^T X, reduced over the
// M rows (nodes x members), bf16 operands, fp32 accumulation.
//
// Both operands are read exactly as they lie in memory ([M, N_out] and [M, K_in] row-major): the
// reduction index (the row) is the SLOW index of both, i.e. both are MN-major UMMA operands.  TMA
// boxes of {64 columns, 64 rows} with SWIZZLE_128B land as the canonical MN-major tile (row r at
// byte 128 r, 8-row groups 1024 B apart, 64-column chunks one box apart), so no transposed copy of
// the activations is ever made.
//
// CTA (tile, split): one 128 (N_out) x BN (K_in) tile of dW over the rows [split * per, +per):
//   warp 0 lane 0  TMA producer: per 64-row block two dY boxes + BN/64 X boxes into a 4-stage ring
//   warp 1 lane 0  4 x tcgen05.mma (M = 128, N = BN, K = 16, A and B MN-major) per block into one of
//                  two TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2..5     tcgen05.ld -> fp32 partial tile -> workspace[split] (16-byte stores)
// The tensor core's fp32 accumulator rounds toward zero on every add -- over the ~100 000 rows a split
// covers at COSMO-1E size that is a systematic shrink of 2-5e-5 (measured).  The row range is
// therefore cut into chunks of kChunkBlocks x 64 rows that alternate between the two accumulators; the
// epilogue adds every finished chunk to the split's partial tile in global memory (round to nearest,
// fixed order) while the next chunk is being accumulated.
// The partials are summed in split order by k_reduce_splits (linear.cu): deterministic.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;
namespace {

constexpr int RK = 64;  // rows (reduction) per stage
constexpr int kWgThreads = 192;
constexpr int kChunkBlocks = 64;   // 4096 rows per TMEM accumulation (shrink ~4e-6; 2048: 2e-6 but +15 % time)

struct WgArgs {
  float* dbpart;         // kDb: [splits][n_out] column sums of dY per split (the bias gradient's partials)
  float* part;           // [splits][n_out][k_in]
  int64_t m, per;        // rows, rows per split (multiple of RK)
  int n_out, k_in, bn, stages, tiles_j;
};

// kDb (round 2): the bias gradient db = column sums of dY out of the SAME pass -- the CTAs of the first K_in tile issue
// one more MMA per 16 rows, dY^T times a constant tile of ones (N = 16, a 64 x 64 box of bf16 1.0 in shared memory:
// all ones, so the 128-byte swizzle leaves it unchanged), into 16 extra TMEM columns per accumulator; the epilogue
// adds column 0 of every finished chunk to the split's db partial.  No pass over dY for the bias gradient is left
// when no ReLU mask has to be applied to it either (masked by the next layer's dgrad epilogue, or no ReLU).
// Needs 2 * bn + 48 <= 512 TMEM columns: bn <= 128, i.e. K_in not a multiple of 256 (the entry point declines those:
// forcing 128-column tiles there costs more than the pass it saves -- 512 -> 1024: 787 -> 1528 us, measured).
// 64 -> 1024 at M = 896 292: 299 -> 345 us, and the 267 us bias-sum pass is gone.
template <bool kDb>
__global__ void __launch_bounds__(kWgThreads, 1)
    k_wgrad_tc(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
               WgArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t box_bytes = RK * 128;                       // one {64 col, RK row} box
  const uint32_t a_bytes = 2 * box_bytes, b_bytes = uint32_t(g.bn / 64) * box_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int i0 = (tile / g.tiles_j) * BM, j0 = (tile % g.tiles_j) * g.bn;
  const int64_t r_begin = int64_t(split) * g.per;
  const int64_t r_end = r_begin + g.per < g.m ? r_begin + g.per : g.m;
  const int n_blocks = r_end > r_begin ? int((r_end - r_begin + RK - 1) / RK) : 0;
  const bool db_cta = kDb && (tile % g.tiles_j) == 0;      // this CTA also reduces dY's columns
  const uint32_t db_col0 = uint32_t(2 * g.bn);             // two 16-column db accumulators behind the dW ones
  const uint32_t ones_addr = base + uint32_t(g.stages) * stage_bytes;
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(2 * g.bn + (kDb ? 48 : 0))) tmem_cols *= 2;
  const int n_chunks = (n_blocks + kChunkBlocks - 1) / kChunkBlocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&dymap);
    tma_prefetch_desc(&xmap);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if constexpr (kDb) {   // the tile of ones: generic stores, read by the tensor core (async proxy)
    for (uint32_t o = threadIdx.x * 16u; o < box_bytes; o += kWgThreads * 16u)
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(ones_addr + o), "r"(0x3F803F80u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int s = blk % g.stages;
        const uint32_t round = uint32_t(blk / g.stages);
        if (round > 0) mbar_wait(smem_u32(&empty_bar[s]), (round - 1) & 1u);
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t dst = base + uint32_t(s) * stage_bytes;
        const int r0 = int(r_begin) + blk * RK;   // rows beyond m are zero-filled by TMA
        mbar_expect_tx(bar, stage_bytes);
        tma_load_3d(dst, &dymap, i0, r0, 0, bar);
        tma_load_3d(dst + box_bytes, &dymap, i0 + 64, r0, 0, bar);
        for (int c = 0; c < g.bn / 64; ++c)
          tma_load_3d(dst + a_bytes + uint32_t(c) * box_bytes, &xmap, j0 + 64 * c, r0, 0, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // kind::f16, D fp32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = bn
      const uint32_t idesc = make_idesc(g.bn) | (1u << 15) | (1u << 16);
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int chunk = blk / kChunkBlocks, cb = blk - chunk * kChunkBlocks;
        const uint32_t acc = uint32_t(chunk) & 1u;
        if (cb == 0 && chunk >= 2) {  // the epilogue must have drained this accumulator (chunk - 2)
          mbar_wait(smem_u32(&tmem_empty_bar[acc]), uint32_t((chunk >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const int s = blk % g.stages;
        mbar_wait(smem_u32(&full_bar[s]), uint32_t(blk / g.stages) & 1u);
        tc_fence_after();
        const uint32_t a_addr = base + uint32_t(s) * stage_bytes;
        const uint64_t adesc = make_smem_desc_mn(a_addr, box_bytes);
        const uint64_t bdesc = make_smem_desc_mn(a_addr + a_bytes, box_bytes);
        const uint32_t d_addr = tmem_d + acc * uint32_t(g.bn);
#pragma unroll
        for (int kk = 0; kk < RK / UMMA_K; ++kk)  // 16 rows of 128 B per UMMA_K
          umma_f16(d_addr, adesc + uint64_t(kk) * (2048 >> 4), bdesc + uint64_t(kk) * (2048 >> 4),
                   idesc, (cb | kk) ? 1u : 0u);
        if (db_cta) {   // dY^T x ones -> 16 identical columns of column sums
          const uint32_t idesc1 = make_idesc(16) | (1u << 15) | (1u << 16);
          const uint64_t odesc = make_smem_desc_mn(ones_addr, box_bytes);
#pragma unroll
          for (int kk = 0; kk < RK / UMMA_K; ++kk)
            umma_f16(tmem_d + db_col0 + acc * 16u, adesc + uint64_t(kk) * (2048 >> 4), odesc + uint64_t(kk) * (2048 >> 4),
                     idesc1, (cb | kk) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (cb == kChunkBlocks - 1 || blk == n_blocks - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lanes 32 (warp % 4) .. +31 = tile rows =====
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    float* prow = g.part + (int64_t(split) * g.n_out + i) * g.k_in + j0;
    float* dbp = db_cta ? g.dbpart + int64_t(split) * g.n_out + i : nullptr;
    if (n_chunks == 0) {  // empty row range: the partial tile is zero
      if (i < g.n_out) {
        for (int c = 0; c < g.bn; c += 4) *reinterpret_cast<uint4*>(prow + c) = make_uint4(0u, 0u, 0u, 0u);
        if (db_cta) *dbp = 0.0f;
      }
    }
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const uint32_t acc = uint32_t(chunk) & 1u;
      mbar_wait(smem_u32(&tmem_full_bar[acc]), uint32_t(chunk >> 1) & 1u);
      tc_fence_after();
      for (int c = 0; c < g.bn; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_d + acc * uint32_t(g.bn) + (uint32_t(q * 32) << 16) + uint32_t(c), r);
        if (i < g.n_out) {
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            uint4 v = make_uint4(r[t], r[t + 1], r[t + 2], r[t + 3]);
            if (chunk > 0) {  // running sum of the chunks, fixed order, round to nearest
              const uint4 p = *reinterpret_cast<const uint4*>(prow + c + t);
              v.x = __float_as_uint(__uint_as_float(p.x) + __uint_as_float(v.x));
              v.y = __float_as_uint(__uint_as_float(p.y) + __uint_as_float(v.y));
              v.z = __float_as_uint(__uint_as_float(p.z) + __uint_as_float(v.z));
              v.w = __float_as_uint(__uint_as_float(p.w) + __uint_as_float(v.w));
            }
            *reinterpret_cast<uint4*>(prow + c + t) = v;
          }
        }
      }
      if (db_cta) {   // column 0 of this chunk's 16 identical db columns (the 32-column load stays inside the allocation)
        uint32_t r[32];
        tmem_ld32(tmem_d + db_col0 + acc * 16u + (uint32_t(q * 32) << 16), r);
        if (i < g.n_out) *dbp = chunk > 0 ? *dbp + __uint_as_float(r[0]) : __uint_as_float(r[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(tmem_cols)
                 : "memory");
  }
}

int pick_bn_wg(int64_t k_in, bool with_db = false) {
  for (int bn : {256, 128, 64})
    if (k_in % bn == 0 && !(with_db && bn > 128)) return bn;
  return 0;
}

}  // namespace

int linear_tc_wgrad_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldx,
                              const void* dy, const void* x) {
  static const bool disabled = getenv("GWEN_DISABLE_TC") != nullptr;
  if (disabled || m < 256 || m > INT32_MAX) return 0;
  if (k_in % 64 || n_out % 64 || lddy % 8 || ldx % 8 || !aligned16(dy) || !aligned16(x)) return 0;
  return 1;
}

// number of row splits (partials) the tensor-core wgrad writes for this problem
int linear_tc_wgrad_splits(int64_t m, int64_t k_in, int64_t n_out, int with_db) {
  const int bn = pick_bn_wg(k_in, with_db != 0);
  const int64_t tiles = ceil_div(n_out, BM) * (k_in / bn);
  int64_t s = std::max<int64_t>(1, sm_count() / tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, m / (4 * RK)));
  return static_cast<int>(std::min<int64_t>(s, 65535));
}

int linear_tc_wgrad_bf16(const void* dy, const void* x, float* part, int64_t m, int64_t k_in,
                         int64_t n_out, int64_t lddy, int64_t ldx, cudaStream_t st, float* dbpart) {
  const int bn = pick_bn_wg(k_in, dbpart != nullptr);
  if (!bn) return set_err(GWEN_E_NOSUPPORT, "tcgen05 wgrad needs k_in %% 64 == 0");
  CUtensorMap dymap, xmap;
  int rc = make_tensor_map_3d(&dymap, dy, GWEN_BF16, n_out, m, 1, lddy, 0, 64, RK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&xmap, x, GWEN_BF16, k_in, m, 1, ldx, 0, 64, RK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  const int splits = linear_tc_wgrad_splits(m, k_in, n_out, dbpart != nullptr);
  const int64_t per = ceil_div(ceil_div(m, splits), RK) * RK;
  const size_t stage_bytes = size_t(2 + bn / 64) * RK * 128;
  const size_t ones_bytes = dbpart ? size_t(RK) * 128 : 0;
  const int stages = static_cast<int>(std::min<size_t>(8, (224 * 1024 - ones_bytes) / stage_bytes));
  const size_t smem = std::max<size_t>(stages * stage_bytes + ones_bytes + 1024, 120 * 1024);  // one CTA per SM
  const int tiles_j = static_cast<int>(k_in / bn);
  WgArgs g{dbpart, part, m, per, static_cast<int>(n_out), static_cast<int>(k_in), bn, stages, tiles_j};
  dim3 grid(static_cast<unsigned>(ceil_div(n_out, BM) * tiles_j), static_cast<unsigned>(splits));
  if (dbpart) {
    GWEN_CUDA(cudaFuncSetAttribute(k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    k_wgrad_tc<true><<<grid, kWgThreads, smem, st>>>(dymap, xmap, g);
    GWEN_LAUNCH_CHECK("k_wgrad_tc");
    return GWEN_OK;
  }
  GWEN_CUDA(cudaFuncSetAttribute(k_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  k_wgrad_tc<false><<<grid, kWgThreads, smem, st>>>(dymap, xmap, g);
  GWEN_LAUNCH_CHECK("k_wgrad_tc");
  return GWEN_OK;
}

}  // namespace gwen
