// K2 backward (weight gradient) for fp32 on the tensor cores:  dW[N_out, K_in] = dY^T X with
// fp32-level accuracy as three tcgen05 kind::tf32 products per K step (see linear_tf32x3.cu for the
// hi / lo splitting).  Both operands are reduced over their SLOW index (the row).  The bf16 kernel
// (linear_wgrad_tc.cu) reads them as MN-major operands; for 32-bit elements the transform warps that
// split the tiles anyway also TRANSPOSE them in shared memory, so the MMAs see plain K-major tiles.
//
// CTA (tile, split): one 128 (N_out) x 128 (K_in) tile of dW over the rows [split * per, +per):
//   warp 0 lane 0   TMA producer: per 32-row block 4 dY boxes + 4 X boxes of {32 fp32, 32 rows}
//                   (SWIZZLE_128B) into a 2-stage RAW ring
//   warps 2..5      transform: thread t owns tile column t of both operands: reads its 32 rows from the
//                   raw boxes (conflict-free: a warp reads 32 consecutive floats per row), splits each
//                   value into hi / lo (nearest TF32, rounded exact remainder) and writes them as ONE
//                   128-byte K-major SWIZZLE_128B row of the hi and lo tiles of a 2-stage operand ring;
//                   fence.proxy.async, arrive on xf_full[stage], release the raw stage
//   warp 1 lane 0   per K = 8 step three MMAs lo_a hi_b + hi_a lo_b + hi_a hi_b into one of two TMEM
//                   accumulators, alternating every kChunkF blocks (the accumulator rounds toward zero on
//                   every add: short chunks keep that bias ~1e-6)
//   warps 6..9      epilogue: add each finished chunk to the split's fp32 partial tile (round to nearest)
// Partials are summed in split order by k_reduce_splits (linear.cu): deterministic.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;
namespace {

constexpr int RKF = 32;            // rows (reduction) per stage
constexpr int BNF = 128;           // K_in columns per tile
constexpr int kWgfThreads = 10 * 32;
constexpr int kChunkF = 16;        // 512 rows = 192 accumulator adds per TMEM accumulation
constexpr uint32_t kBoxF = RKF * 128u;              // one {32 fp32, 32 rows} box = 4 KB
constexpr uint32_t kOperF = 4u * kBoxF;             // 128 columns x 32 rows of one operand = 16 KB
constexpr uint32_t kRawStage = 2u * kOperF;         // raw dY | raw X = 32 KB
constexpr uint32_t kXfStage = 4u * kOperF;          // A_hi | A_lo | B_hi | B_lo (K-major) = 64 KB
constexpr int kRawStages = 2, kXfStages = 2;

struct WgfArgs {
  float* part;           // [splits][n_out][k_in]
  int64_t m, per;
  int n_out, k_in, tiles_j;
};

__device__ __forceinline__ void split_tf32_rn(uint32_t u, uint32_t& h, uint32_t& l) {
  h = (u + 0x1000u) & 0xFFFFE000u;
  const uint32_t d = __float_as_uint(__uint_as_float(u) - __uint_as_float(h));
  l = (d + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kWgfThreads, 1)
    k_wgrad_tf32x3(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                   WgfArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t raw_full[kRawStages], raw_empty[kRawStages], xf_full[kXfStages],
      xf_empty[kXfStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // operand ring, then the raw ring
  const uint32_t raw_base = base + kXfStages * kXfStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int i0 = (tile / g.tiles_j) * BM, j0 = (tile % g.tiles_j) * BNF;
  const int64_t r_begin = int64_t(split) * g.per;
  const int64_t r_end = r_begin + g.per < g.m ? r_begin + g.per : g.m;
  const int n_blocks = r_end > r_begin ? int((r_end - r_begin + RKF - 1) / RKF) : 0;
  const int n_chunks = (n_blocks + kChunkF - 1) / kChunkF;
  const uint32_t tmem_cols = 2u * BNF;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&dymap);
    tma_prefetch_desc(&xmap);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(smem_u32(&raw_full[i]), 1);
      mbar_init(smem_u32(&raw_empty[i]), 4);        // the 4 transform warps
    }
    for (int i = 0; i < kXfStages; ++i) {
      mbar_init(smem_u32(&xf_full[i]), 4);
      mbar_init(smem_u32(&xf_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: raw dY boxes at +0, raw X boxes at +kOperF =====
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int s = blk % kRawStages;
        const uint32_t round = uint32_t(blk / kRawStages);
        if (round > 0) mbar_wait(smem_u32(&raw_empty[s]), (round - 1) & 1u);
        const uint32_t bar = smem_u32(&raw_full[s]);
        const uint32_t dst = raw_base + uint32_t(s) * kRawStage;
        const int r0 = int(r_begin) + blk * RKF;   // rows beyond m are zero-filled by TMA
        mbar_expect_tx(bar, kRawStage);
        for (int c = 0; c < 4; ++c) {
          tma_load_3d(dst + uint32_t(c) * kBoxF, &dymap, i0 + 32 * c, r0, 0, bar);
          tma_load_3d(dst + kOperF + uint32_t(c) * kBoxF, &xmap, j0 + 32 * c, r0, 0, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // kind::tf32, D fp32, A = B = tf32 (format 2), both K-major, M = 128, N = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BNF >> 3) << 17) |
                             (uint32_t(BM >> 4) << 24);
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int chunk = blk / kChunkF, cb = blk - chunk * kChunkF;
        const uint32_t acc = uint32_t(chunk) & 1u;
        if (cb == 0 && chunk >= 2) {
          mbar_wait(smem_u32(&tmem_empty_bar[acc]), uint32_t((chunk >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const int s = blk % kXfStages;
        mbar_wait(smem_u32(&xf_full[s]), uint32_t(blk / kXfStages) & 1u);
        tc_fence_after();
        const uint32_t st0 = base + uint32_t(s) * kXfStage;
        const uint64_t a_hi = make_smem_desc(st0), a_lo = make_smem_desc(st0 + kOperF);
        const uint64_t b_hi = make_smem_desc(st0 + 2 * kOperF), b_lo = make_smem_desc(st0 + 3 * kOperF);
        const uint32_t d_addr = tmem_d + acc * uint32_t(BNF);
#pragma unroll
        for (int kk = 0; kk < RKF / 8; ++kk) {  // K = 8 per MMA: +32 bytes inside the 128-byte K-major row
          const uint64_t o = uint64_t(kk * 2);
          umma_tf32(d_addr, a_lo + o, b_hi + o, idesc, (cb | kk) ? 1u : 0u);
          umma_tf32(d_addr, a_hi + o, b_lo + o, idesc, 1u);
          umma_tf32(d_addr, a_hi + o, b_hi + o, idesc, 1u);
        }
        umma_commit(smem_u32(&xf_empty[s]));
        if (cb == kChunkF - 1 || blk == n_blocks - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else if (warp < 6) {
    // ===== transform warps 2..5: thread t = tile column t (dW row for dY, dW column for X) =====
    const int t = int(threadIdx.x) - 64;                 // 0 .. 127
    const uint32_t box = uint32_t(t >> 5), cin = uint32_t(t & 31);   // raw box and column inside it
    const uint32_t row_out = uint32_t(t) * 128u, sw = uint32_t(t) & 7u;
    for (int blk = 0; blk < n_blocks; ++blk) {
      const int rs = blk % kRawStages, xs = blk % kXfStages;
      mbar_wait(smem_u32(&raw_full[rs]), uint32_t(blk / kRawStages) & 1u);
      if (blk >= kXfStages) mbar_wait(smem_u32(&xf_empty[xs]), uint32_t(blk / kXfStages - 1) & 1u);
      const uint32_t raw0 = raw_base + uint32_t(rs) * kRawStage, xf0 = base + uint32_t(xs) * kXfStage;
#pragma unroll
      for (int op = 0; op < 2; ++op) {
        // raw box: row r at byte 128 r, 16-byte chunk c stored at chunk c ^ (r % 8)
        const uint32_t rb = raw0 + uint32_t(op) * kOperF + box * kBoxF + (cin & 3u) * 4u;
        const uint32_t hi_row = xf0 + uint32_t(op) * 2u * kOperF + row_out, lo_row = hi_row + kOperF;
#pragma unroll
        for (int c4 = 0; c4 < RKF / 4; ++c4) {            // 4 reduction rows -> one 16-byte K-major chunk
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t r = uint32_t(c4 * 4 + e);
            uint32_t u;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(rb + r * 128u + (((cin >> 2) ^ (r & 7u)) << 4)));
            split_tf32_rn(u, h[e], l[e]);
          }
          const uint32_t dst = ((uint32_t(c4) ^ sw) << 4);   // K-major SWIZZLE_128B: chunk ^= row % 8
          sts_v4(hi_row + dst, make_uint4(h[0], h[1], h[2], h[3]));
          sts_v4(lo_row + dst, make_uint4(l[0], l[1], l[2], l[3]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&raw_empty[rs]));
        mbar_arrive(smem_u32(&xf_full[xs]));
      }
    }
  } else {
    // ===== epilogue warps 6..9: TMEM lanes 32 (warp % 4) .. +31 = tile rows =====
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    float* prow = g.part + (int64_t(split) * g.n_out + i) * g.k_in + j0;
    const int cols = g.k_in - j0 < BNF ? g.k_in - j0 : BNF;   // K_in tail (multiple of 4)
    if (n_chunks == 0 && i < g.n_out)
      for (int c = 0; c < cols; c += 4) *reinterpret_cast<uint4*>(prow + c) = make_uint4(0u, 0u, 0u, 0u);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const uint32_t acc = uint32_t(chunk) & 1u;
      mbar_wait(smem_u32(&tmem_full_bar[acc]), uint32_t(chunk >> 1) & 1u);
      tc_fence_after();
      for (int c = 0; c < BNF; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_d + acc * uint32_t(BNF) + (uint32_t(q * 32) << 16) + uint32_t(c), r);
        if (i < g.n_out) {
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            if (c + t < cols) {
              uint4 v = make_uint4(r[t], r[t + 1], r[t + 2], r[t + 3]);
              if (chunk > 0) {
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

}  // namespace

int linear_wgrad_tf32x3_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldx,
                                  const void* dy, const void* x) {
  static const bool disabled = getenv("GWEN_FP32_SIMT") != nullptr || getenv("GWEN_DISABLE_TC") != nullptr;
  if (disabled || m < 4096 || m > INT32_MAX) return 0;
  if (k_in % 4 || n_out % 4 || k_in < 32 || n_out < 32 || lddy % 4 || ldx % 4) return 0;
  return aligned16(dy) && aligned16(x);
}

int linear_wgrad_tf32x3_splits(int64_t m, int64_t k_in, int64_t n_out) {
  const int64_t tiles = ceil_div(n_out, BM) * ceil_div(k_in, BNF);
  int64_t s = std::max<int64_t>(1, sm_count() / tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, m / (8 * RKF)));
  return static_cast<int>(std::min<int64_t>(s, 65535));
}

int linear_wgrad_tf32x3(const void* dy, const void* x, float* part, int64_t m, int64_t k_in, int64_t n_out,
                        int64_t lddy, int64_t ldx, cudaStream_t st) {
  CUtensorMap dymap, xmap;
  int rc = make_tensor_map_3d(&dymap, dy, GWEN_F32, n_out, m, 1, lddy, 0, 32, RKF, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&xmap, x, GWEN_F32, k_in, m, 1, ldx, 0, 32, RKF, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  const int splits = linear_wgrad_tf32x3_splits(m, k_in, n_out);
  const int64_t per = ceil_div(ceil_div(m, splits), RKF) * RKF;
  const size_t smem = kXfStages * size_t(kXfStage) + kRawStages * size_t(kRawStage) + 1024;
  const int tiles_j = static_cast<int>(ceil_div(k_in, BNF));
  WgfArgs g{part, m, per, static_cast<int>(n_out), static_cast<int>(k_in), tiles_j};
  GWEN_CUDA(cudaFuncSetAttribute(k_wgrad_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>(ceil_div(n_out, BM) * tiles_j), static_cast<unsigned>(splits));
  k_wgrad_tf32x3<<<grid, kWgfThreads, smem, st>>>(dymap, xmap, g);
  GWEN_LAUNCH_CHECK("k_wgrad_tf32x3");
  return GWEN_OK;
}

}  // namespace gwen
