// K2 for fp32 on the tensor cores: y = epi(x W^T + bias) with fp32 operands and fp32-level accuracy,
// as three tcgen05 kind::tf32 products per K step (the "3xTF32" error-compensated scheme).
//
// The reference runs its fp32 Linear on SIMT cuBLAS (allow_tf32 = False); a single TF32 product
// (10 mantissa bits) would miss the 1e-5 parity bar.  Split every operand once into
//   hi = x rounded to the nearest TF32 value (10 mantissa bits)
//   lo = (x - hi), an exact difference, rounded to TF32 again (the next 11 bits)
// and accumulate  lo_a * hi_b + hi_a * lo_b + hi_a * hi_b  in the fp32 TMEM accumulator: the dropped
// terms are O(2^-21) relative, i.e. fp32-level.  W is split once per call by k_split_tf32 (it is
// small); x is split IN the kernel: the raw fp32 tile lands by TMA, four transform warps rewrite it
// in place as hi and write lo next to it (the split is elementwise, so the swizzled layout does not
// matter), fence.proxy.async, and arrive on the leader's xf[stage] barrier.  Otherwise this is the
// CTA-pair pipeline of linear_tc3.cu with four operand tiles per stage (128 B = 32 fp32 per swizzle
// row), K = 8 per MMA, and an fp32 TMA-store epilogue.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;
namespace {

constexpr int kEpiWarpsF = 16;
constexpr int kXfWarps = 4;      // transform warps (split the A tile into hi / lo in shared memory)
constexpr int kTfThreads = 64 + 32 * kXfWarps + 32 * kEpiWarpsF;
constexpr int BKF = 32;          // fp32 elements per K block = one 128-byte swizzle row
constexpr int UMMA_KF = 8;       // kind::tf32
constexpr int kMaxStagesF = 6;

struct TfArgs {
  const float* bias;
  int64_t m;
  int n, k_blocks, bn, stages, relu;
};

// x = hi + lo + O(2^-24 |x|): both parts rounded to nearest TF32 (10 mantissa bits; adding half an ulp
// before clearing the low 13 bits rounds the magnitude), so the split error is unbiased.
__device__ __forceinline__ void split_tf32(uint32_t u, uint32_t& h, uint32_t& l) {
  h = (u + 0x1000u) & 0xFFFFE000u;
  const uint32_t d = __float_as_uint(__uint_as_float(u) - __uint_as_float(h));   // exact
  l = (d + 0x1000u) & 0xFFFFE000u;
}

__global__ void k_split_tf32(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                             int64_t n4) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4;
       i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(u[j], h[j], l[j]);
    reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// hi / lo split of the TRANSPOSE of w [rows, cols] (for the dgrad: dx = dy W = dy (W^T)^T); w is small
__global__ void k_split_tf32_t(const float* __restrict__ w, float* __restrict__ hi_t,
                               float* __restrict__ lo_t, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? w[int64_t(r) * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;          // output row c, column r
    if (c < cols && r < rows) {
      uint32_t h, l;
      split_tf32(__float_as_uint(tile[threadIdx.x][i]), h, l);
      hi_t[int64_t(c) * rows + r] = __uint_as_float(h);
      lo_t[int64_t(c) * rows + r] = __uint_as_float(l);
    }
  }
}

__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: D = fp32, A = B = tf32 (format 2), both K-major, M = 256 (pair), N = bn
__host__ __device__ constexpr uint32_t make_idesc_tf32_pair(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTfThreads, 1)
    k_linear_tf32x3(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bhi,
                    const __grid_constant__ CUtensorMap blo, const __grid_constant__ CUtensorMap ymap,
                    TfArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // full_a: this CTA's raw x tile has landed (local); xf: both CTAs' tiles are split (leader's is used);
  // full_b: both CTAs' W hi / lo halves have landed (leader's is used)
  __shared__ __align__(8) uint64_t full_a[kMaxStagesF], xf_bar[kMaxStagesF], full_b[kMaxStagesF],
      empty_bar[kMaxStagesF], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t a_bytes = BM * 128u, b_bytes = uint32_t(g.bn / 2) * 128u;
  const uint32_t stage_bytes = 2u * (a_bytes + b_bytes);   // A_hi | A_lo | B_hi | B_lo
  const uint32_t staging = base + uint32_t(g.stages) * stage_bytes;  // 16 warps x 2 KB
  float* bias_s = reinterpret_cast<float*>(smem_raw + (staging - smem_u32(smem_raw)) + size_t(kEpiWarpsF) * 2048u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = g.n / g.bn;
  const int n_sub = g.bn / 16;   // 16-column (64-byte) epilogue sub-chunks per tile
  const int64_t total = ((g.m + 2 * BM - 1) / (2 * BM)) * n_tiles;
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // two accumulator sets (tile double buffering) x two partial sums (first / second half of K)
  const uint32_t tmem_cols = uint32_t(4 * g.bn);
  const int kb_half = (g.k_blocks + 1) / 2;
  const uint32_t epi_arrivals = 2u * 4u * uint32_t(n_sub < 4 ? n_sub : 4);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bhi);
    tma_prefetch_desc(&blo);
    tma_prefetch_desc(&ymap);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(smem_u32(&full_a[i]), 1);
      mbar_init(smem_u32(&xf_bar[i]), 2 * kXfWarps);
      mbar_init(smem_u32(&full_b[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), epi_arrivals);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.n; i += kTfThreads) bias_s[i] = g.bias ? __ldg(g.bias + i) : 0.0f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      uint32_t it = 0;
      for (int64_t tile = cluster_id; tile < total; tile += n_clusters) {
        const int m0 = int(tile / n_tiles) * (2 * BM) + int(rank) * BM;
        const int n0 = int(tile % n_tiles) * g.bn + int(rank) * (g.bn / 2);
        for (int kb = 0; kb < g.k_blocks; ++kb, ++it) {
          const uint32_t s = it % uint32_t(g.stages), round = it / uint32_t(g.stages);
          if (round > 0) mbar_wait(smem_u32(&empty_bar[s]), (round - 1) & 1u);
          const uint32_t dst = base + s * stage_bytes;
          mbar_expect_tx(smem_u32(&full_a[s]), a_bytes);            // raw x tile: local barrier
          tma_load_3d(dst, &amap, kb * BKF, m0, 0, smem_u32(&full_a[s]));
          if (leader) mbar_expect_tx(smem_u32(&full_b[s]), 4 * b_bytes);
          const uint32_t bar = mapa_u32(smem_u32(&full_b[s]), 0);
          tma_load_3d_pair(dst + 2 * a_bytes, &bhi, kb * BKF, n0, 0, bar);
          tma_load_3d_pair(dst + 2 * a_bytes + b_bytes, &blo, kb * BKF, n0, 0, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {  // ===== MMA issuer: 3 products per K step, small terms first =====
      const uint32_t idesc = make_idesc_tf32_pair(g.bn);
      uint32_t it = 0, seq = 0;
      for (int64_t tile = cluster_id; tile < total; tile += n_clusters, ++seq) {
        const uint32_t acc = seq & 1u, use = seq >> 1;
        if (use > 0) mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use - 1) & 1u);
        tc_fence_after();
        // The tensor core's fp32 accumulator rounds toward zero on every add, a bias that grows with the
        // length of the sum: accumulate the two halves of K separately and add them (round to nearest)
        // in the epilogue -- measured 7.7e-6 -> ~4e-6 normalised error at K = 1024.
        for (int kb = 0; kb < g.k_blocks; ++kb, ++it) {
          const uint32_t part = kb >= kb_half ? 1u : 0u;
          const uint32_t d_addr = tmem_d + (acc * 2u + part) * uint32_t(g.bn);
          const uint32_t first = (kb == 0 || kb == kb_half) ? 1u : 0u;
          const uint32_t s = it % uint32_t(g.stages), par = (it / uint32_t(g.stages)) & 1u;
          mbar_wait(smem_u32(&full_b[s]), par);
          mbar_wait(smem_u32(&xf_bar[s]), par);
          tc_fence_after();
          const uint32_t st0 = base + s * stage_bytes;
          const uint64_t d_ahi = make_smem_desc(st0), d_alo = make_smem_desc(st0 + a_bytes);
          const uint64_t d_bhi = make_smem_desc(st0 + 2 * a_bytes);
          const uint64_t d_blo = make_smem_desc(st0 + 2 * a_bytes + b_bytes);
#pragma unroll
          for (int kk = 0; kk < BKF / UMMA_KF; ++kk) {  // +32 bytes (>>4 = 2) per K = 8 inside the row
            const uint64_t o = uint64_t(kk * 2);
            umma_tf32_pair(d_addr, d_alo + o, d_bhi + o, idesc, (first && kk == 0) ? 0u : 1u);
            umma_tf32_pair(d_addr, d_ahi + o, d_blo + o, idesc, 1u);
            umma_tf32_pair(d_addr, d_ahi + o, d_bhi + o, idesc, 1u);
          }
          umma_commit_pair(smem_u32(&empty_bar[s]));
        }
        umma_commit_pair(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else if (warp < 2 + kXfWarps) {
    // ===== transform warps: raw fp32 tile -> hi (in place) + lo (next slot), 16 bytes per step =====
    const int t = int(threadIdx.x) - 64;                      // 0 .. 127
    const uint32_t xf_remote_base = mapa_u32(smem_u32(&xf_bar[0]), 0);
    uint32_t it = 0;
    for (int64_t tile = cluster_id; tile < total; tile += n_clusters) {
      for (int kb = 0; kb < g.k_blocks; ++kb, ++it) {
        const uint32_t s = it % uint32_t(g.stages), par = (it / uint32_t(g.stages)) & 1u;
        mbar_wait(smem_u32(&full_a[s]), par);
        const uint32_t a_hi = base + s * stage_bytes + uint32_t(t) * 16u, a_lo = a_hi + a_bytes;
#pragma unroll
        for (int j = 0; j < int(BM * 128 / 16 / (32 * kXfWarps)); ++j) {   // 1024 uint4 / 128 threads
          const uint32_t off = uint32_t(j) * (32u * kXfWarps * 16u);
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a_hi + off));
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(u[e], h[e], l[e]);
          sts_v4(a_hi + off, make_uint4(h[0], h[1], h[2], h[3]));
          sts_v4(a_lo + off, make_uint4(l[0], l[1], l[2], l[3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(xf_remote_base + s * 8u);
      }
    }
  } else {
    // ===== epilogue warps 6..21: TMEM lanes 32*(warp%4) .. +31, 16-column sub-chunks g4, g4+4, .. =====
    const int q = warp & 3, g4 = (warp - 2 - kXfWarps) >> 2;
    const uint32_t my_stage = staging + uint32_t(warp - 2 - kXfWarps) * 2048u;
    const uint32_t row_off = uint32_t(lane) * 64u;
    const uint32_t sw = uint32_t(lane >> 1) & 3u;  // SWIZZLE_64B
    const uint32_t empty_remote0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_remote1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    uint32_t seq = 0;
    if (g4 < n_sub) {
      for (int64_t tile = cluster_id; tile < total; tile += n_clusters, ++seq) {
        const int m0 = int(tile / n_tiles) * (2 * BM) + int(rank) * BM;
        const int n0 = int(tile % n_tiles) * g.bn;
        const uint32_t acc = seq & 1u;
        mbar_wait(smem_u32(&tmem_full_bar[acc]), (seq >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_addr = tmem_d + acc * 2u * uint32_t(g.bn) + (uint32_t(q * 32) << 16);
        const bool two_parts = g.k_blocks > 1;
        for (int sc = g4; sc < n_sub; sc += 4) {
          const int c = sc * 16;
          uint32_t r[16], r2[16];
          tmem_ld16_nowait(t_addr + uint32_t(c), r);
          if (two_parts) tmem_ld16_nowait(t_addr + uint32_t(g.bn) + uint32_t(c), r2);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (sc + 4 >= n_sub) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? empty_remote1 : empty_remote0);
          }
          __syncwarp();
          if (two_parts) {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
          }
          const uint32_t sbuf = my_stage + row_off;
          const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 4 chunks of 4 fp32 = 16 bytes
            const float4 b = bp[j];
            float f0 = __uint_as_float(r[4 * j + 0]) + b.x, f1 = __uint_as_float(r[4 * j + 1]) + b.y;
            float f2 = __uint_as_float(r[4 * j + 2]) + b.z, f3 = __uint_as_float(r[4 * j + 3]) + b.w;
            if (g.relu) {
              f0 = fmaxf(f0, 0.0f); f1 = fmaxf(f1, 0.0f); f2 = fmaxf(f2, 0.0f); f3 = fmaxf(f3, 0.0f);
            }
            sts_v4(sbuf + ((uint32_t(j) ^ sw) << 4),
                   make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3)));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&ymap, my_stage, n0 + c, m0 + q * 32, 0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(tmem_cols)
                 : "memory");
  }
}

}  // namespace

int linear_tf32x3_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                            const void* x, const void* w, const void* y) {
  static const bool disabled = getenv("GWEN_FP32_SIMT") != nullptr || getenv("GWEN_DISABLE_TC") != nullptr;
  // k <= 2048: the accumulator's rounding error grows with the length of the sum (4.4e-6 at k = 1024)
  if (disabled || m < 4096 || m > INT32_MAX || k < 32 || k > 2048 || n_out < 64 || n_out > 8192) return 0;
  if (k % 4 || n_out % 64 || ldx != k || ldw != k || ldy % 4 || sm_count() % 2) return 0;
  return aligned16(x) && aligned16(w) && aligned16(y);
}

size_t linear_tf32x3_workspace_bytes(int64_t m, int64_t k, int64_t n_out) {
  (void)m;
  return 2 * align_up(size_t(n_out) * k * 4, 256);   // W hi / lo (x is split inside the kernel)
}

static int tf32x3_launch(const void* x, const float* w_hi, const float* w_lo, void* y, int64_t m, int64_t k,
                         int64_t n_out, int64_t ldy, const float* bias, int relu, cudaStream_t st) {
  const int bn = n_out % 128 == 0 ? 128 : 64;
  CUtensorMap amap, bhi, blo, ymap;
  int rc = make_tensor_map_3d(&amap, x, GWEN_F32, k, m, 1, k, 0, BKF, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&bhi, w_hi, GWEN_F32, k, n_out, 1, k, 0, BKF, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&blo, w_lo, GWEN_F32, k, n_out, 1, k, 0, BKF, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&ymap, y, GWEN_F32, n_out, m, 1, ldy, 0, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != GWEN_OK) return rc;
  const int k_blocks = static_cast<int>(ceil_div(k, BKF));
  const size_t stage_bytes = 2 * size_t(BM + bn / 2) * 128;
  const size_t staging_bytes = size_t(kEpiWarpsF) * 2048 + align_up(size_t(n_out) * 4, 1024);
  const int stages = static_cast<int>(std::min<size_t>(kMaxStagesF, (226 * 1024 - staging_bytes - 1024) / stage_bytes));
  if (stages < 2) return set_err(GWEN_E_NOSUPPORT, "tf32x3 tile does not fit in shared memory");
  const size_t smem = std::max<size_t>(stages * stage_bytes + staging_bytes + 1024, 120 * 1024);
  TfArgs g{bias, m, static_cast<int>(n_out), k_blocks, bn, stages, relu};
  GWEN_CUDA(cudaFuncSetAttribute(k_linear_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t total = ceil_div(m, 2 * BM) * (n_out / bn);
  const int pairs = static_cast<int>(std::min<int64_t>(total, std::max(1, (sm_count() - sm_reserve()) / 2)));
  k_linear_tf32x3<<<2 * pairs, kTfThreads, smem, st>>>(amap, bhi, blo, ymap, g);
  GWEN_LAUNCH_CHECK("k_linear_tf32x3");
  return GWEN_OK;
}

int linear_tf32x3_fwd(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                      int64_t ldy, const float* bias, int relu, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  if (ws_bytes < linear_tf32x3_workspace_bytes(m, k, n_out))
    return set_err(GWEN_E_WORKSPACE, "tf32x3 workspace too small");
  const size_t wa = align_up(size_t(n_out) * k * 4, 256);
  float* w_hi = static_cast<float*>(ws);
  float* w_lo = reinterpret_cast<float*>(static_cast<char*>(ws) + wa);
  const int64_t w4 = n_out * k / 4;
  k_split_tf32<<<static_cast<unsigned>(std::min<int64_t>(ceil_div(w4, 256), int64_t(sm_count()) * 16)), 256, 0, st>>>(
      static_cast<const float*>(w), w_hi, w_lo, w4);
  GWEN_LAUNCH_CHECK("k_split_tf32");
  return tf32x3_launch(x, w_hi, w_lo, y, m, k, n_out, ldy, bias, relu, st);
}

// dx[M, K_in] = dy[M, N_out] W[N_out, K_in]: the forward kernel on dy with the split TRANSPOSE of W
int linear_tf32x3_dgrad(const void* dy, const void* w, void* dx, int64_t m, int64_t k_in, int64_t n_out,
                        int64_t lddx, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < linear_tf32x3_workspace_bytes(m, n_out, k_in))
    return set_err(GWEN_E_WORKSPACE, "tf32x3 workspace too small");
  const size_t wa = align_up(size_t(n_out) * k_in * 4, 256);
  float* wt_hi = static_cast<float*>(ws);
  float* wt_lo = reinterpret_cast<float*>(static_cast<char*>(ws) + wa);
  dim3 grid(static_cast<unsigned>(ceil_div(k_in, 32)), static_cast<unsigned>(ceil_div(n_out, 32)));
  k_split_tf32_t<<<grid, dim3(32, 8), 0, st>>>(static_cast<const float*>(w), wt_hi, wt_lo,
                                                static_cast<int>(n_out), static_cast<int>(k_in));
  GWEN_LAUNCH_CHECK("k_split_tf32_t");
  return tf32x3_launch(dy, wt_hi, wt_lo, dx, m, n_out, k_in, lddx, nullptr, 0, st);
}

}  // namespace gwen
