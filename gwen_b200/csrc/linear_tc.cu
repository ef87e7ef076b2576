// K2 (tensor-core path): y = epi(x W^T + bias) for bf16 on the 5th-generation tensor cores.
//
//   x : [M, K] bf16 row-major (K-major A operand)      W : [N, K] bf16 row-major (K-major B)
//   y : [M, N] bf16, fp32 accumulation in TMEM
//
// One CTA computes a 128 x BN output tile (BN = 32..256).  Warp roles (192 threads):
//   warp 0 (one lane)  TMA producer: per 64-wide K block one cp.async.bulk.tensor for the A tile
//                      (128 x 64) and one for the B tile (BN x 64), SWIZZLE_128B, landing on
//                      full[stage]; waits empty[stage] before reusing a stage.
//   warp 1             allocates BN TMEM columns; one lane issues 4 x tcgen05.mma (M=128, N=BN,
//                      K=16, kind::f16) per K block from shared-memory descriptors and commits to
//                      empty[stage]; the last commit also arrives on tmem_full.
//   warps 2..5         epilogue: tcgen05.ld their 32 TMEM lanes (= 32 output rows) 32 columns at
//                      a time, add bias, ReLU, round to bf16, 16-byte global stores.
// 1-2 CTAs are resident per SM (shared memory / TMEM columns permitting) so one CTA's epilogue
// overlaps another's main loop.  Rows beyond M and columns of K beyond the tensor are zero-filled
// by TMA; stores are masked.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;

// linear_tc3.cu (CTA-pair kernel)
int linear_tc3_supported(int64_t m, int64_t k, int64_t n_out, int b_mn);
int linear_tc3_fwd_bf16(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                        int64_t ldx, int64_t ldw, int64_t ldy, const float* bias, int relu,
                        int b_mn, cudaStream_t st, int64_t batch = 1, int64_t x_bstride = 0,
                        int64_t y_bstride = 0, const void* mask = nullptr, int64_t ldm = 0, int64_t m_bstride = 0);

namespace {

struct TcArgs {
  __nv_bfloat16* y;
  const float* bias;
  int64_t m, ldy;
  int n, k_blocks, bn, stages, relu;
};

__global__ void __launch_bounds__(kTcThreads) k_linear_tc(const __grid_constant__ CUtensorMap amap,
                                                          const __grid_constant__ CUtensorMap bmap,
                                                          TcArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B: 1024-B aligned
  const uint32_t a_bytes = BM * BK * 2, b_bytes = uint32_t(g.bn) * BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = g.n / g.bn;
  const int64_t m0 = (int64_t(blockIdx.x) / n_tiles) * BM;
  const int n0 = int(blockIdx.x % n_tiles) * g.bn;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_fence_init();
  }
  if (warp == 1) {  // TMEM allocation (power of two >= 32 columns), owned by warp 1
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(uint32_t(g.bn < 32 ? 32 : g.bn))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        const int s = kb % g.stages;
        const uint32_t round = uint32_t(kb / g.stages);
        if (round > 0) mbar_wait(smem_u32(&empty_bar[s]), (round - 1) & 1u);
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t dst = tiles + uint32_t(s) * stage_bytes;
        mbar_expect_tx(bar, stage_bytes);
        tma_load_3d(dst, &amap, kb * BK, int(m0), 0, bar);
        tma_load_3d(dst + a_bytes, &bmap, kb * BK, n0, 0, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(g.bn);
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        const int s = kb % g.stages;
        mbar_wait(smem_u32(&full_bar[s]), uint32_t(kb / g.stages) & 1u);
        tc_fence_after();
        const uint32_t a_addr = tiles + uint32_t(s) * stage_bytes;
        const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(a_addr + a_bytes);
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk)  // +32 bytes (>>4 = 2) per UMMA_K inside the row
          umma_f16(tmem_d, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc,
                   (kb | kk) ? 1u : 0u);
        umma_commit(smem_u32(&empty_bar[s]));  // frees the stage when these MMAs retire
      }
      umma_commit(smem_u32(&tmem_full_bar));   // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int64_t row = m0 + q * 32 + lane;
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    __nv_bfloat16* yrow = g.y + row * g.ldy + n0;
    for (int c = 0; c < g.bn; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(c), r);
      if (row < g.m) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            f[t] = __uint_as_float(r[j + t]);
            if (g.bias) f[t] += __ldg(g.bias + n0 + c + j + t);
            if (g.relu) f[t] = fmaxf(f[t], 0.0f);
          }
          uint4 o;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(f[6], f[7]);
          o.x = *reinterpret_cast<uint32_t*>(&p0);
          o.y = *reinterpret_cast<uint32_t*>(&p1);
          o.z = *reinterpret_cast<uint32_t*>(&p2);
          o.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(yrow + c + j) = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(uint32_t(g.bn < 32 ? 32 : g.bn))
                 : "memory");
  }
}

// ---- v2: persistent, double-buffered accumulator, TMA-store epilogue ---------------------------
//   grid = #SMs; tiles (128 x BN) are taken round-robin, N tiles fastest so the CTAs that run
//   together share their A tile through L2.
//   warp 0 / warp 1: as above, but looping over tiles with a continuous stage ring; the MMA
//     issuer alternates between two TMEM accumulators (2 x BN columns) and waits tmem_empty[acc]
//     before overwriting one.
//   warps 2..5: per tile, per 64-column chunk: tcgen05.ld -> bias/ReLU -> bf16 -> st.shared into a
//     128B-swizzled 32 x 64 staging tile (conflict-free) -> one TMA tensor store per warp
//     (cp.async.bulk.tensor, bulk-group completion; rows beyond M are clipped by the tensor map).
//     After the last TMEM read of a tile the warp arrives on tmem_empty[acc], so the next tile's
//     MMAs run under this tile's stores.
struct Tc2Args {
  const float* bias;
  int64_t m;
  int n, k_blocks, bn, stages, relu;
};

__global__ void __launch_bounds__(kTcThreads, 1)
    k_linear_tc2(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                 const __grid_constant__ CUtensorMap ymap, Tc2Args g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[2][256];  // the tile's bias slice, double-buffered
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * BK * 2, b_bytes = uint32_t(g.bn) * BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t staging = base + uint32_t(g.stages) * stage_bytes;  // 4 warps x 2 bufs x 4 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = g.n / g.bn;
  const int64_t total = ((g.m + BM - 1) / BM) * n_tiles;
  const uint32_t tmem_cols = uint32_t(2 * g.bn);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
    tma_prefetch_desc(&ymap);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);  // one arrival per epilogue warp
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
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = int(tile / n_tiles) * BM, n0 = int(tile % n_tiles) * g.bn;
        for (int kb = 0; kb < g.k_blocks; ++kb, ++it) {
          const uint32_t s = it % uint32_t(g.stages), round = it / uint32_t(g.stages);
          if (round > 0) mbar_wait(smem_u32(&empty_bar[s]), (round - 1) & 1u);
          const uint32_t bar = smem_u32(&full_bar[s]);
          const uint32_t dst = base + s * stage_bytes;
          mbar_expect_tx(bar, stage_bytes);
          tma_load_3d(dst, &amap, kb * BK, m0, 0, bar);
          tma_load_3d(dst + a_bytes, &bmap, kb * BK, n0, 0, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(g.bn);
      uint32_t it = 0, seq = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x, ++seq) {
        const uint32_t acc = seq & 1u, use = seq >> 1;
        if (use > 0) mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use - 1) & 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_d + acc * uint32_t(g.bn);
        for (int kb = 0; kb < g.k_blocks; ++kb, ++it) {
          const uint32_t s = it % uint32_t(g.stages);
          mbar_wait(smem_u32(&full_bar[s]), (it / uint32_t(g.stages)) & 1u);
          tc_fence_after();
          const uint32_t a_addr = base + s * stage_bytes;
          const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(a_addr + a_bytes);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            umma_f16(d_addr, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc,
                     (kb | kk) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[s]));
        }
        umma_commit(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const uint32_t my_stage = staging + uint32_t(warp - 2) * 8192u;  // two 4 KB buffers
    const uint32_t row_off = uint32_t(lane) * 128u;
    const uint32_t sw = uint32_t(lane & 7);
    uint32_t seq = 0, buf = 0;
    for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x, ++seq) {
      const int m0 = int(tile / n_tiles) * BM, n0 = int(tile % n_tiles) * g.bn;
      const uint32_t acc = seq & 1u;
      // stage this tile's bias slice once (128 epilogue threads, <= 256 values) instead of
      // re-reading it from global memory per row; buffer `acc` was last read two tiles ago and
      // every epilogue warp has passed the named barrier of the tile in between.
      {
        const int et = int(threadIdx.x) - 64;
        for (int i = et; i < g.bn; i += 128) bias_s[acc][i] = g.bias ? __ldg(g.bias + n0 + i) : 0.0f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(smem_u32(&tmem_full_bar[acc]), (seq >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_d + acc * uint32_t(g.bn) + (uint32_t(q * 32) << 16);
      for (int c = 0; c < g.bn; c += 64) {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_addr + uint32_t(c), r0);
        tmem_ld32(t_addr + uint32_t(c + 32), r1);
        if (c + 64 >= g.bn) {  // last TMEM read of this tile: release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
        }
        // the staging buffer we are about to overwrite must have been read by its TMA store
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t sbuf = my_stage + buf * 4096u + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 chunks of 8 columns = 16 bytes
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            f[t] = __uint_as_float(j < 4 ? r0[j * 8 + t] : r1[(j - 4) * 8 + t]) +
                   bias_s[acc][c + j * 8 + t];
            if (g.relu) f[t] = fmaxf(f[t], 0.0f);
          }
          uint4 o;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(f[6], f[7]);
          o.x = *reinterpret_cast<uint32_t*>(&p0);
          o.y = *reinterpret_cast<uint32_t*>(&p1);
          o.z = *reinterpret_cast<uint32_t*>(&p2);
          o.w = *reinterpret_cast<uint32_t*>(&p3);
          sts_v4(sbuf + ((uint32_t(j) ^ sw) << 4), o);  // SWIZZLE_128B: chunk ^= row % 8
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&ymap, my_stage + buf * 4096u, n0 + c, m0 + q * 32, 0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1u;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

int pick_bn(int64_t n) {
  for (int bn : {256, 128, 64, 32})
    if (n % bn == 0) return bn;
  return 0;
}

}  // namespace

int linear_tc_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                        const void* x, const void* w, const void* y) {
  static const bool disabled = getenv("GWEN_DISABLE_TC") != nullptr;
  if (disabled) return 0;
  if (m < 1 || k < 8 || n_out < 32 || n_out > INT32_MAX || m > INT32_MAX) return 0;
  if (k % 8 || ldx % 8 || ldw % 8 || ldy % 8) return 0;  // 16-byte pitches for TMA / v4 stores
  if (!aligned16(x) || !aligned16(w) || !aligned16(y)) return 0;
  return pick_bn(n_out) != 0;
}

int linear_tc2_fwd_bf16(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                        int64_t ldx, int64_t ldw, int64_t ldy, const float* bias, int relu,
                        cudaStream_t st) {
  int bn = 0;
  for (int c : {256, 128, 64})
    if (n_out % c == 0) { bn = c; break; }
  if (!bn) return set_err(GWEN_E_NOSUPPORT, "tcgen05 GEMM v2 needs n_out %% 64 == 0");
  CUtensorMap amap, bmap, ymap;
  int rc = make_tensor_map_3d(&amap, x, GWEN_BF16, k, m, 1, ldx, 0, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&bmap, w, GWEN_BF16, k, n_out, 1, ldw, 0, BK, bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&ymap, y, GWEN_BF16, n_out, m, 1, ldy, 0, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  const int k_blocks = static_cast<int>(ceil_div(k, BK));
  const size_t stage_bytes = size_t(BM + bn) * BK * 2;
  const size_t staging_bytes = 4 * 2 * 4096;
  int stages = static_cast<int>(std::min<size_t>(8, (224 * 1024 - staging_bytes - 1024) / stage_bytes));
  if (stages < 2) return set_err(GWEN_E_NOSUPPORT, "tile does not fit in shared memory");
  // >= 120 KB keeps one CTA per SM (each CTA owns up to all 512 TMEM columns)
  const size_t smem = std::max<size_t>(stages * stage_bytes + staging_bytes + 1024, 120 * 1024);
  Tc2Args g{bias, m, static_cast<int>(n_out), k_blocks, bn, stages, relu};
  GWEN_CUDA(cudaFuncSetAttribute(k_linear_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  const int64_t total = ceil_div(m, BM) * (n_out / bn);
  const int grid = static_cast<int>(std::min<int64_t>(total, sm_count()));
  k_linear_tc2<<<grid, kTcThreads, smem, st>>>(amap, bmap, ymap, g);
  GWEN_LAUNCH_CHECK("k_linear_tc2");
  return GWEN_OK;
}

int linear_tc_fwd_bf16(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                       int64_t ldx, int64_t ldw, int64_t ldy, const float* bias, int relu,
                       cudaStream_t st) {
  static const bool v1_only = getenv("GWEN_TC_V1") != nullptr;
  if (!v1_only && linear_tc3_supported(m, k, n_out, 0))
    return linear_tc3_fwd_bf16(x, w, y, m, k, n_out, ldx, ldw, ldy, bias, relu, 0, st);
  if (!v1_only && n_out % 64 == 0)
    return linear_tc2_fwd_bf16(x, w, y, m, k, n_out, ldx, ldw, ldy, bias, relu, st);
  const int bn = pick_bn(n_out);
  if (!bn) return set_err(GWEN_E_NOSUPPORT, "tcgen05 GEMM needs n_out %% 32 == 0");
  CUtensorMap amap, bmap;
  int rc = make_tensor_map_3d(&amap, x, GWEN_BF16, k, m, 1, ldx, 0, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&bmap, w, GWEN_BF16, k, n_out, 1, ldw, 0, BK, bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  const int k_blocks = static_cast<int>(ceil_div(k, BK));
  const size_t stage_bytes = size_t(BM + bn) * BK * 2;
  // up to 4 stages, at most ~100 KB so that two CTAs share an SM
  int stages = static_cast<int>(std::min<int64_t>(std::min<int64_t>(4, k_blocks), (100 * 1024) / stage_bytes));
  if (stages < 1) stages = 1;
  if (stages < 2 && k_blocks > 1) stages = 2;
  const size_t smem = stages * stage_bytes + 1024;
  TcArgs g{static_cast<__nv_bfloat16*>(y), bias, m, ldy, static_cast<int>(n_out), k_blocks, bn,
           stages, relu};
  GWEN_CUDA(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  const int64_t grid = ceil_div(m, BM) * (n_out / bn);
  if (grid > INT32_MAX) return set_err(GWEN_E_NOSUPPORT, "GEMM grid too large");
  k_linear_tc<<<static_cast<unsigned>(grid), kTcThreads, smem, st>>>(amap, bmap, g);
  GWEN_LAUNCH_CHECK("k_linear_tc");
  return GWEN_OK;
}


// dx[M, K_in] = dy[M, N_out] W[N_out, K_in]: the forward kernel with the reduction over N_out and
// W read as an MN-major B operand (no transposed copy of the weights).
int linear_tc_dgrad_supported(int64_t m, int64_t k_in, int64_t n_out, int64_t lddy, int64_t ldw,
                              int64_t lddx, const void* dy, const void* w, const void* dx) {
  if (!linear_tc_supported(m, n_out, k_in, lddy, ldw, lddx, dy, w, dx)) return 0;
  return linear_tc3_supported(m, n_out, k_in, 1);
}
int linear_tc_dgrad_bf16(const void* dy, const void* w, void* dx, int64_t m, int64_t k_in,
                         int64_t n_out, int64_t lddy, int64_t ldw, int64_t lddx, cudaStream_t st,
                         const void* mask, int64_t ldm) {
  return linear_tc3_fwd_bf16(dy, w, dx, m, n_out, k_in, lddy, ldw, lddx, nullptr, 0, 1, st, 1, 0, 0, mask, ldm, 0);
}

// batch-strided forward / dgrad in ONE launch (bf16, CTA-pair kernel); 0 = not served
int linear_tc_batched_supported(int64_t m, int64_t k, int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy,
                                int64_t x_bstride, int64_t y_bstride, const void* x, const void* w,
                                const void* y, int b_mn) {
  if (x_bstride % 8 || y_bstride % 8) return 0;
  if (!linear_tc_supported(m, k, n_out, ldx, ldw, ldy, x, w, y)) return 0;
  return linear_tc3_supported(m, k, n_out, b_mn);
}
int linear_tc_batched_bf16(const void* x, const void* w, void* y, int64_t batch, int64_t m, int64_t k,
                           int64_t n_out, int64_t ldx, int64_t ldw, int64_t ldy, int64_t x_bstride,
                           int64_t y_bstride, const float* bias, int relu, int b_mn, cudaStream_t st,
                           const void* mask, int64_t ldm, int64_t m_bstride) {
  return linear_tc3_fwd_bf16(x, w, y, m, k, n_out, ldx, ldw, ldy, bias, relu, b_mn, st, batch, x_bstride,
                             y_bstride, mask, ldm, m_bstride);
}

}  // namespace gwen
