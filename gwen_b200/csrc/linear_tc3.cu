// K2 (tensor-core path, CTA pairs): y = epi(x W^T + bias) for bf16 with tcgen05.mma.cta_group::2.
//
// Why pairs: k_linear_tc2 (one CTA per 128 x 256 tile) streams 48 KB of operands per 64-deep K
// block into each SM for 512 cycles of MMA -- 94 B/clk/SM, more than an SM can pull from L2, and
// ncu shows the tensor pipe 59 % active with the issuer waiting on full barriers.  Here two CTAs
// of a cluster (one TPC) compute ONE 256 x BN tile: each CTA stages its own 128 rows of x and
// only HALF of the W tile (BN/2 rows), the leader CTA issues M = 256 MMAs that read both CTAs'
// shared memory, and each CTA's TMEM receives its own 128 accumulator rows.  Operand traffic per
// SM drops to 32 KB per K block (64 B/clk).
//
// Roles per CTA (576 threads):
//   warp 0 lane 0   TMA producer (both CTAs): A half + B half per K block, completing on the
//                   LEADER's full[stage] barrier (cp.async.bulk.tensor ... .cta_group::2).
//   warp 1 lane 0   MMA issuer (leader only): 4 x tcgen05.mma.cta_group::2 per K block;
//                   tcgen05.commit ... multicast::cluster frees the stage in BOTH CTAs and
//                   publishes the accumulator to both epilogues.
//   warps 2..17     epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31, the 32-column sub-chunks
//                   (w - 2) / 4, +4, ...: tcgen05.ld -> +bias (packed fp32x2, bias vector staged
//                   in shared memory once) -> bf16 -> ReLU (bf16x2 max) -> 64B-swizzled st.shared
//                   -> one TMA tensor store per warp and sub-chunk.  Two accumulators (2 x BN TMEM
//                   columns) so the next tile's MMAs run under this tile's epilogue; the epilogue
//                   warps of both CTAs arrive on the leader's tmem_empty barrier.
// Measured at M = 896 292 (ms, cuBLAS matmul without epilogue in brackets): 64->1024 0.38 (0.31),
// 1024->512 0.78 (0.72), 512->256 0.23 (0.23), 256->512 0.27 (0.28), 512->1024 0.81 (0.80),
// 1024->64 0.31 (0.32).  The ring depth matters (3/4/5 stages: 0.92/0.81/0.77 ms at 1024->512), an
// L2 prefetch of A ahead of the ring did not help and was removed; the 64->1024 case waits on
// cp.async.bulk.wait_group.read (output write-back), independent of tile order.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;
namespace {

constexpr int kEpiWarps = 16;
constexpr int kTc3Threads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;

struct Tc3Args {
  const float* bias;
  int64_t m;
  int n, k_blocks, bn, stages, relu, bufs, row_major_tiles;
  int epi_groups;  // epilogue warp groups (of 4 warps) that work: 4, or 2 to trade staging for ring depth
  int batch;  // independent [m, k] x-slices / [m, n] y-slices (third TMA coordinate), same W
  int b_mn;  // B operand is MN-major: w is [reduction][n] row-major (dgrad: dx = dy W)
  int epi_cols;  // output columns per epilogue chunk and TMA store: 32 (64-byte rows) or 64 (128-byte rows)
  uint32_t wait_ns;  // poll interval of the epilogue warps waiting for an accumulator (0 = spin)
};

__device__ __forceinline__ uint4 lds_v4_tc3(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
// 0xffff per bf16 half of `y2` that is > 0 (one HSET2.BF16; k_relu_bias_bwd's integer rule "sign clear and magnitude
// non-zero" gives the same mask for every value a ReLU can output)
__device__ __forceinline__ uint32_t gt0_mask(uint32_t y2) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&y2), __floats2bfloat162_rn(0.0f, 0.0f));
}

// kMask (round 2): the dgrad with the ReLU backward of the PREVIOUS layer folded in -- the output tile is multiplied
// by (mask > 0) elementwise, mask = a bf16 tensor of the output's shape (the layer input = the previous layer's ReLU
// output).  Per 32-column sub-chunk the mask tile is brought by ONE TMA load (same box and swizzle as the store) into
// the staging buffer the output will be packed into, one chunk ahead (two staging buffers per warp); each thread
// reads back the 64 bytes of its row, ANDs its packed bf16 pairs with the compare mask and overwrites them.
// Measured (512->1024 dgrad, M = 896 292, plain kernel 751 us): this form with an integer compare per half 1105 us;
// + the operand producer prefetching the next tile's mask block into L2 (UTMAPF) 1145 us (and, with the HSET2 compare,
// the epilogue warp prefetching its chunk-after-next tile: 980 -> 1000 us: the mask's latency is not the bound); per-thread 16-byte
// read-only loads of the mask row instead of the TMA tile 1164 us.
template <bool kMask>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc3Threads, 1)
    k_linear_tc3(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                 const __grid_constant__ CUtensorMap ymap, const __grid_constant__ CUtensorMap mmap, Tc3Args g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar[2],
      tmem_empty_bar[2], mask_bar[kMask ? 2 * kEpiWarps : 1];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t a_bytes = BM * BK * 2, b_bytes = uint32_t(g.bn / 2) * BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t staging = base + uint32_t(g.stages) * stage_bytes;  // 16 warps x bufs x 2 KB (4 KB: 64-column chunks)
  const uint32_t stg_bytes = g.epi_cols == 64 ? 4096u : 2048u;
  // the whole bias vector (n floats, zeros when absent) sits behind the staging buffers
  float* bias_s = reinterpret_cast<float*>(smem_raw + (staging - smem_u32(smem_raw)) +
                                           size_t(4 * g.epi_groups) * stg_bytes * size_t(g.bufs));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = g.n / g.bn;
  const int n_sub = g.bn / g.epi_cols;  // epilogue sub-chunks per tile
  const int64_t mb_per_batch = (g.m + 2 * BM - 1) / (2 * BM);
  const int64_t m_blocks = mb_per_batch * g.batch;   // row blocks of all batch slices
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // Tile order.  row_major_tiles: a pair owns whole 256-row blocks (c, c + n_clusters, ..) and walks
  // their N tiles one after the other, so it completes full output rows within a few tiles (DRAM
  // pages are written once) and re-reads its A rows from L2.  Otherwise tiles are dealt round-robin
  // with N fastest (better balance when there are few row blocks).
  const int64_t my_tiles = g.row_major_tiles
      ? (m_blocks > cluster_id ? ((m_blocks - 1 - cluster_id) / n_clusters + 1) * n_tiles : 0)
      : (m_blocks * n_tiles > cluster_id ? (m_blocks * n_tiles - 1 - cluster_id) / n_clusters + 1 : 0);
  auto tile_of = [&](int64_t idx, int& mb, int& nt) {
    if (g.row_major_tiles) {
      mb = int(cluster_id + (idx / n_tiles) * n_clusters);
      nt = int(idx % n_tiles);
    } else {
      const int64_t t = cluster_id + idx * n_clusters;
      mb = int(t / n_tiles);
      nt = int(t % n_tiles);
    }
  };
  const uint32_t tmem_cols = uint32_t(2 * g.bn);
  const uint32_t epi_arrivals = 2u * 4u * uint32_t(n_sub < g.epi_groups ? n_sub : g.epi_groups);

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
      mbar_init(smem_u32(&tmem_empty_bar[i]), epi_arrivals);
    }
    if constexpr (kMask) {
      tma_prefetch_desc(&mmap);
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(smem_u32(&mask_bar[i]), 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.n; i += kTc3Threads) bias_s[i] = g.bias ? __ldg(g.bias + i) : 0.0f;
  if (warp == 1) {  // the same warp of both CTAs allocates the pair's TMEM columns
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
      // stage index / phase kept incrementally: a division by the run-time ring depth costs this single thread ~100
      // cycles per K block
      uint32_t s = 0, ph = 0;
      bool wrapped = false;
      for (int64_t idx = 0; idx < my_tiles; ++idx) {
        int mb, nt;
        tile_of(idx, mb, nt);
        const int bi = int(mb / mb_per_batch);
        const int m0 = int(mb % mb_per_batch) * (2 * BM) + int(rank) * BM;
        const int n0 = nt * g.bn + int(rank) * (g.bn / 2);
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          if (wrapped) mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          if (leader) mbar_expect_tx(smem_u32(&full_bar[s]), 2 * stage_bytes);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          const uint32_t dst = base + s * stage_bytes;
          tma_load_3d_pair(dst, &amap, kb * BK, m0, bi, bar);
          if (!g.b_mn) {
            tma_load_3d_pair(dst + a_bytes, &bmap, kb * BK, n0, 0, bar);
          } else {  // 64 (n) x 64 (reduction) boxes, one per 64 output columns, 8 KB apart
            for (int i = 0; i < g.bn / 128; ++i)
              tma_load_3d_pair(dst + a_bytes + uint32_t(i) * 8192u, &bmap, n0 + 64 * i, kb * BK, 0, bar);
          }
          if (++s == uint32_t(g.stages)) { s = 0; ph ^= 1u; wrapped = true; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {  // ===== MMA issuer (leader CTA) =====
      const uint32_t idesc = make_idesc_pair(g.bn) | (g.b_mn ? (1u << 16) : 0u);  // bit 16: B MN-major
      uint32_t seq = 0, s = 0, ph = 0;
      for (int64_t idx = 0; idx < my_tiles; ++idx, ++seq) {
        const uint32_t acc = seq & 1u, use = seq >> 1;
        if (use > 0) mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use - 1) & 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_d + acc * uint32_t(g.bn);
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t a_addr = base + s * stage_bytes;
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = g.b_mn ? make_smem_desc_mn(a_addr + a_bytes, 8192u)
                                        : make_smem_desc(a_addr + a_bytes);
          // per UMMA_K = 16: K-major +32 B inside the swizzle row; MN-major +16 rows of 128 B
          const uint64_t bstep = g.b_mn ? uint64_t(2048 >> 4) : uint64_t(2);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            umma_f16_pair(d_addr, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk) * bstep, idesc,
                          (kb | kk) ? 1u : 0u);
          umma_commit_pair(smem_u32(&empty_bar[s]));
          if (++s == uint32_t(g.stages)) { s = 0; ph ^= 1u; }
        }
        umma_commit_pair(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===== epilogue warps 2..17: TMEM lanes 32*(warp%4) .. +31, 32-column sub-chunks g4, g4+4, .. =====
    const int q = warp & 3, g4 = (warp - 2) >> 2;
    const uint32_t my_stage = staging + uint32_t(warp - 2) * stg_bytes * uint32_t(g.bufs);
    const uint32_t row_off = uint32_t(lane) * 64u;
    const uint32_t sw = uint32_t(lane >> 1) & 3u;  // SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
    const bool relu = g.relu != 0;
    const uint32_t empty_remote0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_remote1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    uint32_t seq = 0, buf = 0;
    [[maybe_unused]] uint32_t mseq = 0;   // chunks this warp has masked: buffer mseq & 1, barrier phase (mseq >> 1) & 1
    if (g4 < n_sub && g4 < g.epi_groups) {
      if constexpr (kMask) {   // prologue: the first chunk's mask tile
        if (lane == 0 && my_tiles > 0) {
          int fmb, fnt;
          tile_of(0, fmb, fnt);
          const uint32_t mb_addr = smem_u32(&mask_bar[(warp - 2) * 2]);
          mbar_expect_tx(mb_addr, 2048u);
          tma_load_3d(my_stage, &mmap, fnt * g.bn + g4 * 32, int(fmb % mb_per_batch) * (2 * BM) + int(rank) * BM + q * 32,
                      int(fmb / mb_per_batch), mb_addr);
        }
      }
      for (int64_t idx = 0; idx < my_tiles; ++idx, ++seq) {
        int mb, nt;
        tile_of(idx, mb, nt);
        const int bi = int(mb / mb_per_batch);
        const int m0 = int(mb % mb_per_batch) * (2 * BM) + int(rank) * BM;
        const int n0 = nt * g.bn;
        const uint32_t acc = seq & 1u;
        mbar_wait_backoff(smem_u32(&tmem_full_bar[acc]), (seq >> 1) & 1u, g.wait_ns);
        tc_fence_after();
        const uint32_t t_addr = tmem_d + acc * uint32_t(g.bn) + (uint32_t(q * 32) << 16);
        if (g.epi_cols == 64) {
          // 64-column chunks: two TMEM loads fill one 32-row x 128-byte staging tile (SWIZZLE_128B), ONE TMA
          // store of 128-byte rows per chunk -- half as many stores, each row a full 128-byte line (the
          // write-bound shapes, K < 512, were limited by the store path at 64-byte rows)
          const uint32_t sw7 = uint32_t(lane) & 7u;
          for (int sc = g4; sc < n_sub; sc += g.epi_groups) {
            const int c = sc * 64;
            uint32_t r[32];
            tmem_ld32_nowait(t_addr + uint32_t(c), r);
            if (lane == 0) {
              if (g.bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            __syncwarp();
            const uint32_t sbuf = my_stage + buf * stg_bytes + uint32_t(lane) * 128u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + c + 32 * half);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b0 = bp[2 * j], b1 = bp[2 * j + 1];
                uint4 o;
                o.x = bias_pack(r[8 * j + 0], r[8 * j + 1], b0.x, b0.y, relu);
                o.y = bias_pack(r[8 * j + 2], r[8 * j + 3], b0.z, b0.w, relu);
                o.z = bias_pack(r[8 * j + 4], r[8 * j + 5], b1.x, b1.y, relu);
                o.w = bias_pack(r[8 * j + 6], r[8 * j + 7], b1.z, b1.w, relu);
                sts_v4(sbuf + ((uint32_t(4 * half + j) ^ sw7) << 4), o);
              }
              if (half == 0) {
                tmem_ld32_nowait(t_addr + uint32_t(c + 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (sc + g.epi_groups >= n_sub) {  // last TMEM read of this tile by this warp
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive_cluster(acc ? empty_remote1 : empty_remote0);
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&ymap, my_stage + buf * stg_bytes, n0 + c, m0 + q * 32, bi);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf = (buf + 1u) & uint32_t(g.bufs - 1);
          }
          continue;
        }
        for (int sc = g4; sc < n_sub; sc += g.epi_groups) {
          const int c = sc * 32;
          uint32_t r[32];
          tmem_ld32_nowait(t_addr + uint32_t(c), r);
          if constexpr (kMask) {
            // mask tile of THIS chunk is in flight into staging buffer `buf` (issued one chunk ago, or by the
            // prologue); the OTHER buffer's store (previous chunk) must have been read before the next chunk's mask
            // tile lands in it
            if (lane == 0) {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              int64_t nidx = idx;
              int nsc = sc + g.epi_groups;
              if (nsc >= n_sub) { nsc = g4; ++nidx; }
              if (nidx < my_tiles) {
                int nmb, nnt;
                tile_of(nidx, nmb, nnt);
                const uint32_t nb = buf ^ 1u;
                const uint32_t mb_addr = smem_u32(&mask_bar[(warp - 2) * 2 + int(nb)]);
                mbar_expect_tx(mb_addr, 2048u);
                tma_load_3d(my_stage + nb * 2048u, &mmap, nnt * g.bn + nsc * 32,
                            int(nmb % mb_per_batch) * (2 * BM) + int(rank) * BM + q * 32, int(nmb / mb_per_batch), mb_addr);
              }
            }
          } else
          // the staging buffer we are about to overwrite must have been read by its TMA store
          if (lane == 0) {
            if (g.bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (sc + g.epi_groups >= n_sub) {  // last TMEM read of this tile by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? empty_remote1 : empty_remote0);
          }
          __syncwarp();
          const uint32_t sbuf = my_stage + buf * 2048u + row_off;
          // bias slice of this sub-chunk (same address for every lane: broadcast LDS.128)
          const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + c);
          if constexpr (kMask) {   // the chunk's mask tile has landed in the buffer we are about to pack into
            mbar_wait(smem_u32(&mask_bar[(warp - 2) * 2 + int(buf)]), (mseq >> 1) & 1u);
            ++mseq;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 4 chunks of 8 columns = 16 bytes
            const float4 b0 = bp[2 * j], b1 = bp[2 * j + 1];
            uint4 o;
            o.x = bias_pack(r[8 * j + 0], r[8 * j + 1], b0.x, b0.y, relu);
            o.y = bias_pack(r[8 * j + 2], r[8 * j + 3], b0.z, b0.w, relu);
            o.z = bias_pack(r[8 * j + 4], r[8 * j + 5], b1.x, b1.y, relu);
            o.w = bias_pack(r[8 * j + 6], r[8 * j + 7], b1.z, b1.w, relu);
            if constexpr (kMask) {
              const uint4 mk = lds_v4_tc3(sbuf + ((uint32_t(j) ^ sw) << 4));
              o.x &= gt0_mask(mk.x);
              o.y &= gt0_mask(mk.y);
              o.z &= gt0_mask(mk.z);
              o.w &= gt0_mask(mk.w);
            }
            sts_v4(sbuf + ((uint32_t(j) ^ sw) << 4), o);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&ymap, my_stage + buf * 2048u, n0 + c, m0 + q * 32, bi);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          buf = (buf + 1u) & uint32_t(g.bufs - 1);
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading our operands / arriving on our barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(tmem_cols)
                 : "memory");
  }
}

}  // namespace

int linear_tc3_supported(int64_t m, int64_t k, int64_t n_out, int b_mn) {
  static const bool disabled = getenv("GWEN_TC_NO_PAIR") != nullptr;
  if (b_mn && n_out % 128) return 0;  // each CTA of the pair stages whole 64-column boxes
  return !disabled && n_out % 64 == 0 && n_out <= 8192 && m >= 256 && k >= 8 && sm_count() % 2 == 0;
}

int linear_tc3_fwd_bf16(const void* x, const void* w, void* y, int64_t m, int64_t k, int64_t n_out,
                        int64_t ldx, int64_t ldw, int64_t ldy, const float* bias, int relu,
                        int b_mn, cudaStream_t st, int64_t batch, int64_t x_bstride, int64_t y_bstride,
                        const void* mask, int64_t ldm, int64_t m_bstride) {
  if (batch < 1 || batch > 65535) return set_err(GWEN_E_BADARG, "batch out of range");
  int bn = 0;
  for (int c : {256, 128, 64})
    if (n_out % c == 0) { bn = c; break; }
  if (!bn) return set_err(GWEN_E_NOSUPPORT, "tcgen05 pair GEMM needs n_out %% 64 == 0");
  CUtensorMap amap, bmap, ymap;
  int rc = make_tensor_map_3d(&amap, x, GWEN_BF16, k, m, batch, ldx, x_bstride, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = b_mn ? make_tensor_map_3d(&bmap, w, GWEN_BF16, n_out, k, 1, ldw, 0, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B)
            : make_tensor_map_3d(&bmap, w, GWEN_BF16, k, n_out, 1, ldw, 0, BK, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  const int k_blocks = static_cast<int>(ceil_div(k, BK));
  const size_t stage_bytes = size_t(BM + bn / 2) * BK * 2;
  // staging buffers per epilogue warp: two (one TMA store in flight while the next chunk packs) for the shallow,
  // store-bound reductions -- measured at M = 896 292 (round 2, 1 / 2 buffers): 64->1024 0.359 / 0.326 ms (cuBLAS
  // 0.310), 1024->64 0.390 / 0.357, but 256->512 0.283 / 0.309 and 1024->512 0.801 / 0.812 (ring depth matters more)
  static const int bufs_env = [] {
    const char* v = getenv("GWEN_TC3_BUFS");
    return v ? atoi(v) : 0;
  }();
  const int bufs = mask ? 2 : ((bufs_env == 1 || bufs_env == 2) ? bufs_env : ((k <= 128 || n_out <= 64) ? 2 : 1));
  // deep reductions are MMA-bound and want ring depth (6 stages need the room of 8 staging buffers);
  // shallow ones are store-bound and want all 16 epilogue warps
  static const int groups_env = [] {
    const char* v = getenv("GWEN_TC3_EPI_GROUPS");
    return v ? atoi(v) : 0;
  }();
  // (masked dgrad: the epilogue does more per chunk and wants all 16 warps -- 512->1024 998 -> 936 us)
  const int epi_groups = (groups_env == 2 || groups_env == 4) ? groups_env : ((k >= 512 && !mask) ? 2 : 4);
  // 64-column chunks (128-byte store rows) only for the narrow output tile (bn = 64: one chunk per warp).  Measured
  // at M = 896 292 (round 2, 32 / 64 columns): 64->1024 0.359 / 0.416 ms, 256->512 0.283 / 0.342, 1024->64 0.390 / 0.369
  // -- the wider chunk serialises two TMEM loads per store and costs a ring stage; the store row width was not the limit.
  static const int cols_env = [] {
    const char* v = getenv("GWEN_TC3_EPI_COLS");
    return v ? atoi(v) : 0;
  }();
  const int epi_cols = mask ? 32 : ((cols_env == 32 || cols_env == 64) ? (bn % cols_env ? 32 : cols_env)
                                                                       : (epi_groups == 4 && bn == 64 ? 64 : 32));
  rc = epi_cols == 64
           ? make_tensor_map_3d(&ymap, y, GWEN_BF16, n_out, m, batch, ldy, y_bstride, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)
           : make_tensor_map_3d(&ymap, y, GWEN_BF16, n_out, m, batch, ldy, y_bstride, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != GWEN_OK) return rc;
  CUtensorMap mmap = ymap;
  if (mask) {
    rc = make_tensor_map_3d(&mmap, mask, GWEN_BF16, n_out, m, batch, ldm, m_bstride, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != GWEN_OK) return rc;
  }
  const size_t staging_bytes = size_t(4 * epi_groups) * bufs * (epi_cols == 64 ? 4096 : 2048) + align_up(size_t(n_out) * 4, 1024);
  static const int stage_cap = [] {
    const char* v = getenv("GWEN_TC3_STAGES");
    return v ? std::max(2, std::min(kMaxStages, atoi(v))) : kMaxStages;
  }();
  int stages = static_cast<int>(std::min<size_t>(stage_cap, (226 * 1024 - staging_bytes - 1024) / stage_bytes));
  if (stages < 2) return set_err(GWEN_E_NOSUPPORT, "tile does not fit in shared memory");
  // >= 120 KB keeps one CTA per SM (a pair owns up to all 512 TMEM columns of both SMs)
  const size_t smem = std::max<size_t>(stages * stage_bytes + staging_bytes + 1024, 120 * 1024);
  const int64_t total = ceil_div(m, 2 * BM) * batch * (n_out / bn);
  const int pairs = static_cast<int>(std::min<int64_t>(total, std::max(1, (sm_count() - sm_reserve()) / 2)));
  static const int order_env = [] {
    const char* v = getenv("GWEN_TC3_ORDER");
    return v ? atoi(v) : -1;
  }();
  // measured at the GWEN shapes (M = 896 292): both orders within 3 %; round-robin is the default
  const int row_major = order_env >= 0 ? order_env : 0;
  Tc3Args g{bias, m, static_cast<int>(n_out), k_blocks, bn, stages, relu, bufs, row_major, epi_groups, static_cast<int>(batch), b_mn, epi_cols, wait_backoff_ns()};
  if (mask) {
    GWEN_CUDA(cudaFuncSetAttribute(k_linear_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    k_linear_tc3<true><<<2 * pairs, kTc3Threads, smem, st>>>(amap, bmap, ymap, mmap, g);
  } else {
    GWEN_CUDA(cudaFuncSetAttribute(k_linear_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    k_linear_tc3<false><<<2 * pairs, kTc3Threads, smem, st>>>(amap, bmap, ymap, mmap, g);
  }
  GWEN_LAUNCH_CHECK("k_linear_tc3");
  return GWEN_OK;
}

}  // namespace gwen
