// K1 + K2 in one kernel for mesh graphs (bf16):  Y = epi( (A_hat X) W^T + b ).
//
// The layers that aggregate before they project (in_channels < out_channels: conv1, upconv3,
// upconv4 of the GWEN stack) used to run k_grid_stencil (write A_hat X to HBM) and then the GEMM
// (read it back).  Here the aggregated rows never leave the SM: CUDA-core warps run the separable
// mesh stencil on TMA-staged source boxes and write their result straight into shared memory in
// the K-major SWIZZLE_128B layout the tensor core reads as its A operand; the projection is the
// CTA-pair tcgen05 pipeline of linear_tc3.cu (W tiles by TMA, M = 256 MMAs issued by the leader,
// two TMEM accumulators, TMA-store epilogue with bias / ReLU).
//
// Geometry: a CTA owns an 8 x 16 block of destination nodes = 128 A rows (row i = 16 * tr + tc =
// TMEM lane i); a pair owns 8 x 32.  Work item = (member b, tile row, pair column).
// Per CTA (19 warps):
//   warp 0 lane 0    source producer: per 64-feature slab ONE 4-D TMA box {64, 18, 10, 1} of x viewed
//                    as [B][H][W][K] (zero fill outside the mesh = the truncated stencil of border
//                    nodes) + the 10 dis rows, two stages; K block kb uses stage kb & 1, which feeds
//                    stencil group kb & 1.
//   warp 1 lane 0    MMA issuer (leader CTA): per N tile, per K block: waits a_full[kb] (both CTAs'
//                    stencil warps) and b_full[stage], 4 x tcgen05.mma.cta_group::2; commits free the
//                    W stage, (on the last N tile) the A block, and publish the accumulator.
//   warp 2 lane 0    W producer (both CTAs: its half of the BN x 64 tile per K block).
//   warps 3..10      two stencil groups of 4 warps: sub-warps of 8 lanes (= 64 bf16 features) slide
//                    along 8 destinations of a tile row with the column sums in packed fp32x2
//                    registers (as k_grid_stencil), round to bf16, st.shared into A block kb,
//                    fence.proxy.async, arrive on the leader's a_full[kb].
//   warps 11..18     epilogue: tcgen05.ld -> bias -> bf16 -> ReLU -> 64B-swizzled st.shared -> one 4-D
//                    TMA store {32 cols, 16, 2, 1} per warp and 32-column sub-chunk (the store clips
//                    tiles that overhang the mesh).
// A operand: 8 resident K blocks of 16 KB.  k_in <= 256: two buffers of k_blocks blocks, the stencil of item
// t+1 fills one while the MMAs of item t read the other.  k_in = 320 .. 512: ONE buffer; every block is
// released (tcgen05.commit -> a_empty[kb]) as soon as the LAST N tile's MMAs on it have retired, so the stencil
// of item t+1 refills the blocks behind the MMAs of item t's last N pass.
// Round 2: (1) the source producer issues a TMA L2 PREFETCH of a later item's boxes (ncu, round 1: the stencil
// warps spent 35 % of their time on src_full -- with one stage per stencil group the HBM latency of every box
// was exposed; a prefetched box arrives from L2); (2) the stencil warps can finish the PREVIOUS layer on the
// aggregated row -- a = relu(A_hat p + pre_bias), rounded to bf16 exactly as k_grid_stencil stores it -- so a
// layer that aggregates last can hand its un-aggregated projection p straight to the next layer's fused
// kernel:  Y = epi( relu(A_hat P + b_prev) W^T + b ), and the aggregated tensor of that layer never exists in HBM.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"
#include "stencil_common.cuh"

namespace gwen {
using namespace tc;
using namespace st;
namespace {

constexpr int kFusedThreads = 19 * 32;
constexpr int kMaxKb = 8;      // resident A blocks of 16 KB: two buffers for K <= 256, one for K <= 512
constexpr int kMaxSB = 6;      // W stage ring
constexpr int FT_W = 16;       // tile width (destinations), TH = 8 rows -> 128 A rows
constexpr uint32_t kSrcRow = (FT_W + 2) * 128u;                  // one staged mesh row of a slab
constexpr uint32_t kSrcBox = (TH + 2) * kSrcRow;                 // 23040 B
constexpr uint32_t kDisRow = 80u;                                // 18 floats padded to 20
constexpr uint32_t kSrcStage = 24576u;                           // box + 10 dis rows, 1 KB aligned
constexpr uint32_t kABlock = BM * 128u;                          // 128 rows x 64 bf16
constexpr uint32_t kRingBlocks = 4;                              // A block ring of the one-N-tile mode (power of two)

struct FusedArgs {
  const float* bias;
  const float* disb;        // zero-bordered dis (graph.bordered_dis)
  int64_t disb_pitch;
  int batch, h, w, k_blocks, n, bn, sb, relu;
  int tiles_y, pairs_x;
  int nbuf;                 // resident A buffers: 2 (k_blocks <= 4) or 1
  const float* pre_bias;    // [k_in] added to the aggregated row before it becomes the A operand (nullable)
  int pre_relu;             // ReLU on the aggregated row (after pre_bias)
  int prefetch;             // L2 prefetch distance of the source boxes, in items (0 = off)
  int epi_groups;           // epilogue warp groups (of 4 warps) that work: 2, or 1 when the A blocks fill the SM
  int epi_bufs;             // 2 KB staging buffers per epilogue warp: 2 = one TMA store in flight while the next packs
  uint32_t wait_ns;         // poll interval of the stencil / epilogue warps on their barriers (0 = spin)
  int ring;                 // 1: ONE N tile (n_out == bn), so an A block is read by exactly one K step: the A region is a
                            // ring of kRingBlocks blocks (global block b -> slot b % 4, stencil group b & 1) instead of
                            // a resident tile, and the shared memory it frees holds two source boxes per stencil group
  int sg;                   // source stages per stencil group (1 or 2)
  int kouter;               // ring mode with TWO N tiles whose accumulators are both resident (2 bn = 512 TMEM columns):
                            // K block outer, N tile inner, so an A block is read by two consecutive K steps and freed
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFusedThreads, 1)
    k_gcn_fused(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                const __grid_constant__ CUtensorMap ymap, FusedArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t src_full[4], src_empty[4], a_full[kMaxKb], a_empty[kMaxKb],
      b_full[kMaxSB], b_empty[kMaxSB], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // shared-memory carve-up
  const uint32_t a_base = base;                                            // nbuf buffers x k_blocks x 16 KB
  const uint32_t a_buf_bytes = uint32_t(g.k_blocks) * kABlock;
  const uint32_t src_base = a_base + (g.ring ? kRingBlocks * kABlock : uint32_t(g.nbuf) * a_buf_bytes);  // 2 sg x 24 KB
  const uint32_t SG = uint32_t(g.sg);
  const uint32_t nbuf = uint32_t(g.nbuf);
  const int kbs = g.k_blocks;                                              // barrier index = abuf * kbs + kb
  const uint32_t b_bytes = uint32_t(g.bn / 2) * 128u;
  const uint32_t b_base = src_base + 2u * SG * kSrcStage;                  // sb x b_bytes
  const uint32_t epi_base = b_base + uint32_t(g.sb) * b_bytes;             // 8 warps x 2 KB
  float* bias_s = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw)) +
                                           uint32_t(4 * g.epi_groups * g.epi_bufs) * 2048u);

  const int n_tiles = g.n / g.bn;
  const int n_sub = g.bn / 32;
  const int64_t items = int64_t(g.batch) * g.tiles_y * g.pairs_x;
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t my_items = items > pair_id ? (items - 1 - pair_id) / n_pairs + 1 : 0;
  const uint32_t tmem_cols = uint32_t(2 * g.bn);
  auto item_of = [&](int64_t seq, int& b, int& r0, int& c0) {
    const int64_t it = pair_id + seq * n_pairs;
    const int per_b = g.tiles_y * g.pairs_x;
    b = int(it / per_b);
    const int rem = int(it % per_b);
    r0 = (rem / g.pairs_x) * TH;
    c0 = (rem % g.pairs_x) * (2 * FT_W) + int(rank) * FT_W;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&wmap);
    tma_prefetch_desc(&ymap);
    for (int i = 0; i < 2 * g.sg; ++i) {
      mbar_init(smem_u32(&src_full[i]), 1);
      mbar_init(smem_u32(&src_empty[i]), 4);           // the 4 warps of the stencil group that owns the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 2u * 4u * uint32_t(n_sub < g.epi_groups ? n_sub : g.epi_groups));
    }
    for (int i = 0; i < kMaxKb; ++i) {                 // [A buffer][K block]
      mbar_init(smem_u32(&a_full[i]), 8);              // 4 stencil warps x 2 CTAs (leader's is used)
      mbar_init(smem_u32(&a_empty[i]), 1);
    }
    for (int i = 0; i < g.sb; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.n; i += kFusedThreads) bias_s[i] = g.bias ? __ldg(g.bias + i) : 0.0f;
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
    if (lane == 0) {  // ===== source producer =====
      // Block (seq, kb) goes to stencil group kb & 1 (ring mode: global block index & 1), whose private ring has SG
      // stages: a group then sees EVERY phase of the barriers it waits on (a parity wait must never skip a phase).
      uint32_t cnt[2] = {0u, 0u};
      for (int64_t seq = 0; seq < my_items; ++seq) {
        int b, r0, c0;
        item_of(seq, b, r0, c0);
        if (g.prefetch > 0 && seq + g.prefetch < my_items) {   // pull a later item's boxes into L2 now
          int pb, pr0, pc0;
          item_of(seq + g.prefetch, pb, pr0, pc0);
          for (int kb = 0; kb < g.k_blocks; ++kb)
            asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(&xmap),
                         "r"(kb * BK), "r"(pc0 - 1), "r"(pr0 - 1), "r"(pb)
                         : "memory");
        }
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          const uint32_t grp = g.ring ? uint32_t(seq * g.k_blocks + kb) & 1u : uint32_t(kb) & 1u;
          const uint32_t jj = cnt[grp]++;
          const uint32_t s = grp * SG + (jj & (SG - 1u)), round = jj >> (SG - 1u);   // SG is 1 or 2
          if (round > 0) mbar_wait(smem_u32(&src_empty[s]), (round - 1) & 1u);
          const uint32_t bar = smem_u32(&src_full[s]);
          const uint32_t dst = src_base + s * kSrcStage;
          mbar_expect_tx(bar, kSrcBox + uint32_t(TH + 2) * kDisRow);
          tma_load_4d(dst, &xmap, kb * BK, c0 - 1, r0 - 1, b, bar);
          // bordered dis: element [r + 1][c + 1] = dis[r][c]; mesh node (r0 - 1 + i, c0 - 1 + t)
          const float* dsrc = g.disb + int64_t(r0) * g.disb_pitch + c0;
          for (int i = 0; i < TH + 2; ++i)
            bulk_g2s(dst + kSrcBox + uint32_t(i) * kDisRow, dsrc + int64_t(i) * g.disb_pitch, kDisRow, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_pair(g.bn);
      // W ring stage / phase and the A buffer index are kept incrementally: a division by a run-time ring depth costs
      // this single thread ~100 cycles per K block (measured in linear_b2b.cu: a third of the issuer's time)
      uint32_t seq_n = 0, ws = 0, wph = 0;
      const uint32_t sbn = uint32_t(g.sb);
      if (g.kouter) {
        // both accumulators (N tiles 0 and 1) stay resident for the whole item; accumulator nt is tile nt
        for (int64_t seq = 0; seq < my_items; ++seq) {
          for (int kb = 0; kb < g.k_blocks; ++kb) {
            const uint32_t blk = uint32_t(seq) * uint32_t(g.k_blocks) + uint32_t(kb);
            const uint32_t aslot = blk % kRingBlocks;
            mbar_wait(smem_u32(&a_full[aslot]), (blk / kRingBlocks) & 1u);
            const uint64_t adesc = make_smem_desc(a_base + aslot * kABlock);
            for (int nt = 0; nt < 2; ++nt) {
              // the epilogue must have drained this accumulator of the previous item before its first MMA
              if (kb == 0 && seq > 0) mbar_wait(smem_u32(&tmem_empty_bar[nt]), uint32_t(seq - 1) & 1u);
              const uint32_t s = ws;
              mbar_wait(smem_u32(&b_full[s]), wph);
              if (++ws == sbn) { ws = 0; wph ^= 1u; }
              tc_fence_after();
              const uint64_t bdesc = make_smem_desc(b_base + s * b_bytes);
              const uint32_t d_addr = tmem_d + uint32_t(nt) * uint32_t(g.bn);
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk)
                umma_f16_pair(d_addr, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc, (kb | kk) ? 1u : 0u);
              umma_commit_pair(smem_u32(&b_empty[s]));
            }
            umma_commit_pair(smem_u32(&a_empty[aslot]));       // both N tiles have read the block
          }
          umma_commit_pair(smem_u32(&tmem_full_bar[0]));
          umma_commit_pair(smem_u32(&tmem_full_bar[1]));
        }
      } else
      for (int64_t seq = 0; seq < my_items; ++seq) {
        const uint32_t abuf = nbuf == 2 ? uint32_t(seq) & 1u : 0u;
        const uint32_t aphase = (nbuf == 2 ? uint32_t(seq >> 1) : uint32_t(seq)) & 1u;
        for (int nt = 0; nt < n_tiles; ++nt, ++seq_n) {
          const uint32_t acc = seq_n & 1u, use = seq_n >> 1;
          if (use > 0) mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use - 1) & 1u);
          tc_fence_after();
          const uint32_t d_addr = tmem_d + acc * uint32_t(g.bn);
          for (int kb = 0; kb < g.k_blocks; ++kb) {
            // A block of (seq, kb): resident slot abuf * kbs + kb, or (ring mode) slot blk % ra of the block ring
            const uint32_t blk = uint32_t(seq) * uint32_t(g.k_blocks) + uint32_t(kb);
            const uint32_t aslot = g.ring ? blk % kRingBlocks : abuf * uint32_t(kbs) + uint32_t(kb);
            const uint32_t afull_phase = g.ring ? (blk / kRingBlocks) & 1u : aphase;
            if (nt == 0) mbar_wait(smem_u32(&a_full[aslot]), afull_phase);
            const uint32_t s = ws;
            mbar_wait(smem_u32(&b_full[s]), wph);
            if (++ws == sbn) { ws = 0; wph ^= 1u; }
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(a_base + aslot * kABlock);
            const uint64_t bdesc = make_smem_desc(b_base + s * b_bytes);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk)
              umma_f16_pair(d_addr, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc,
                            (kb | kk) ? 1u : 0u);
            umma_commit_pair(smem_u32(&b_empty[s]));
            if (nt == n_tiles - 1) umma_commit_pair(smem_u32(&a_empty[aslot]));  // block is free
          }
          umma_commit_pair(smem_u32(&tmem_full_bar[acc]));
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {  // ===== W producer (both CTAs) =====
      uint32_t s = 0, ph = 0;
      bool wrapped = false;
      const int outer = g.kouter ? g.k_blocks : n_tiles, inner = g.kouter ? n_tiles : g.k_blocks;
      for (int64_t seq = 0; seq < my_items; ++seq) {
        for (int o = 0; o < outer; ++o) {
          for (int i = 0; i < inner; ++i) {     // the order in which the MMA issuer consumes the W tiles
            const int nt = g.kouter ? i : o, kb = g.kouter ? o : i;
            const int n0 = nt * g.bn + int(rank) * (g.bn / 2);
            if (wrapped) mbar_wait(smem_u32(&b_empty[s]), ph ^ 1u);
            if (leader) mbar_expect_tx(smem_u32(&b_full[s]), 2 * b_bytes);
            tma_load_3d_pair(b_base + s * b_bytes, &wmap, kb * BK, n0, 0, mapa_u32(smem_u32(&b_full[s]), 0));
            if (++s == uint32_t(g.sb)) { s = 0; ph ^= 1u; wrapped = true; }
          }
        }
      }
    }
  } else if (warp < 11) {
    // ===== stencil groups: warps 3..6 -> stage 0, warps 7..10 -> stage 1 =====
    const int grp = (warp - 3) >> 2, wg = (warp - 3) & 3;
    const int sub = lane >> 3, l = lane & 7;               // 4 sub-warps of 8 lanes (64 features)
    const int su = wg * 4 + sub;                           // 0..15: (tile row, half row)
    const int tr = su >> 1, cb = (su & 1) * SEG;
    const int row0 = tr * FT_W + cb;                       // first A row of this unit
    const bool pre_relu = g.pre_relu != 0;
    uint32_t jj = 0;                                       // blocks this group has aggregated so far
    for (int64_t seq = 0; seq < my_items; ++seq)
    for (int kb = 0; kb < g.k_blocks; ++kb) {
      const uint32_t blk = uint32_t(seq) * uint32_t(g.k_blocks) + uint32_t(kb);
      if ((g.ring ? blk & 1u : uint32_t(kb) & 1u) != uint32_t(grp)) continue;   // the other group's block
      const uint32_t sstage = uint32_t(grp) * SG + (jj & (SG - 1u)), sphase = (jj >> (SG - 1u)) & 1u;   // SG is 1 or 2
      ++jj;
      const uint32_t xs = src_base + sstage * kSrcStage;
      const uint32_t ds = xs + kSrcBox;
      const uint32_t x0 = xs + uint32_t(tr) * kSrcRow + uint32_t(cb) * 128u + uint32_t(l) * 16u;
      const uint32_t x1 = x0 + kSrcRow, x2 = x1 + kSrcRow;
      const uint32_t q0 = ds + uint32_t(tr) * kDisRow + uint32_t(cb) * 4u;
      const uint32_t q1 = q0 + kDisRow, q2 = q1 + kDisRow;
      const uint32_t abuf = nbuf == 2 ? uint32_t(seq) & 1u : 0u;
      // A block slot and how often it has been used before (ring mode: global block ring)
      const uint32_t aslot = g.ring ? blk % kRingBlocks : abuf * uint32_t(kbs) + uint32_t(kb);
      const uint32_t use = g.ring ? blk / kRingBlocks : (nbuf == 2 ? uint32_t(seq >> 1) : uint32_t(seq));
      // per-feature bias of the aggregated row (the previous layer's epilogue moved into this producer)
      uint64_t pb2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        pb2[k] = g.pre_bias ? pk2(__ldg(g.pre_bias + kb * BK + l * 8 + 2 * k), __ldg(g.pre_bias + kb * BK + l * 8 + 2 * k + 1))
                            : 0ull;
      mbar_wait_backoff(smem_u32(&src_full[sstage]), sphase, g.wait_ns);
      // this A block was used before: wait until the MMAs of its last N tile have retired
      if (use > 0) mbar_wait_backoff(smem_u32(&a_empty[aslot]), (use - 1u) & 1u, g.wait_ns);
      const uint32_t a_blk = a_base + aslot * kABlock;
      uint64_t s0[4], s1[4], s2[4];
      float dmid_prev = 0.0f;
#pragma unroll
      for (int tt = 0; tt < SEG + 2; ++tt) {
        const uint4 v0 = lds128(x0 + uint32_t(tt) * 128u);
        const uint4 v1 = lds128(x1 + uint32_t(tt) * 128u);
        const uint4 v2 = lds128(x2 + uint32_t(tt) * 128u);
        const float d0 = lds32(q0 + uint32_t(tt) * 4u);
        const float d1 = lds32(q1 + uint32_t(tt) * 4u);
        const float d2 = lds32(q2 + uint32_t(tt) * 4u);
        uint64_t f0[4], f1[4], f2[4];
        V16<__nv_bfloat16>::unpack2(v0, f0);
        V16<__nv_bfloat16>::unpack2(v1, f1);
        V16<__nv_bfloat16>::unpack2(v2, f2);
        const uint64_t e0 = pk2(d0, d0), e1 = pk2(d1, d1), e2 = pk2(d2, d2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s0[k] = s1[k];
          s1[k] = s2[k];
          s2[k] = fma2(e2, f2[k], fma2(e1, f1[k], mul2(e0, f0[k])));
        }
        if (tt >= 2) {  // destination column cb + tt - 2 -> A row row0 + tt - 2
          const uint64_t dm = pk2(dmid_prev, dmid_prev);
          uint64_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = fma2(dm, add2(add2(s0[k], s1[k]), s2[k]), pb2[k]);
          const uint32_t i = uint32_t(row0 + tt - 2);
          sts_v4(a_blk + i * 128u + ((uint32_t(l) ^ (i & 7u)) << 4), V16<__nv_bfloat16>::pack2(o, pre_relu));
        }
        dmid_prev = d1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&src_empty[sstage]));
        mbar_arrive_cluster(mapa_u32(smem_u32(&a_full[aslot]), 0));
      }
    }
  } else {
    // ===== epilogue warps 11..18: TMEM lanes 32 (warp % 4) .. +31 = tile rows 2q, 2q + 1 =====
    const int q = warp & 3, g2 = (warp - 11) >> 2;
    const uint32_t my_stage0 = epi_base + uint32_t(warp - 11) * 2048u * uint32_t(g.epi_bufs);
    uint32_t ebuf = 0;
    const uint32_t row_off = uint32_t(lane) * 64u;
    const uint32_t sw = uint32_t(lane >> 1) & 3u;  // SWIZZLE_64B
    const bool relu = g.relu != 0;
    const uint32_t empty_remote0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t empty_remote1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    uint32_t seq_n = 0;
    if (g2 < n_sub && g2 < g.epi_groups) {
      for (int64_t seq = 0; seq < my_items; ++seq) {
        int b, r0, c0;
        item_of(seq, b, r0, c0);
        for (int nt = 0; nt < n_tiles; ++nt, ++seq_n) {
          const int n0 = nt * g.bn;
          const uint32_t acc = seq_n & 1u;
          mbar_wait_backoff(smem_u32(&tmem_full_bar[acc]), (seq_n >> 1) & 1u, g.wait_ns);
          tc_fence_after();
          const uint32_t t_addr = tmem_d + acc * uint32_t(g.bn) + (uint32_t(q * 32) << 16);
          for (int sc = g2; sc < n_sub; sc += g.epi_groups) {
            const int c = sc * 32;
            uint32_t r[32];
            tmem_ld32_nowait(t_addr + uint32_t(c), r);
            const uint32_t my_stage = my_stage0 + ebuf * 2048u;
            if (lane == 0) {   // the staging buffer about to be overwritten must have been read by its TMA store
              if (g.epi_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            ebuf = (ebuf + 1u) & uint32_t(g.epi_bufs - 1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (sc + g.epi_groups >= n_sub) {  // last TMEM read of this N tile by this warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(acc ? empty_remote1 : empty_remote0);
            }
            __syncwarp();
            const uint32_t sbuf = my_stage + row_off;
            const float4* bp = reinterpret_cast<const float4*>(bias_s + n0 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = bp[2 * j], b1 = bp[2 * j + 1];
              uint4 o;
              o.x = bias_pack(r[8 * j + 0], r[8 * j + 1], b0.x, b0.y, relu);
              o.y = bias_pack(r[8 * j + 2], r[8 * j + 3], b0.z, b0.w, relu);
              o.z = bias_pack(r[8 * j + 4], r[8 * j + 5], b1.x, b1.y, relu);
              o.w = bias_pack(r[8 * j + 6], r[8 * j + 7], b1.z, b1.w, relu);
              sts_v4(sbuf + ((uint32_t(j) ^ sw) << 4), o);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {  // rows 2q, 2q + 1 of the tile: box {32 cols, 16 nodes, 2 mesh rows, 1}
              tma_store_4d(&ymap, my_stage, n0 + c, c0, r0 + 2 * q, b);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
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
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_gcn_fused_fwd(const void* x, const void* weight, void* y, const float* dis_padded,
                                  int64_t dis_pitch, int64_t dis_rows, int64_t batch, int64_t h, int64_t w, int64_t k_in,
                                  int64_t n_out, int dtype, const float* bias, int epilogue,
                                  const float* pre_bias, int pre_epilogue, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && h >= 0 && w >= 0 && k_in >= 0 && n_out >= 0, "negative size");
  if (batch == 0 || h == 0 || w == 0 || n_out == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && weight && y && dis_padded, "null pointer");
  if (dtype != GWEN_BF16) return set_err(GWEN_E_NOSUPPORT, "the fused layer kernel is bf16 only");
  if (k_in < 64 || k_in % 64 || k_in > 64 * kMaxKb || n_out % 128 || n_out > 8192)
    return set_err(GWEN_E_NOSUPPORT, "fused layer needs k_in in {64, 128, .., 512} and n_out %% 128 == 0");
  if (!aligned16(x) || !aligned16(weight) || !aligned16(y) || !aligned16(dis_padded) || dis_pitch % 4)
    return set_err(GWEN_E_ALIGN, "fused layer needs 16-byte aligned tensors");
  if (h * w >= INT32_MAX || batch >= 65536 || sm_count() % 2)
    return set_err(GWEN_E_NOSUPPORT, "mesh too large for the fused layer kernel");
  const int k_blocks = static_cast<int>(k_in / 64);
  const int tiles_y = static_cast<int>(ceil_div(h, TH)), pairs_x = static_cast<int>(ceil_div(w, 2 * FT_W));
  GWEN_CHECK_ARG(dis_pitch >= int64_t(pairs_x) * 2 * FT_W + 4, "bordered dis pitch too small");
  GWEN_CHECK_ARG(dis_rows >= int64_t(tiles_y) * TH + 2, "bordered dis has too few rows");
  // two A buffers while they fit in 8 blocks (k_in <= 256), else one
  const int nbuf = 2 * k_blocks <= kMaxKb ? 2 : 1;
  // double-buffered epilogue staging when the A blocks leave room (k_in <= 128: the store-bound layers, 64 -> 1024)
  static const int ebufs_env = [] {
    const char* v = getenv("GWEN_FUSED_EPI_BUFS");
    return v ? atoi(v) : 0;
  }();
  const int epi_bufs = (ebufs_env == 1 || ebufs_env == 2) ? ebufs_env : (k_blocks <= 2 ? 2 : 1);
  static const int ring_env = [] {   // GWEN_FUSED_RING=0: never use the A ring; 1: one-N-tile layers only (A/B tests)
    const char* v = getenv("GWEN_FUSED_RING");
    return v ? atoi(v) : 2;
  }();
  const size_t cap = 226 * 1024;
  // Configuration: N tile bn in {256, 128}; ring mode when the layer has ONE N tile (n_out == bn: an A block is read by
  // a single K step, so 4 ring blocks replace the resident tile); two source stages per stencil group when they fit
  // next to at least two W stages.  Among the candidates: the one that keeps more W bytes in flight, 256 on a tie.
  static const int sg_env = [] {     // GWEN_FUSED_SG=1|2: force the source stages per stencil group (sweeps)
    const char* v = getenv("GWEN_FUSED_SG");
    return v ? atoi(v) : 0;
  }();
  int bn = 0, ring = 0, sg = 1, epi_groups = 2, kouter = 0;
  size_t best = 0, fixed = 0;
  for (int c : {256, 128}) {
    if (n_out % c) continue;
    const size_t bb = size_t(c / 2) * 128;
    // two N tiles whose accumulators fill tensor memory exactly (n_out = 512): K block outer, ring mode as well
    // (measured at the cfg 3 mesh, B = 8: 256->512 3.42 -> 2.84 ms; 128->512 1.89 -> 2.00, hence k_in >= 256)
    const int kouter_c = (ring_env >= 2 && n_out == 2 * c && 2 * c == 512 && k_blocks >= 4) ? 1 : 0;
    const int ring_c = ((ring_env && n_out == c && k_blocks >= 2) || kouter_c) ? 1 : 0;
    // one working epilogue group (4 staging buffers) when a single resident A buffer already fills the SM
    const int eg = (ring_c || nbuf == 2) ? 2 : 1;
    const size_t a_bytes = ring_c ? size_t(kRingBlocks) * kABlock : size_t(nbuf) * size_t(k_blocks) * kABlock;
    for (int sgc : {2, 1}) {
      if (sg_env && sgc != sg_env) continue;
      if (kouter_c && sgc == 2 && !sg_env) continue;   // K-outer: the W ring (two tiles per K block) needs the room
      const size_t fx = a_bytes + size_t(2 * sgc) * kSrcStage + size_t(4 * eg * epi_bufs) * 2048 +
                        align_up(size_t(n_out) * 4, 1024) + 1024;
      if (fx + 2 * bb > cap) continue;
      const size_t fl = std::min<size_t>(kMaxSB, (cap - fx) / bb) * bb;
      if (sgc == 2 && fl < 3 * bb && fl < 32768) continue;      // the second source stage must not starve the W ring
      if (fl > best) { best = fl; bn = c; ring = ring_c; sg = sgc; epi_groups = eg; fixed = fx; kouter = kouter_c; }
      break;                                                     // the deepest source ring that fits for this bn
    }
  }
  if (!bn) return set_err(GWEN_E_NOSUPPORT, "fused layer does not fit in shared memory");
  const size_t b_bytes = size_t(bn / 2) * 128;
  const int sb = static_cast<int>(std::min<size_t>(kMaxSB, (cap - fixed) / b_bytes));
  auto enc = tensor_map_encoder();
  if (!enc) return set_err(GWEN_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap xmap, wmap, ymap;
  {
    cuuint64_t dims[4] = {cuuint64_t(k_in), cuuint64_t(w), cuuint64_t(h), cuuint64_t(batch)};
    cuuint64_t strides[3] = {cuuint64_t(k_in) * 2, cuuint64_t(w) * k_in * 2, cuuint64_t(h) * w * k_in * 2};
    cuuint32_t box[4] = {64, FT_W + 2, TH + 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GWEN_E_CUDA, "x tensor map encode failed (%d)", int(r));
  }
  {
    cuuint64_t dims[4] = {cuuint64_t(n_out), cuuint64_t(w), cuuint64_t(h), cuuint64_t(batch)};
    cuuint64_t strides[3] = {cuuint64_t(n_out) * 2, cuuint64_t(w) * n_out * 2, cuuint64_t(h) * w * n_out * 2};
    cuuint32_t box[4] = {32, FT_W, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GWEN_E_CUDA, "y tensor map encode failed (%d)", int(r));
  }
  int rc = make_tensor_map_3d(&wmap, weight, GWEN_BF16, k_in, n_out, 1, k_in, 0, BK, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  FusedArgs g{bias, dis_padded, dis_pitch, static_cast<int>(batch), static_cast<int>(h), static_cast<int>(w),
              k_blocks, static_cast<int>(n_out), bn, sb, (epilogue & GWEN_EPI_RELU) ? 1 : 0, tiles_y, pairs_x,
              nbuf, pre_bias, (pre_epilogue & GWEN_EPI_RELU) ? 1 : 0, 0, epi_groups, epi_bufs, wait_backoff_ns(), ring, sg, kouter};
  static const int prefetch_env = [] {
    const char* v = getenv("GWEN_FUSED_PREFETCH");
    // measured at the cfg 3 shapes (round 2, distance 0 / 1 / 2): 512->1024 8.65 / 9.62 / 9.70 ms, 512->256
    // 4.01 / 3.93 / 4.14, 256->512 and 64->1024 unchanged: the boxes' halos are L2 hits anyway and the extra
    // requests compete with the W tiles -- off by default
    return v ? std::max(0, std::min(8, atoi(v))) : 0;
  }();
  g.prefetch = prefetch_env;
  const size_t smem = fixed + size_t(sb) * b_bytes;
  GWEN_CUDA(cudaFuncSetAttribute(k_gcn_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t items = batch * tiles_y * pairs_x;
  const int pairs = static_cast<int>(std::min<int64_t>(items, std::max(1, (sm_count() - sm_reserve()) / 2)));
  k_gcn_fused<<<2 * pairs, kFusedThreads, smem, static_cast<cudaStream_t>(stream)>>>(xmap, wmap, ymap, g);
  GWEN_LAUNCH_CHECK("k_gcn_fused");
  return GWEN_OK;
}
