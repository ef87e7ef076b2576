// Two projections back to back in one kernel (bf16):  Y = epi2( act(X W1^T + b1) W2^T + b2 ).
//
// Where a layer that projects LAST is followed by a layer that projects FIRST, the wide hidden tensor between
// them is written once and read once for nothing: in the GWEN stack conv1 (64 -> 1024, bias + ReLU) feeds the
// projection of conv2 (1024 -> 512), and upconv4 (512 -> 1024, bias + ReLU) feeds the projection of upconv5
// (1024 -> 64): 14.7 GB written + 14.7 GB read per forward at the cfg 3 size, each.  Here the hidden rows never
// leave the SM: a CTA pair owns 256 rows; per 256-column chunk of the hidden layer the first product lands in a
// tensor-memory accumulator, sixteen epilogue warps turn it into bf16 (bias, ReLU) and write it to shared memory in
// the K-major SWIZZLE_128B layout of a tcgen05 A operand, and the second product consumes it as four 64-deep K
// steps into a second accumulator that stays resident for the whole row block.
//
// Roles per CTA (20 warps, cluster of 2, every MMA is cta_group::2 with M = 256):
//   warp 0 lane 0   X producer: the 128 x K1 row block as K1/64 resident 16 KB blocks (two buffers for K1 <= 128, else
//                   one whose blocks are released behind the LAST chunk's MMAs, so the next row block streams in
//                   behind them)
//   warp 1 lane 0   MMA issuer (leader): software pipeline  M1(c) , M2(c-1) , M1(c+1) , M2(c) ...  over ALL chunks of
//                   all row blocks: the hidden accumulator is drained under M2(c-1), converted and stored under M1(c+1)
//   warp 2 / 3      W1 / W2 tile producers (each CTA its half of the N rows, as linear_tc3.cu)
//   warps 4..19     epilogue: group g = 64-column block of the chunk, q = TMEM lane quarter.  Hidden chunk: both TMEM
//                   loads and the conversion come first (the accumulator is free again after ~2 loads), then the wait
//                   for the shared-memory block, then eight 16-byte stores.  The same warps drain the OUTPUT accumulator
//                   of a finished pass between two hidden chunks (32-column sub-chunks, 64-byte runs stored straight
//                   from registers: the output is the narrow side, ~1 TB/s at most) -- with four dedicated warps and
//                   TMA staging the drain of the single output accumulator stalled the issuer for 25 % of its time
// N2 > 256 runs as N2/256 passes over the row block (the first product is recomputed per pass: tensor memory holds
// 256 hidden columns + 256 output columns = all 512).
// Why 256-column chunks and one hidden accumulator: the single issuing thread needs ~50 cycles per tcgen05.mma and ~100
// per commit; with N = 128 instructions (first version: two 128-column accumulators) it, not the tensor pipe, was the
// bound (issuer 63 % busy issuing, 11 % waiting, tensor pipe 32 % active).  A hidden block and the W2 tile that multiplies
// it share one slot index and ONE release barrier (one commit per K step of the second product).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "tcgen05.cuh"

namespace gwen {
using namespace tc;
namespace {

constexpr int kEpi1Warps = 16;
constexpr int kB2bThreads = (4 + kEpi1Warps) * 32;
constexpr int kBn1 = 256;                 // hidden columns per chunk
constexpr int kCb = kBn1 / BK;            // 64-column hidden blocks per chunk
constexpr int kMaxA = 8;                  // resident X blocks (K1 <= 512)
constexpr int kMaxS1 = 6;                 // W1 ring
constexpr int kMaxH = 4;                  // hidden blocks of 16 KB (64 columns) = W2 stages
constexpr uint32_t kBlk = BM * 128u;      // 128 rows x 64 bf16
constexpr uint32_t kW1Stage = (kBn1 / 2) * 128u;

struct B2bArgs {
  const float* bias1;
  const float* bias2;
  int64_t m;
  int k1_blocks, n1, n2, bn2, relu1, relu2;
  int nbuf_a, s1, hb;
  void* y;
  int64_t ldy;
  uint32_t wait_ns;
  unsigned long long* prof;   // GWEN_B2B_PROF=1: per pair, the issuer's cycles in each wait / issue section (developer tool)
};

// Ring bookkeeping without integer division: the single-thread roles are latency-bound (a division by a run-time
// ring depth costs ~100 cycles; a dozen of them per chunk was a third of the issuer's time in the first version).
struct Ring {
  uint32_t i, ph, n;
  __device__ __forceinline__ void next() { if (++i == n) { i = 0; ph ^= 1u; } }
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kB2bThreads, 1)
    k_linear_b2b(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap w1map,
                 const __grid_constant__ CUtensorMap w2map, B2bArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxA], a_empty[kMaxA], w1_full[kMaxS1], w1_empty[kMaxS1],
      w2_full[kMaxH], h_full[kMaxH], blk_empty[kMaxH], d1_full, d1_empty, d2_full, d2_empty;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbs = g.k1_blocks;
  const uint32_t w2_stage = uint32_t(g.bn2 / 2) * 128u;
  // shared-memory carve-up
  const uint32_t a_base = base;
  const uint32_t h_base = a_base + uint32_t(g.nbuf_a * kbs) * kBlk;
  const uint32_t w1_base = h_base + uint32_t(g.hb) * kBlk;
  const uint32_t w2_base = w1_base + uint32_t(g.s1) * kW1Stage;
  const uint32_t bias_base = w2_base + uint32_t(g.hb) * w2_stage;
  float* bias1_s = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));
  float* bias2_s = bias1_s + g.n1;

  const int chunks = g.n1 / kBn1, passes = g.n2 / g.bn2, n_sub2 = g.bn2 / 32;
  const int64_t items = (g.m + 2 * BM - 1) / (2 * BM);
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t my_items = items > pair_id ? (items - 1 - pair_id) / n_pairs + 1 : 0;
  const uint32_t tmem_cols = 512u;
  const uint32_t d2_col0 = uint32_t(kBn1);
  const bool a2 = g.nbuf_a == 2;
  const uint32_t hmask = uint32_t(g.hb - 1), hshift = g.hb == 4 ? 2u : 1u;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&w1map);
    tma_prefetch_desc(&w2map);
    for (int i = 0; i < kMaxA; ++i) {
      mbar_init(smem_u32(&a_full[i]), 1);
      mbar_init(smem_u32(&a_empty[i]), 1);
    }
    for (int i = 0; i < kMaxS1; ++i) {
      mbar_init(smem_u32(&w1_full[i]), 1);
      mbar_init(smem_u32(&w1_empty[i]), 1);
    }
    for (int i = 0; i < kMaxH; ++i) {
      mbar_init(smem_u32(&w2_full[i]), 1);
      mbar_init(smem_u32(&h_full[i]), 8);             // 4 hidden-epilogue warps x 2 CTAs (the leader's is used)
      mbar_init(smem_u32(&blk_empty[i]), 1);
    }
    mbar_init(smem_u32(&d1_full), 1);
    mbar_init(smem_u32(&d1_empty), 2u * kEpi1Warps);
    mbar_init(smem_u32(&d2_full), 1);
    mbar_init(smem_u32(&d2_empty), 2u * 4u * uint32_t(n_sub2 < 4 ? n_sub2 : 4));
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.n1; i += kB2bThreads) bias1_s[i] = g.bias1 ? __ldg(g.bias1 + i) : 0.0f;
  for (int i = threadIdx.x; i < g.n2; i += kB2bThreads) bias2_s[i] = g.bias2 ? __ldg(g.bias2 + i) : 0.0f;
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
    if (lane == 0) {  // ===== X producer (both CTAs) =====
      for (int64_t seq = 0; seq < my_items; ++seq) {
        const int64_t it = pair_id + seq * n_pairs;
        const int m0 = int(it * (2 * BM)) + int(rank) * BM;
        const uint32_t abuf = a2 ? uint32_t(seq) & 1u : 0u, use = a2 ? uint32_t(seq >> 1) : uint32_t(seq);
        for (int kb = 0; kb < kbs; ++kb) {
          const uint32_t slot = abuf * uint32_t(kbs) + uint32_t(kb);
          if (use > 0) mbar_wait(smem_u32(&a_empty[slot]), (use - 1) & 1u);
          if (leader) mbar_expect_tx(smem_u32(&a_full[slot]), 2 * kBlk);
          tma_load_3d_pair(a_base + slot * kBlk, &xmap, kb * BK, m0, 0, mapa_u32(smem_u32(&a_full[slot]), 0));
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc1 = make_idesc_pair(kBn1), idesc2 = make_idesc_pair(g.bn2);
      Ring r1{0u, 0u, uint32_t(g.s1)};
      // waits: d1_empty, a_full, w1_full, h_full, w2_full, d2_empty; total; chunks; M1 issue, M1 commit, M2 issue, M2 commit
      long long pw[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      const bool prof = g.prof != nullptr;
      const long long t_begin = prof ? clock64() : 0;
      auto twait = [&](int which, uint32_t bar, uint32_t parity) {
        if (!prof) { mbar_wait(bar, parity); return; }
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        pw[which] += clock64() - t0;
      };
      const uint64_t adesc0 = make_smem_desc(a_base), hdesc0 = make_smem_desc(h_base);
      const uint64_t w1desc0 = make_smem_desc(w1_base), w2desc0 = make_smem_desc(w2_base);
      const uint32_t d2_addr = tmem_d + d2_col0;
      // second product of chunk c2 (gp2 = global pass of that chunk, j2 = its index in the pass)
      auto issue_m2 = [&](uint32_t c2, uint32_t gp2, int j2) {
        if (j2 == 0 && gp2 > 0) twait(5, smem_u32(&d2_empty), (gp2 - 1) & 1u);
#pragma unroll
        for (int blk = 0; blk < kCb; ++blk) {
          const uint32_t hblk = c2 * uint32_t(kCb) + uint32_t(blk), slot = hblk & hmask, ph = (hblk >> hshift) & 1u;
          twait(3, smem_u32(&h_full[slot]), ph);
          twait(4, smem_u32(&w2_full[slot]), ph);
          tc_fence_after();
          const uint64_t adesc = hdesc0 + uint64_t((slot * kBlk) >> 4);
          const uint64_t bdesc = w2desc0 + uint64_t((slot * w2_stage) >> 4);
          const long long tm0 = prof ? clock64() : 0;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            umma_f16_pair(d2_addr, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc2, (j2 | blk | kk) ? 1u : 0u);
          const long long tm1 = prof ? clock64() : 0;
          umma_commit_pair(smem_u32(&blk_empty[blk]));      // frees the hidden block AND the W2 stage (one barrier per
                                                            // block index: exactly one completion per chunk each)
          if (prof) { pw[10] += tm1 - tm0; pw[11] += clock64() - tm1; }
        }
        if (j2 == chunks - 1) umma_commit_pair(smem_u32(&d2_full));
      };
      uint32_t c = 0, gp = 0;
      for (int64_t seq = 0; seq < my_items; ++seq) {
        const uint32_t abuf = a2 ? uint32_t(seq) & 1u : 0u, aphase = (a2 ? uint32_t(seq >> 1) : uint32_t(seq)) & 1u;
        for (int pass = 0; pass < passes; ++pass, ++gp) {
          for (int j = 0; j < chunks; ++j, ++c) {
            // first product of chunk c: the accumulator must have been read by the hidden epilogue of chunk c - 1
            if (c > 0) twait(0, smem_u32(&d1_empty), (c - 1) & 1u);
            tc_fence_after();
            const bool first = pass == 0 && j == 0, last = pass == passes - 1 && j == chunks - 1;
            for (int kb = 0; kb < kbs; ++kb) {
              const uint32_t aslot = abuf * uint32_t(kbs) + uint32_t(kb);
              if (first) twait(1, smem_u32(&a_full[aslot]), aphase);
              twait(2, smem_u32(&w1_full[r1.i]), r1.ph);
              tc_fence_after();
              const uint64_t adesc = adesc0 + uint64_t((aslot * kBlk) >> 4);
              const uint64_t bdesc = w1desc0 + uint64_t((r1.i * kW1Stage) >> 4);
              const long long tm0 = prof ? clock64() : 0;
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk)
                umma_f16_pair(tmem_d, adesc + uint64_t(kk * 2), bdesc + uint64_t(kk * 2), idesc1, (kb | kk) ? 1u : 0u);
              const long long tm1 = prof ? clock64() : 0;
              umma_commit_pair(smem_u32(&w1_empty[r1.i]));
              if (last) umma_commit_pair(smem_u32(&a_empty[aslot]));
              if (prof) { pw[8] += tm1 - tm0; pw[9] += clock64() - tm1; }
              r1.next();
            }
            umma_commit_pair(smem_u32(&d1_full));
            // second product of the chunk before it (its hidden blocks are converted under the MMAs just issued)
            if (c > 0) {
              if (j > 0) issue_m2(c - 1, gp, j - 1);
              else issue_m2(c - 1, gp - 1, chunks - 1);
            }
          }
        }
      }
      if (c > 0) issue_m2(c - 1, gp - 1, chunks - 1);
      if (prof) {
        pw[6] = clock64() - t_begin;
        pw[7] = c;
        for (int i = 0; i < 12; ++i) g.prof[pair_id * 16 + i] = static_cast<unsigned long long>(pw[i]);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {  // ===== W1 producer (both CTAs): the order of the issuer's first products =====
      Ring r{0u, 0u, uint32_t(g.s1)};
      bool wrapped = false;
      const int64_t my_passes = my_items * passes;
      for (int64_t gp = 0; gp < my_passes; ++gp)
        for (int j = 0; j < chunks; ++j) {
          const int n0 = j * kBn1 + int(rank) * (kBn1 / 2);
          for (int kb = 0; kb < kbs; ++kb) {
            if (wrapped) mbar_wait(smem_u32(&w1_empty[r.i]), r.ph ^ 1u);
            if (leader) mbar_expect_tx(smem_u32(&w1_full[r.i]), 2 * kW1Stage);
            tma_load_3d_pair(w1_base + r.i * kW1Stage, &w1map, kb * BK, n0, 0, mapa_u32(smem_u32(&w1_full[r.i]), 0));
            r.next();
            if (r.i == 0) wrapped = true;
          }
        }
    }
  } else if (warp == 3) {
    if (lane == 0) {  // ===== W2 producer (both CTAs): the order of the issuer's second products =====
      uint32_t c = 0;
      for (int64_t seq = 0; seq < my_items; ++seq)
        for (int pass = 0; pass < passes; ++pass) {
          const int n0 = pass * g.bn2 + int(rank) * (g.bn2 / 2);
          for (int j = 0; j < chunks; ++j, ++c)
            for (int blk = 0; blk < kCb; ++blk) {
              // the slot's previous occupant: four slots -- the same block of the chunk before; two slots -- block
              // blk + 2 of the chunk before (blk < 2) or block blk - 2 of this chunk
              const uint32_t slot = uint32_t(blk) & hmask;
              const bool same_chunk = g.hb == 2 && blk >= 2;
              const uint32_t prev_blk = g.hb == 4 ? uint32_t(blk) : uint32_t(blk ^ 2);
              if (same_chunk) mbar_wait(smem_u32(&blk_empty[prev_blk]), c & 1u);
              else if (c > 0) mbar_wait(smem_u32(&blk_empty[prev_blk]), (c - 1) & 1u);
              if (leader) mbar_expect_tx(smem_u32(&w2_full[slot]), 2 * w2_stage);
              tma_load_3d_pair(w2_base + slot * w2_stage, &w2map, j * kBn1 + blk * BK, n0, 0,
                               mapa_u32(smem_u32(&w2_full[slot]), 0));
            }
        }
    }
  } else {
    // ===== epilogue warps 4..19 =====
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const bool relu1 = g.relu1 != 0, relu2 = g.relu2 != 0;
    const uint32_t row = uint32_t(q * 32 + lane);
    const uint32_t sw = row & 7u;
    const uint32_t d1_empty_remote = mapa_u32(smem_u32(&d1_empty), 0);
    const uint32_t d2_empty_remote = mapa_u32(smem_u32(&d2_empty), 0);
    const uint32_t t_addr = tmem_d + uint32_t(grp * 64) + (uint32_t(q * 32) << 16);
    const uint32_t t2_addr = tmem_d + d2_col0 + (uint32_t(q * 32) << 16);
    const uint32_t slot = uint32_t(grp) & hmask;      // the hidden slot this warp writes (fixed: 4 c + grp mod hb)
    __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(g.y);
    // output accumulator of global pass gp2 (row block seq2, pass p2) -> bias2, ReLU, bf16 -> global memory
    auto drain_d2 = [&](uint32_t gp2, int64_t seq2, int p2) {
      if (grp >= n_sub2) return;
      const int64_t it = pair_id + seq2 * n_pairs;
      const int64_t grow = it * (2 * BM) + int64_t(rank) * BM + row;
      mbar_wait_backoff(smem_u32(&d2_full), gp2 & 1u, g.wait_ns);
      tc_fence_after();
      for (int sc = grp; sc < n_sub2; sc += 4) {
        const int cc = sc * 32;
        uint32_t r[32];
        tmem_ld32_nowait(t2_addr + uint32_t(cc), r);
        const float4* bp = reinterpret_cast<const float4*>(bias2_s + p2 * g.bn2 + cc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (sc + 4 >= n_sub2) {  // last TMEM read of this pass by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(d2_empty_remote);
        }
        uint4* dstg = reinterpret_cast<uint4*>(yb + grow * g.ldy + p2 * g.bn2 + cc);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 b0 = bp[2 * jj], b1 = bp[2 * jj + 1];
          uint4 o;
          o.x = bias_pack(r[8 * jj + 0], r[8 * jj + 1], b0.x, b0.y, relu2);
          o.y = bias_pack(r[8 * jj + 2], r[8 * jj + 3], b0.z, b0.w, relu2);
          o.z = bias_pack(r[8 * jj + 4], r[8 * jj + 5], b1.x, b1.y, relu2);
          o.w = bias_pack(r[8 * jj + 6], r[8 * jj + 7], b1.z, b1.w, relu2);
          if (grow < g.m) dstg[jj] = o;
        }
      }
    };
    // the previous occupant of this warp's hidden slot (see the W2 producer)
    const bool same_chunk = g.hb == 2 && grp >= 2;
    const uint32_t prev_blk = g.hb == 4 ? uint32_t(grp) : uint32_t(grp ^ 2);
    const bool drain_first = g.hb == 2;
    uint32_t c = 0, gp = 0;
    for (int64_t seq = 0; seq < my_items; ++seq)
      for (int pass = 0; pass < passes; ++pass, ++gp)
        for (int j = 0; j < chunks; ++j, ++c) {
          mbar_wait_backoff(smem_u32(&d1_full), c & 1u, g.wait_ns);
          tc_fence_after();
          const float* bsrc = bias1_s + j * kBn1 + grp * 64;
          uint32_t r[32];
          uint4 o[8];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld32_nowait(t_addr + uint32_t(32 * half), r);
            const float4* bp = reinterpret_cast<const float4*>(bsrc + 32 * half);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float4 b0 = bp[2 * jj], b1 = bp[2 * jj + 1];
              o[4 * half + jj].x = bias_pack(r[8 * jj + 0], r[8 * jj + 1], b0.x, b0.y, relu1);
              o[4 * half + jj].y = bias_pack(r[8 * jj + 2], r[8 * jj + 3], b0.z, b0.w, relu1);
              o[4 * half + jj].z = bias_pack(r[8 * jj + 4], r[8 * jj + 5], b1.x, b1.y, relu1);
              o[4 * half + jj].w = bias_pack(r[8 * jj + 6], r[8 * jj + 7], b1.z, b1.w, relu1);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(d1_empty_remote);   // the accumulator is read: the next chunk's MMAs may start
          // The tensor pipe runs  M1(c) , M2(c - 1) : the pass before this one finishes right behind this chunk's first
          // product, so its output accumulator is drained at this chunk.  With two hidden slots the drain must come
          // BEFORE the wait for the slot (that wait can depend on M2(c), which the issuer starts only once the drain has
          // freed the output accumulator); with four slots the hidden block goes first (it is what M2(c) waits for).
          const bool drain_now = j == 0 && gp > 0;
          if (drain_now && drain_first) {
            if (pass > 0) drain_d2(gp - 1, seq, pass - 1);
            else drain_d2(gp - 1, seq - 1, passes - 1);
          }
          // the block about to be overwritten must have been read by the second product that used it last
          if (same_chunk) mbar_wait_backoff(smem_u32(&blk_empty[prev_blk]), c & 1u, g.wait_ns);
          else if (c > 0) mbar_wait_backoff(smem_u32(&blk_empty[prev_blk]), (c - 1u) & 1u, g.wait_ns);
          const uint32_t dst = h_base + slot * kBlk + row * 128u;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) sts_v4(dst + ((uint32_t(jj) ^ sw) << 4), o[jj]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> tensor-core reads
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&h_full[slot]), 0));
          if (drain_now && !drain_first) {
            if (pass > 0) drain_d2(gp - 1, seq, pass - 1);
            else drain_d2(gp - 1, seq - 1, passes - 1);
          }
        }
    if (gp > 0) drain_d2(gp - 1, my_items - 1, passes - 1);
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

extern "C" int gwen_linear_b2b_supported(int64_t m, int64_t k1, int64_t n1, int64_t n2, int dtype) {
  if (dtype != GWEN_BF16 || sm_count() % 2) return 0;
  if (m < 256 || m * std::max<int64_t>(std::max(k1, n1), n2) >= (int64_t(1) << 40) || m >= INT32_MAX - 512) return 0;
  if (k1 < 64 || k1 % 64 || k1 > 64 * kMaxA) return 0;
  if (n1 < kBn1 || n1 % kBn1 || n1 > 8192) return 0;
  if (n2 < 64 || n2 % 64 || (n2 > 256 && n2 % 256) || n2 > 2048) return 0;
  return 1;
}

extern "C" int gwen_linear_b2b_fwd(const void* x, const void* w1, const float* bias1, int epilogue1, const void* w2,
                                   const float* bias2, int epilogue2, void* y, int64_t m, int64_t k1, int64_t n1,
                                   int64_t n2, int64_t ldx, int64_t ldy, int dtype, void* stream) {
  GWEN_CHECK_ARG(m >= 0 && k1 >= 0 && n1 >= 0 && n2 >= 0, "negative size");
  if (m == 0 || n2 == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && w1 && w2 && y, "null pointer");
  GWEN_CHECK_ARG(ldx >= k1 && ldy >= n2, "row pitch smaller than the row");
  if (!gwen_linear_b2b_supported(m, k1, n1, n2, dtype))
    return set_err(GWEN_E_NOSUPPORT,
                   "back-to-back projection needs bf16, m >= 256, k1 in {64, .., 512}, n1 %% 256 == 0, n2 %% 64 == 0 "
                   "(n2 > 256: n2 %% 256 == 0)");
  if (!aligned16(x) || !aligned16(w1) || !aligned16(w2) || !aligned16(y) || ldx % 8 || ldy % 8)
    return set_err(GWEN_E_ALIGN, "back-to-back projection needs 16-byte aligned rows");
  const int k1_blocks = static_cast<int>(k1 / 64);
  const int bn2 = static_cast<int>(n2 <= 256 ? n2 : 256);
  const int nbuf_a = 2 * k1_blocks <= 4 ? 2 : 1;
  const size_t w2_stage = size_t(bn2 / 2) * 128;
  const size_t cap = 226 * 1024;
  const size_t fixed = size_t(nbuf_a * k1_blocks) * kBlk + align_up(size_t(n1 + n2) * 4, 1024) + 1024;
  // hidden blocks (= W2 stages): a whole chunk's four when they fit next to two W1 stages, else two
  static const int hb_env = [] { const char* v = getenv("GWEN_B2B_HB"); return v ? atoi(v) : 0; }();
  static const int s1_env = [] { const char* v = getenv("GWEN_B2B_S1"); return v ? atoi(v) : 0; }();
  int hb = 0;
  for (int c : {4, 2}) {
    if ((hb_env == 2 || hb_env == 4) && c != hb_env) continue;
    if (fixed + size_t(c) * (kBlk + w2_stage) + 2 * kW1Stage <= cap) { hb = c; break; }
  }
  if (!hb) return set_err(GWEN_E_NOSUPPORT, "back-to-back projection does not fit in shared memory");
  size_t used = fixed + size_t(hb) * (kBlk + w2_stage);
  int s1 = static_cast<int>(std::min<size_t>(kMaxS1, (cap - used) / kW1Stage));
  if (s1_env >= 2 && s1_env <= s1) s1 = s1_env;
  used += size_t(s1) * kW1Stage;
  CUtensorMap xmap, w1map, w2map;
  int rc = make_tensor_map_3d(&xmap, x, GWEN_BF16, k1, m, 1, ldx, 0, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&w1map, w1, GWEN_BF16, k1, n1, 1, k1, 0, BK, kBn1 / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  rc = make_tensor_map_3d(&w2map, w2, GWEN_BF16, n1, n2, 1, n1, 0, BK, bn2 / 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != GWEN_OK) return rc;
  B2bArgs g{bias1, bias2, m, k1_blocks, static_cast<int>(n1), static_cast<int>(n2), bn2,
            (epilogue1 & GWEN_EPI_RELU) ? 1 : 0, (epilogue2 & GWEN_EPI_RELU) ? 1 : 0, nbuf_a, s1, hb, y, ldy,
            wait_backoff_ns(), nullptr};
  static const bool prof_env = getenv("GWEN_B2B_PROF") != nullptr;
  static unsigned long long* prof_buf = nullptr;
  if (prof_env) {
    if (!prof_buf) GWEN_CUDA(cudaMalloc(&prof_buf, 128 * 16 * sizeof(unsigned long long)));
    GWEN_CUDA(cudaMemsetAsync(prof_buf, 0, 128 * 16 * sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
    g.prof = prof_buf;
  }
  const size_t smem = std::max<size_t>(used, 120 * 1024);   // one CTA per SM: the pair owns all 512 TMEM columns
  GWEN_CUDA(cudaFuncSetAttribute(k_linear_b2b, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t items = ceil_div(m, 2 * BM);
  const int pairs = static_cast<int>(std::min<int64_t>(items, std::max(1, (sm_count() - sm_reserve()) / 2)));
  k_linear_b2b<<<2 * pairs, kB2bThreads, smem, static_cast<cudaStream_t>(stream)>>>(xmap, w1map, w2map, g);
  GWEN_LAUNCH_CHECK("k_linear_b2b");
  if (prof_env) {   // developer tool: where the MMA issuer of each pair spent its cycles (averaged over the pairs)
    unsigned long long hbuf[128 * 16];
    GWEN_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    GWEN_CUDA(cudaMemcpy(hbuf, prof_buf, sizeof(hbuf), cudaMemcpyDeviceToHost));
    double acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < pairs && p < 128; ++p)
      for (int i = 0; i < 12; ++i) acc[i] += double(hbuf[p * 16 + i]) / std::min(pairs, 128);
    fprintf(stderr,
            "[b2b prof] k1 %lld n1 %lld n2 %lld hb %d s1 %d: issuer cycles total %.0f (%.0f per chunk); waits d1_empty %.1f%% "
            "a_full %.1f%% w1_full %.1f%% h_full %.1f%% w2_full %.1f%% d2_empty %.1f%%; M1 issue %.1f%% commit %.1f%%, M2 issue "
            "%.1f%% commit %.1f%%\n",
            (long long)k1, (long long)n1, (long long)n2, hb, s1, acc[6], acc[6] / std::max(1.0, acc[7]),
            100 * acc[0] / acc[6], 100 * acc[1] / acc[6], 100 * acc[2] / acc[6], 100 * acc[3] / acc[6],
            100 * acc[4] / acc[6], 100 * acc[5] / acc[6], 100 * acc[8] / acc[6], 100 * acc[9] / acc[6],
            100 * acc[10] / acc[6], 100 * acc[11] / acc[6]);
  }
  return GWEN_OK;
}
