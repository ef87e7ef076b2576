// tcgen05 / TMEM inline-PTX wrappers shared by the tensor-core projection kernels
// (linear_tc.cu: cta_group::1, linear_tc3.cu: cta_group::2 CTA pairs).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace gwen {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;        // bf16 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kTcThreads = 192;

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 B
// apart (SBO), LBO unused (=1) for swizzled K-major, descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);         // start address, bits [0,14)
  d |= uint64_t(1) << 16;                         // leading byte offset (ignored), bits [16,30)
  d |= uint64_t(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
  d |= uint64_t(1) << 46;                         // version = 1
  d |= uint64_t(2) << 61;                         // layout type: SWIZZLE_128B
  return d;
}
// MN-major operand tile as TMA writes it with SWIZZLE_128B boxes of {64 elements (MN, contiguous
// in global memory), R rows (K)}: K row r at byte 128 r (chunks XOR-swizzled by r % 8), 8-row K
// groups 1024 B apart (SBO), successive 64-element MN chunks `lbo_bytes` apart (LBO) -- the
// canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) layout in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, M = 128, N = bn.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
      "r"(src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}


// ---- CTA pairs (cta_group::2) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope): what must be ordered before the arrival is this
  // warp's TMEM reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already order;
  // a .release.cluster here costs a MEMBAR.ALL.GPU per chunk (28 % of all stall samples in ncu).
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion bytes are counted on an mbarrier of either CTA of the pair.
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, int c0, int c1,
                                                 int c2, uint32_t cluster_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(cluster_bar)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs retire) on the barrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(uint16_t(3))
      : "memory");
}
// Instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, M = 256 (pair), N = bn.
__host__ __device__ constexpr uint32_t make_idesc_pair(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(bn >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// {lo, hi} fp32 pair + {blo, bhi} (one FADD2), rounded to a bf16x2, optional ReLU on the pair.
__device__ __forceinline__ uint32_t bias_pack(uint32_t lo, uint32_t hi, float blo, float bhi,
                                              bool relu) {
  uint64_t a, b, s;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(lo), "r"(hi));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(blo), "f"(bhi));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(a), "l"(b));
  float f0, f1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(s));
  __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
  uint32_t u = *reinterpret_cast<uint32_t*>(&p);
  if (relu) asm("max.bf16x2 %0, %0, %1;" : "+r"(u) : "r"(0u));  // round then clamp == clamp then round
  return u;
}


}  // namespace tc
}  // namespace gwen
