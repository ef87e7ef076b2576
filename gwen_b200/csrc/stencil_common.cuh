// Device helpers shared by the mesh-stencil aggregation (stencil.cu) and the fused
// aggregate -> projection kernel (gcn_fused.cu): packed fp32x2 arithmetic, 16-byte pack / unpack,
// shared-memory and TMA wrappers.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace gwen {
namespace st {

constexpr int TH = 8;     // destination rows per tile
constexpr int SEG = 8;    // destination columns a sub-warp walks per unit

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 lanes per instruction,
// each rounded exactly like the scalar op).
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <typename T>
struct V16;
template <>
struct V16<float> {
  static constexpr int N = 4;
  __device__ static void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
  __device__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
  // 16 bytes <-> N/2 packed fp32 pairs
  __device__ static void unpack2(const uint4& r, uint64_t* p) {
    p[0] = pk2u(r.x, r.y);
    p[1] = pk2u(r.z, r.w);
  }
  __device__ static uint4 pack2(const uint64_t* p, bool relu) {
    float f[4];
    unpk2(p[0], f[0], f[1]);
    unpk2(p[1], f[2], f[3]);
    if (relu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) f[i] = fmaxf(f[i], 0.0f);
    }
    return pack(f);
  }
};
template <>
struct V16<__nv_bfloat16> {
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
  __device__ static void unpack2(const uint4& r, uint64_t* p) {
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = pk2u(u[i] << 16, u[i] & 0xffff0000u);
  }
  __device__ static uint4 pack2(const uint64_t* p, bool relu) {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float lo, hi;
      unpk2(p[i], lo, hi);
      __nv_bfloat162 q = __floats2bfloat162_rn(lo, hi);
      u[i] = *reinterpret_cast<uint32_t*>(&q);
      if (relu) asm("max.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(0u));  // clamp after rounding == before
    }
    return make_uint4(u[0], u[1], u[2], u[3]);
  }
};

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1,
                                            int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// Plain bulk copy global -> shared (contiguous, multiple of 16 bytes), completing on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}


}  // namespace st
}  // namespace gwen
