// TMA / mbarrier plumbing shared by the tiled aggregation and the tcgen05 GEMM:
// host-side tensor-map encoding through the driver entry point (no link against libcuda) and
// the inline-PTX wrappers for mbarrier and cp.async.bulk.tensor.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "common.cuh"

namespace gwen {

// cuTensorMapEncodeTiled resolved once through the runtime (driver symbol, versioned).
inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                         &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// Row-major [batch][rows][cols] tensor (element pitch ld, batch stride in elements) with a box of
// box_cols x box_rows x 1.  swizzle: CU_TENSOR_MAP_SWIZZLE_NONE / _128B.
inline int make_tensor_map_3d(CUtensorMap* map, const void* base, int dtype, uint64_t cols,
                              uint64_t rows, uint64_t batch, uint64_t ld, uint64_t bstride,
                              uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle) {
  auto enc = tensor_map_encoder();
  if (!enc) return set_err(GWEN_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const uint64_t esz = dtype == GWEN_F32 ? 4 : 2;
  const CUtensorMapDataType dt =
      dtype == GWEN_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[3] = {cols, rows, batch ? batch : 1};
  cuuint64_t strides[2] = {ld * esz, (batch > 1 ? bstride : ld * rows) * esz};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  static const CUtensorMapL2promotion promo = [] {   // developer knob: L2 promotion size of operand loads
    const char* v = getenv("GWEN_TMA_PROMO");
    return v && atoi(v) == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                               : (v && atoi(v) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                     : (v && atoi(v) == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
  }();
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(GWEN_E_CUDA,
                   "cuTensorMapEncodeTiled failed (%d): cols %llu rows %llu batch %llu ld %llu box "
                   "%ux%u",
                   int(r), (unsigned long long)cols, (unsigned long long)rows,
                   (unsigned long long)batch, (unsigned long long)ld, box_cols, box_rows);
  return GWEN_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Wait of a whole warp (or many warps) on a barrier that completes microseconds later: back off with nanosleep
// between polls so the pollers leave the issue slots (and the power budget) to the warps that work.  ns = 0: spin.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  while (!done) {
    if (ns) __nanosleep(ns);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
inline uint32_t wait_backoff_ns() {   // GWEN_WAIT_NS: poll interval of the bulk waiters (default 0 = spin)
  static const uint32_t v = [] {
    const char* e = getenv("GWEN_WAIT_NS");
    const int x = e ? atoi(e) : 0;
    return static_cast<uint32_t>(x < 0 ? 0 : (x > 100000 ? 100000 : x));
  }();
  return v;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// Tile load global -> shared, box origin (c0 = column, c1 = row, c2 = batch), completes on `bar`.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, int c0, int c1,
                                            int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
#endif

}  // namespace gwen
