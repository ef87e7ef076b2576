// K2 (tensor-core path) -- placeholder until the tcgen05 kernel lands: reports "not supported"
// so gwen_linear_fwd takes the CUDA-core GEMM.
#include "common.cuh"

namespace gwen {
int linear_tc_supported(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const void*,
                        const void*, const void*) {
  return 0;
}
int linear_tc_fwd_bf16(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t,
                       int64_t, int64_t, const float*, int, cudaStream_t) {
  return set_err(GWEN_E_NOSUPPORT, "tcgen05 GEMM not built");
}
}  // namespace gwen
