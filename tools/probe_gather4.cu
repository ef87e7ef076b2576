// Probe: which tensor-map box does cp.async.bulk.tensor.2d.tile::gather4 (UTMALDG.2D.GATHER4) expect on sm_100a,
// where do the four rows land in shared memory, and how many such copies does one SM retire per microsecond.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/probe_gather4 tools/probe_gather4.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap m, const int* rows, int col, float* out, int* flag,
                        int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char buf[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = s32(&bar), dst = s32(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int ok = 1;
  long long t0 = clock64();
  const int n_rounds = reps > 1 ? (reps & 0xffff) : 1;
  long long t_issue = 0;
  for (int r = 0; r < n_rounds && ok; ++r) {
    if (threadIdx.x == 0) {
      const int n_ops = reps > 1 ? (reps >> 16) : 1;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024 * n_ops) : "memory");
      for (int o = 0; o < n_ops; ++o)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], "
            "[%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst + o * 1024),
            "l"(&m), "r"(col), "r"(rows[(4 * o) % 64]), "r"(rows[(4 * o + 1) % 64]), "r"(rows[(4 * o + 2) % 64]),
            "r"(rows[(4 * o + 3) % 64]), "r"(b)
            : "memory");
      if (r == n_rounds - 1) t_issue = clock64();
      uint32_t done = 0;
      for (int spin = 0; spin < 2000000 && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(r & 1) : "memory");
      ok = done;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) { *flag = ok; cycles[0] = t1 - t0; cycles[1] = t1 - t_issue; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = reinterpret_cast<float*>(buf)[i];
}

int main(int argc, char** argv) {
  const int R = 4096, C = 256;
  float* x; cudaMalloc(&x, sizeof(float) * R * C);
  float* hx = (float*)malloc(sizeof(float) * R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) hx[r * C + c] = r * 1000.0f + c;
  cudaMemcpy(x, hx, sizeof(float) * R * C, cudaMemcpyHostToDevice);
  int hrows[64]; for (int i = 0; i < 64; ++i) hrows[i] = (i * 977 + 5) % R;
  int* rows; cudaMalloc(&rows, sizeof(hrows)); cudaMemcpy(rows, hrows, sizeof(hrows), cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 1024); int* flag; cudaMalloc(&flag, 4); long long* cyc; cudaMalloc(&cyc, 16);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int boxrows = 1; boxrows <= 1; boxrows += 3) {   // box rows = 4 is an illegal instruction (measured)
    CUtensorMap m;
    cuuint64_t dims[2] = {C, R}; cuuint64_t strides[1] = {C * sizeof(float)};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows %d: encode rc=%d\n", boxrows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    cudaMemset(out, 0, 1024);
    k_probe<<<1, 32, 40 * 1024>>>(m, rows, 64, out, flag, 1, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    int hf = -1; float ho[256]; cudaMemcpy(&hf, flag, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost);
    printf("  sync=%s completed=%d  rows asked %d %d %d %d (col 64)\n", cudaGetErrorString(e), hf, hrows[0], hrows[1], hrows[2], hrows[3]);
    printf("  smem[0]=%.0f smem[63]=%.0f smem[64]=%.0f smem[128]=%.0f smem[192]=%.0f smem[255]=%.0f\n", ho[0], ho[63], ho[64], ho[128], ho[192], ho[255]);
    if (e != cudaSuccess) { printf("  device error, stopping\n"); return 1; }
    if (hf == 1) {   // throughput: 200 rounds of 32 gather4 copies (32 KB per round) from ONE SM, then from all SMs
      for (int blocks = 1; blocks <= 148; blocks += 147)
        for (int n_ops = 8; n_ops <= 32; n_ops *= 2) {
          k_probe<<<blocks, 32, 40 * 1024>>>(m, rows, 64, out, flag, (n_ops << 16) | 200, cyc);
          cudaDeviceSynchronize();
          long long hc[2]; cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(&hf, flag, 4, cudaMemcpyDeviceToHost);
          printf("  %3d CTAs, %2d copies per round: completed=%d, %.0f cycles per round, last round's wait %lld cycles\n", blocks, n_ops, hf,
                 hc[0] / 200.0, hc[1]);
        }
    }
  }
  return 0;
}
