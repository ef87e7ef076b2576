// Adam step of the reference training loop (configs_train_gnn.optimizer.step(), src/gwen/models_gnn.py:373; the
// optimizer is torch.optim.Adam(model.parameters(), lr = config["lr"] * 10), src/gwen/train_gnn.py:111): every
// parameter tensor of the model in ONE launch (torch's foreach implementation runs ~10 multi-tensor kernels per
// step; the model has 20 parameter tensors, 12 of them live).  fp32 parameters, gradients and moments; the
// arithmetic follows torch.optim.Adam (amsgrad = False, maximize = False):
//   g' = g + wd * p;  m += (g' - m) (1 - b1);  v = b2 v + (1 - b2) g'^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace gwen {
namespace {

constexpr int kAdamMaxTensors = 32;
constexpr int kAdamThreads = 256;

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  int64_t n[kAdamMaxTensors];
  int64_t blk0[kAdamMaxTensors + 1];   // first block of every tensor (exclusive scan of ceil(n / chunk))
  int count;
};

constexpr int64_t kAdamChunk = int64_t(kAdamThreads) * 8;   // elements per block

__global__ void __launch_bounds__(kAdamThreads) k_adam(const __grid_constant__ AdamTable t, float step_size,
                                                       float omb1, float b2, float omb2, float inv_bc2_sqrt,
                                                       float eps, float wd) {
  // which tensor does this block belong to? (at most 32 entries: linear search)
  int ti = 0;
  while (ti + 1 < t.count && int64_t(blockIdx.x) >= t.blk0[ti + 1]) ++ti;
  const int64_t base = (int64_t(blockIdx.x) - t.blk0[ti]) * kAdamChunk;
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const int64_t n = t.n[ti];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t i = base + int64_t(k) * kAdamThreads + threadIdx.x;
    if (i >= n) break;
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(gi - m[i], omb1, m[i]);          // lerp_(g, 1 - b1)
    const float vi = fmaf(omb2, gi * gi, b2 * v[i]);      // mul_(b2).addcmul_(g, g, value = 1 - b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_adam_step(int32_t count, void* const* params, const void* const* grads, void* const* exp_avg,
                              void* const* exp_avg_sq, const int64_t* numel, int64_t step, double lr, double beta1,
                              double beta2, double eps, double weight_decay, void* stream) {
  GWEN_CHECK_ARG(count >= 0 && step >= 1, "count must be >= 0 and step >= 1");
  if (count == 0) return GWEN_OK;
  GWEN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel, "null table");
  // hyper-parameters arrive as doubles (Python floats): 1 - beta and beta^step are formed in fp64 like torch does
  const double bc1 = 1.0 - pow(beta1, double(step));
  const double bc2 = 1.0 - pow(beta2, double(step));
  const float step_size = static_cast<float>(double(lr) / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  for (int32_t first = 0; first < count; first += kAdamMaxTensors) {
    AdamTable t{};
    t.count = std::min<int32_t>(kAdamMaxTensors, count - first);
    int64_t blocks = 0;
    for (int i = 0; i < t.count; ++i) {
      GWEN_CHECK_ARG(numel[first + i] >= 0, "negative numel");
      GWEN_CHECK_ARG(numel[first + i] == 0 || (params[first + i] && grads[first + i] && exp_avg[first + i] &&
                                               exp_avg_sq[first + i]), "null tensor pointer");
      t.p[i] = static_cast<float*>(params[first + i]);
      t.g[i] = static_cast<const float*>(grads[first + i]);
      t.m[i] = static_cast<float*>(exp_avg[first + i]);
      t.v[i] = static_cast<float*>(exp_avg_sq[first + i]);
      t.n[i] = numel[first + i];
      t.blk0[i] = blocks;
      blocks += ceil_div(numel[first + i], kAdamChunk);
    }
    t.blk0[t.count] = blocks;
    if (blocks == 0) continue;
    GWEN_CHECK_ARG(blocks < INT32_MAX, "too many elements for one launch");
    k_adam<<<static_cast<unsigned>(blocks), kAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        t, step_size, static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
        inv_bc2_sqrt, static_cast<float>(eps), static_cast<float>(weight_decay));
    GWEN_LAUNCH_CHECK("k_adam");
  }
  return GWEN_OK;
}
