// Fused Adam(+L2) over one flat fp32 parameter buffer (reference main.py:22-26,37: torch.optim.Adam, weight_decay on
// every tensor whose name lacks "bias").  `weight_decay` is a per-element array so bias elements simply carry 0.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ wd, long n, float step_size,
                                                   float beta1, float beta2, float eps, float inv_bc2_sqrt, float grad_scale,
                                                   const float* __restrict__ shard_count) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  // DataParallel averages over the replicas that actually received a chunk (main.py:34 loss.mean()): when the all-reduced bucket
  // carries that count (a short last batch yields fewer chunks than ranks) it replaces the static 1/world scale
  if (shard_count) grad_scale = 1.f / fmaxf(shard_count[0], 1.f);
  const float pi = p[i];
  const float gi = g[i] * grad_scale + wd[i] * pi;
  const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
  const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
}
}  // namespace umpr

extern "C" int umpr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* weight_decay, long n,
                              float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                              const float* shard_count, void* stream) {
  if (n <= 0) return 0;
  if (step < 1) return umpr::fail_arg("adam: step=%d must be >= 1", step);
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  umpr::adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, weight_decay, n,
                                                                                   (float)(lr / bc1), beta1, beta2, eps,
                                                                                   (float)(1.0 / sqrt(bc2)), grad_scale, shard_count);
  return umpr::check_launch("adam_step");
}
