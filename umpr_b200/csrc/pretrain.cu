// R-Net pre-training head (reference pretrain/pretrain_rnet.py:148-168): result = sigmoid(Linear(256 -> 1)([att_u | att_i])),
// loss = BCELoss(result, target) (mean; log terms clamped at -100 as torch.nn.BCELoss does).  One warp per sample.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

__global__ void __launch_bounds__(256) bce_head_fwd_kernel(const float* __restrict__ att_u, const float* __restrict__ att_i,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ target, int B, float* __restrict__ result,
                                                           float* __restrict__ loss /* zero-initialised */) {
  const int lane = threadIdx.x & 31, b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float4 u = *reinterpret_cast<const float4*>(att_u + (size_t)b * D + lane * 4);
  const float4 v = *reinterpret_cast<const float4*>(att_i + (size_t)b * D + lane * 4);
  const float4 wu = *reinterpret_cast<const float4*>(w + lane * 4);
  const float4 wv = *reinterpret_cast<const float4*>(w + D + lane * 4);
  float z = u.x * wu.x + u.y * wu.y + u.z * wu.z + u.w * wu.w + v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
  z = warp_sum(z) + bias[0];
  if (lane == 0) {
    const float p = sigmoidf_acc(z), t = target[b];
    result[b] = p;
    const float l = -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
    atomicAdd(loss, l / (float)B);
  }
}

// d_loss: scalar upstream gradient; d_result: optional (B,) upstream gradient of the returned probabilities.
// The weight / bias gradients are sums of (p - t)-signed terms that largely cancel (|sum| ~ 1e-2 of the sum of magnitudes for balanced
// targets): they are accumulated in double per lane and per block, and only a few dozen block results meet in fp32 atomics.
__global__ void __launch_bounds__(256) bce_head_bwd_kernel(const float* __restrict__ att_u, const float* __restrict__ att_i,
                                                           const float* __restrict__ w, const float* __restrict__ result,
                                                           const float* __restrict__ target, const float* __restrict__ d_loss,
                                                           const float* __restrict__ d_result, int B, float* __restrict__ d_att_u,
                                                           float* __restrict__ d_att_i, float* __restrict__ d_w, float* __restrict__ d_b) {
  __shared__ double s_dw[8][2 * D];
  __shared__ double s_db[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  const float4 wu = *reinterpret_cast<const float4*>(w + lane * 4);
  const float4 wv = *reinterpret_cast<const float4*>(w + D + lane * 4);
  const float dl = (d_loss ? d_loss[0] : 0.f) / (float)B;
  double au[4] = {0., 0., 0., 0.}, av[4] = {0., 0., 0., 0.}, ab = 0.;
  for (int b = gw; b < B; b += nw) {
    const float p = result[b], t = target[b];
    // BCELoss backward as ATen computes it: (p - t) / max((1 - p) p, 1e-12), then through the sigmoid
    float dz = dl * (p - t) / fmaxf((1.f - p) * p, 1e-12f) * p * (1.f - p);
    if (d_result) dz += d_result[b] * p * (1.f - p);
    const float4 u = *reinterpret_cast<const float4*>(att_u + (size_t)b * D + lane * 4);
    const float4 v = *reinterpret_cast<const float4*>(att_i + (size_t)b * D + lane * 4);
    *reinterpret_cast<float4*>(d_att_u + (size_t)b * D + lane * 4) = make_float4(dz * wu.x, dz * wu.y, dz * wu.z, dz * wu.w);
    *reinterpret_cast<float4*>(d_att_i + (size_t)b * D + lane * 4) = make_float4(dz * wv.x, dz * wv.y, dz * wv.z, dz * wv.w);
    au[0] += (double)dz * u.x; au[1] += (double)dz * u.y; au[2] += (double)dz * u.z; au[3] += (double)dz * u.w;
    av[0] += (double)dz * v.x; av[1] += (double)dz * v.y; av[2] += (double)dz * v.z; av[3] += (double)dz * v.w;
    ab += dz;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { s_dw[warp][lane * 4 + q] = au[q]; s_dw[warp][D + lane * 4 + q] = av[q]; }
  if (lane == 0) s_db[warp] = ab;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    double a = 0.;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += s_dw[k][i];
    atomicAdd(&d_w[i], (float)a);
  }
  if (threadIdx.x == 0) {
    double a = 0.;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += s_db[k];
    atomicAdd(d_b, (float)a);
  }
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_bce_head_fwd(const float* att_u, const float* att_i, const float* w, const float* bias, const float* target, int B,
                                 float* result, float* loss, void* stream) {
  if (B <= 0) return fail_arg("bce_head: empty batch");
  bce_head_fwd_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(att_u, att_i, w, bias, target, B, result, loss);
  return check_launch("bce_head_fwd");
}

extern "C" int umpr_bce_head_bwd(const float* att_u, const float* att_i, const float* w, const float* result, const float* target,
                                 const float* d_loss, const float* d_result, int B, float* d_att_u, float* d_att_i, float* d_w,
                                 float* d_b, void* stream) {
  if (B <= 0) return 0;
  const int blocks = (B + 7) / 8 < 32 ? (B + 7) / 8 : 32;
  bce_head_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(att_u, att_i, w, result, target, d_loss, d_result, B, d_att_u, d_att_i,
                                                                     d_w, d_b);
  return check_launch("bce_head_bwd");
}
