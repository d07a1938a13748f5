// Text matching (reference src/model.py:166-168): y = tanh(linear_u([atte_u | senti_u]) + linear_i([atte_i | senti_i])), both
// linears bias-free 256 -> 128, and its input gradients.  2 x 128 x 256 MACs per sample: far too small for a GEMM launch each
// (eight weight-stationary GEMM launches of ~18 us before); one fused fp32 kernel each way, 8 samples per CTA, the two weight
// matrices (256 KB) stream from L2.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

constexpr int TM_S = 8;           // samples per CTA

// warp w computes outputs n = w, w+8, ...; lanes split the 512-long reduction (coalesced weight rows), then a warp sum per sample
__global__ void __launch_bounds__(256) text_match_fwd_kernel(const float* __restrict__ a_u, const float* __restrict__ s_u,
                                                             const float* __restrict__ a_i, const float* __restrict__ s_i,
                                                             const float* __restrict__ Wu, const float* __restrict__ Wi, int B,
                                                             float* __restrict__ y) {
  __shared__ float xs[TM_S][4 * D];            // [sample][a_u | s_u | a_i | s_i]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b0 = blockIdx.x * TM_S;
  for (int idx = tid; idx < TM_S * 4 * D; idx += 256) {
    const int s = idx / (4 * D), k = idx - s * 4 * D, j = k >> 7, c = k & 127;
    const float* src = j == 0 ? a_u : j == 1 ? s_u : j == 2 ? a_i : s_i;
    xs[s][k] = b0 + s < B ? src[(size_t)(b0 + s) * D + c] : 0.f;
  }
  __syncthreads();
  // the weight row slice of this lane (16 independent loads), fetched one output ahead: the L2 round trip of output n+8 runs under the
  // 128 FMAs of output n
  float wn[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wn[j] = Wu[(size_t)warp * 2 * D + lane + 32 * j]; wn[8 + j] = Wi[(size_t)warp * 2 * D + lane + 32 * j]; }
  for (int n = warp; n < D; n += 8) {
    float acc[TM_S];
#pragma unroll
    for (int s = 0; s < TM_S; ++s) acc[s] = 0.f;
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = wn[j];
    if (n + 8 < D) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { wn[j] = Wu[(size_t)(n + 8) * 2 * D + lane + 32 * j]; wn[8 + j] = Wi[(size_t)(n + 8) * 2 * D + lane + 32 * j]; }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int s = 0; s < TM_S; ++s) acc[s] += w[j] * xs[s][lane + 32 * j];
    }
#pragma unroll
    for (int s = 0; s < TM_S; ++s) {
      const float v = warp_sum(acc[s]);
      if (lane == 0 && b0 + s < B) y[(size_t)(b0 + s) * D + n] = tanhf(v);
    }
  }
}

// thread k (0..511) owns one input column of [a_u | s_u | a_i | s_i]: d in[b][k] = sum_n dpre[b][n] W[n][k] (coalesced weight rows)
__global__ void __launch_bounds__(512) text_match_bwd_kernel(const float* __restrict__ dpre, const float* __restrict__ Wu,
                                                             const float* __restrict__ Wi, int B, float* __restrict__ d_au,
                                                             float* __restrict__ d_su, float* __restrict__ d_ai, float* __restrict__ d_si) {
  __shared__ float ds[TM_S][D];
  const int k = threadIdx.x, b0 = blockIdx.x * TM_S;
  for (int idx = k; idx < TM_S * D; idx += 512) {
    const int s = idx >> 7, n = idx & 127;
    ds[s][n] = b0 + s < B ? dpre[(size_t)(b0 + s) * D + n] : 0.f;
  }
  __syncthreads();
  const float* W = k < 2 * D ? Wu + k : Wi + (k - 2 * D);
  float acc[TM_S];
#pragma unroll
  for (int s = 0; s < TM_S; ++s) acc[s] = 0.f;
#pragma unroll 16
  for (int n = 0; n < D; ++n) {
    const float w = W[(size_t)n * 2 * D];
#pragma unroll
    for (int s = 0; s < TM_S; ++s) acc[s] += w * ds[s][n];
  }
  float* dst = (k >> 7) == 0 ? d_au : (k >> 7) == 1 ? d_su : (k >> 7) == 2 ? d_ai : d_si;
#pragma unroll
  for (int s = 0; s < TM_S; ++s)
    if (b0 + s < B) dst[(size_t)(b0 + s) * D + (k & 127)] = acc[s];
}

// dW[n][k] += sum_b dpre[b][n] in[b][k] over [a_u | s_u | a_i | s_i] (k = 0..511; dWu takes k < 256, dWi the rest).
// grid (8 blocks of 64 columns, batch splits); thread = (32 outputs n, one column k); atomics fold the batch splits
__global__ void __launch_bounds__(256) text_match_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ a_u,
                                                               const float* __restrict__ s_u, const float* __restrict__ a_i,
                                                               const float* __restrict__ s_i, int B, int per_split,
                                                               float* __restrict__ dWu, float* __restrict__ dWi) {
  __shared__ float dp[8][D];
  __shared__ float xin[8][64];
  const int tid = threadIdx.x, k = tid & 63, n0 = (tid >> 6) * 32;
  const int kb = blockIdx.x, j = kb >> 1, c0 = (kb & 1) * 64;           // input j, its columns c0..c0+63
  const float* src = j == 0 ? a_u : j == 1 ? s_u : j == 2 ? a_i : s_i;
  const int b_beg = blockIdx.y * per_split, b_end = min(B, b_beg + per_split);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int b0 = b_beg; b0 < b_end; b0 += 8) {
    __syncthreads();
    for (int idx = tid; idx < 8 * D; idx += 256) {
      const int s = idx >> 7, n = idx & 127;
      dp[s][n] = b0 + s < b_end ? dpre[(size_t)(b0 + s) * D + n] : 0.f;
    }
    for (int idx = tid; idx < 8 * 64; idx += 256) {
      const int s = idx >> 6, c = idx & 63;
      xin[s][c] = b0 + s < b_end ? src[(size_t)(b0 + s) * D + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const float x = xin[s][k];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += dp[s][n0 + i] * x;
    }
  }
  float* dW = (j < 2 ? dWu : dWi) + (j & 1) * D + c0 + k;
#pragma unroll
  for (int i = 0; i < 32; ++i) atomicAdd(dW + (size_t)(n0 + i) * 2 * D, acc[i]);
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_text_match_wgrad(const float* dpre, const float* atte_u, const float* senti_u, const float* atte_i,
                                     const float* senti_i, int B, float* dWu, float* dWi, void* stream) {
  if (B <= 0) return 0;
  int splits = (B + 63) / 64;
  if (splits > 32) splits = 32;
  const int per_split = ((B + splits - 1) / splits + 7) / 8 * 8;
  text_match_wgrad_kernel<<<dim3(8, splits), 256, 0, (cudaStream_t)stream>>>(dpre, atte_u, senti_u, atte_i, senti_i, B, per_split, dWu, dWi);
  return check_launch("text_match_wgrad");
}

extern "C" int umpr_text_match_fwd(const float* atte_u, const float* senti_u, const float* atte_i, const float* senti_i, const float* Wu,
                                   const float* Wi, int B, float* y, void* stream) {
  if (B <= 0) return 0;
  text_match_fwd_kernel<<<(B + TM_S - 1) / TM_S, 256, 0, (cudaStream_t)stream>>>(atte_u, senti_u, atte_i, senti_i, Wu, Wi, B, y);
  return check_launch("text_match_fwd");
}

extern "C" int umpr_text_match_bwd(const float* dpre, const float* Wu, const float* Wi, int B, float* d_atte_u, float* d_senti_u,
                                   float* d_atte_i, float* d_senti_i, void* stream) {
  if (B <= 0) return 0;
  text_match_bwd_kernel<<<(B + TM_S - 1) / TM_S, 512, 0, (cudaStream_t)stream>>>(dpre, Wu, Wi, B, d_atte_u, d_senti_u, d_atte_i, d_senti_i);
  return check_launch("text_match_bwd");
}
