// Strided fp32 GEMM on CUDA cores with split-K, bias and activation epilogues.
// Used for the small dense contractions around the fused kernels: gi·M (model.py:50), its two
// backward products, the text-matching linears (model.py:168) and their gradients.
//   C[m][n] = act( beta*C[m][n] + sum_k A(m,k) * B(k,n) + bias[n] ),  A(m,k) = A[m*ars + k*acs],  B(k,n) = B[k*brs + n*bcs]
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long ars, long acs,
                                                    const float* __restrict__ B, long brs, long bcs,
                                                    float* __restrict__ C, long ldc, int M, int N, int K, int k_chunk,
                                                    int accumulate, const float* __restrict__ bias, int act) {
  __shared__ __align__(16) float As[BK][LDS_];
  __shared__ __align__(16) float Bs[BK][LDS_];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);
  const bool a_kvec = (acs == 1) && ((ars & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_nvec = (bcs == 1) && ((brs & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (a_kvec && k0 + BK <= kend) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = tid + i * 256, m = e >> 2, k4 = e & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + m < M) v = *reinterpret_cast<const float4*>(A + (long)(m0 + m) * ars + k0 + k4 * 4);
        As[k4 * 4 + 0][m] = v.x; As[k4 * 4 + 1][m] = v.y; As[k4 * 4 + 2][m] = v.z; As[k4 * 4 + 3][m] = v.w;
      }
    } else if (acs == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * 256, m = e >> 4, k = e & 15;
        As[k][m] = (m0 + m < M && k0 + k < kend) ? A[(long)(m0 + m) * ars + (k0 + k)] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * 256, m = e & 127, k = e >> 7;
        As[k][m] = (m0 + m < M && k0 + k < kend) ? A[(long)(m0 + m) * ars + (long)(k0 + k) * acs] : 0.f;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (b_nvec && n0 + BN <= N && k0 + BK <= kend) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = tid + i * 256, k = e >> 5, n4 = e & 31;
        *reinterpret_cast<float4*>(&Bs[k][n4 * 4]) = *reinterpret_cast<const float4*>(B + (long)(k0 + k) * brs + n0 + n4 * 4);
      }
    } else if (bcs == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * 256, n = e & 127, k = e >> 7;
        Bs[k][n] = (n0 + n < N && k0 + k < kend) ? B[(long)(k0 + k) * brs + (n0 + n)] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * 256, n = e >> 4, k = e & 15;
        Bs[k][n] = (n0 + n < N && k0 + k < kend) ? B[(long)(k0 + k) * brs + (long)(n0 + n) * bcs] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= N) continue;
      float* c = C + (long)m * ldc + n;
      if (split) {
        atomicAdd(c, acc[i][j]);
      } else {
        float v = acc[i][j];
        if (accumulate) v += *c;
        if (bias) v += bias[n];
        if (act == 1) v = tanhf(v);
        else if (act == 2) v = fmaxf(v, 0.f);
        else if (act == 3) v = sigmoidf_acc(v);
        *c = v;
      }
    }
  }
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_sgemm(const float* A, long ars, long acs, const float* B, long brs, long bcs, float* C, long ldc, int M,
                          int N, int K, int splits, int accumulate, const float* bias, int act, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K < 0) return fail_arg("sgemm: K=%d", K);
  if (splits < 1) splits = 1;
  if (splits > 1 && (bias || act)) return fail_arg("sgemm: split-K cannot apply bias/activation");
  int k_chunk = (K + splits - 1) / splits;
  k_chunk = ((k_chunk + BK - 1) / BK) * BK;
  if (k_chunk == 0) k_chunk = BK;
  splits = (K + k_chunk - 1) / k_chunk;
  if (splits < 1) splits = 1;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, ars, acs, B, brs, bcs, C, ldc, M, N, K, k_chunk, accumulate, bias, act);
  return check_launch("sgemm");
}
