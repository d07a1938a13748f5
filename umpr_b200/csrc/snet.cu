// S-Net sentence self-attention (reference src/model.py:71-81), forward and backward.
//   score[n,l] = Ws · tanh(Ms · x[n,l]);  soft = softmax_l(score) over ALL L positions (unmasked);
//   self_atte[n] = sum_l soft[n,l] x[n,l];  sentiment[b] = sum_s (sum_q word_soft[b,s,q]) self_atte[b,s]
// One streaming read of gru_repr; a CTA owns whole sentences (<= 128 rows) so the softmax never leaves the SM.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

constexpr int XLD = D + 4;      // 132
constexpr int TLD = ATT + 4;    // 68
constexpr int SN_ROWS = 128;    // largest row tile (and the longest sentence supported)
constexpr int SN_MAXG = 16;     // sentences per CTA tile

// ROWS = rows of one CTA tile: 128, or 64 when the sentences are short enough - half the shared memory, two CTAs per SM, so
// one CTA's loads and barrier phases overlap the other's math
template <int ROWS>
struct SnetSmemT {
  float xs[ROWS * XLD];         // x rows, row-major
  float ths[ROWS * TLD];        // tanh(Ms x) rows (forward: scratch; backward: th then d(pre-tanh))
  float ms[D * TLD];            // forward: Ms^T [k][a] ; backward: Ms [a][c] uses [ATT][XLD] (same bytes: 128*68 == 64*132+256)
  float score[ROWS];
  float soft[ROWS];
  float ws[ATT];
  float red[32];
};
static_assert(D * TLD >= ATT * XLD, "ms buffer reuse");

template <int ROWS>
__device__ __forceinline__ void load_rows(float* dst, int ld, const float* __restrict__ src, int rows, int width, int tid) {
  // rows x width floats (width % 4 == 0) -> dst[r*ld + c]; rows..ROWS zero-filled
  const int w4 = width >> 2;
  for (int idx = tid; idx < ROWS * w4; idx += 256) {
    const int r = idx / w4, c4 = idx - r * w4;
    if (r < rows) cp_async16(&dst[r * ld + c4 * 4], src + (size_t)r * width + c4 * 4);
    else *reinterpret_cast<float4*>(&dst[r * ld + c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_commit();
}

template <int ROWS>
__global__ void __launch_bounds__(256, ROWS == 64 ? 2 : 1) snet_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Ms,
                                                          const float* __restrict__ Ws, int N, int L, int gs,
                                                          float* __restrict__ self_atte, float* __restrict__ soft_out,
                                                          float* __restrict__ th_out) {
  extern __shared__ __align__(16) unsigned char raw[];
  using Smem = SnetSmemT<ROWS>;
  constexpr int RPT = ROWS / 16;           // rows per thread of the 16 x 16 thread grid
  Smem& S = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
  for (int idx = tid; idx < ATT * D; idx += 256) {
    const int a = idx >> 7, k = idx & 127;
    S.ms[k * TLD + a] = Ms[idx];
  }
  if (tid < ATT) S.ws[tid] = Ws[tid];
  const int n_groups = (N + gs - 1) / gs;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int n0 = grp * gs;
    const int ns = min(gs, N - n0);
    const int rows = ns * L;
    __syncthreads();
    load_rows<ROWS>(S.xs, XLD, x + (size_t)n0 * L * D, rows, D, tid);
    cp_async_wait_all();
    __syncthreads();
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int k4 = 0; k4 < D / 4; ++k4) {
      float4 a[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(&S.xs[(ty * RPT + i) * XLD + k4 * 4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b = *reinterpret_cast<const float4*>(&S.ms[(k4 * 4 + kk) * TLD + tx * 4]);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
          acc[i][0] += av * b.x; acc[i][1] += av * b.y; acc[i][2] += av * b.z; acc[i][3] += av * b.w;
        }
      }
    }
    const float4 w4 = *reinterpret_cast<const float4*>(&S.ws[tx * 4]);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = ty * RPT + i;
      const float t0 = tanhf(acc[i][0]), t1 = tanhf(acc[i][1]), t2 = tanhf(acc[i][2]), t3 = tanhf(acc[i][3]);
      float sc = t0 * w4.x + t1 * w4.y + t2 * w4.z + t3 * w4.w;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      if (tx == 0) S.score[r] = sc;
      if (th_out && r < rows)
        *reinterpret_cast<float4*>(th_out + ((size_t)n0 * L + r) * ATT + tx * 4) = make_float4(t0, t1, t2, t3);
    }
    __syncthreads();
    // softmax over the L positions of each sentence: one warp per sentence
    for (int s = (tid >> 5); s < ns; s += 8) {
      float mx = -INFINITY;
      for (int l = lane; l < L; l += 32) mx = fmaxf(mx, S.score[s * L + l]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int l = lane; l < L; l += 32) { const float e = expf(S.score[s * L + l] - mx); S.soft[s * L + l] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int l = lane; l < L; l += 32) {
        const float v = S.soft[s * L + l] * inv;
        S.soft[s * L + l] = v;
        if (soft_out) soft_out[(size_t)(n0 + s) * L + l] = v;
      }
    }
    __syncthreads();
    // self_atte[s][c] = sum_l soft[s,l] x[s,l,c]
    for (int idx = tid; idx < ns * (D / 4); idx += 256) {
      const int s = idx >> 5, c4 = idx & 31;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int l = 0; l < L; ++l) {
        const float w = S.soft[s * L + l];
        const float4 v = *reinterpret_cast<const float4*>(&S.xs[(s * L + l) * XLD + c4 * 4]);
        a.x += w * v.x; a.y += w * v.y; a.z += w * v.z; a.w += w * v.w;
      }
      *reinterpret_cast<float4*>(self_atte + (size_t)(n0 + s) * D + c4 * 4) = a;
    }
  }
}

// wsum[n] = sum_q word_soft[n, q] (model.py:79);  sentiment[b] = sum_s wsum[b,s] self_atte[b,s] (model.py:80)
__global__ void __launch_bounds__(128) snet_sentiment_fwd_kernel(const float* __restrict__ self_atte, const float* __restrict__ word_soft,
                                                                 int S_, int Wd, float* __restrict__ wsum, float* __restrict__ sentiment) {
  extern __shared__ float w_s[];   // [S_]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int s = warp; s < S_; s += 4) {
    float a = 0.f;
    const float* src = word_soft + ((size_t)b * S_ + s) * Wd;
    for (int q = lane; q < Wd; q += 32) a += src[q];
    a = warp_sum(a);
    if (lane == 0) { w_s[s] = a; wsum[(size_t)b * S_ + s] = a; }
  }
  __syncthreads();
  float acc = 0.f;
  for (int s = 0; s < S_; ++s) acc += w_s[s] * self_atte[((size_t)b * S_ + s) * D + tid];
  sentiment[(size_t)b * D + tid] = acc;
}

// d_self_atte[n] = wsum[n] d_sentiment[b] (+ upstream);  d_wsum[n] = <self_atte[n], d_sentiment[b]>
__global__ void __launch_bounds__(128) snet_sentiment_bwd_kernel(const float* __restrict__ self_atte, const float* __restrict__ wsum,
                                                                 const float* __restrict__ d_sentiment, const float* __restrict__ d_sa_up,
                                                                 int S_, float* __restrict__ d_sa, float* __restrict__ d_wsum) {
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float ds = d_sentiment ? d_sentiment[(size_t)b * D + tid] : 0.f;
  for (int s = 0; s < S_; ++s) {
    const size_t n = (size_t)b * S_ + s;
    float v = (wsum ? wsum[n] : 0.f) * ds;
    if (d_sa_up) v += d_sa_up[n * D + tid];
    d_sa[n * D + tid] = v;
    if (d_wsum) {
      const float dot = block_sum(self_atte[n * D + tid] * ds, red);
      if (tid == 0) d_wsum[n] = dot;
    }
  }
}

template <int ROWS>
__global__ void __launch_bounds__(256, ROWS == 64 ? 2 : 1) snet_bwd_kernel(const float* __restrict__ x, const float* __restrict__ th,
                                                          const float* __restrict__ soft, const float* __restrict__ d_sa,
                                                          const float* __restrict__ Ms, const float* __restrict__ Ws, int N, int L,
                                                          int gs, float* __restrict__ dx, float* __restrict__ dMs,
                                                          float* __restrict__ dWs) {
  extern __shared__ __align__(16) unsigned char raw[];
  using Smem = SnetSmemT<ROWS>;
  constexpr int RPT = ROWS / 16;
  Smem& S = *reinterpret_cast<Smem*>(raw);
  float* dsa = reinterpret_cast<float*>(raw + sizeof(Smem));    // [SN_MAXG][128]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  float* msn = S.ms;                                                 // Ms natural layout [a][XLD]
  for (int idx = tid; idx < ATT * D; idx += 256) {
    const int a = idx >> 7, c = idx & 127;
    msn[a * XLD + c] = Ms[idx];
  }
  if (tid < ATT) S.ws[tid] = Ws[tid];
  float dms_acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) dms_acc[i][c] = 0.f;
  float dws_acc = 0.f;
  const int n_groups = (N + gs - 1) / gs;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int n0 = grp * gs;
    const int ns = min(gs, N - n0);
    const int rows = ns * L;
    __syncthreads();
    load_rows<ROWS>(S.xs, XLD, x + (size_t)n0 * L * D, rows, D, tid);
    load_rows<ROWS>(S.ths, TLD, th + (size_t)n0 * L * ATT, rows, ATT, tid);
    for (int idx = tid; idx < ROWS; idx += 256) S.soft[idx] = idx < rows ? soft[(size_t)n0 * L + idx] : 0.f;
    for (int idx = tid; idx < ns * D; idx += 256) dsa[idx] = d_sa[(size_t)n0 * D + idx];
    cp_async_wait_all();
    __syncthreads();
    // d_soft[r] = <x[r], d_self_atte[s(r)]>
    for (int r = warp; r < rows; r += 8) {
      const int s = r / L;
      const float4 v = *reinterpret_cast<const float4*>(&S.xs[r * XLD + lane * 4]);
      const float4 d = *reinterpret_cast<const float4*>(&dsa[s * D + lane * 4]);
      float a = v.x * d.x + v.y * d.y + v.z * d.z + v.w * d.w;
      a = warp_sum(a);
      if (lane == 0) S.score[r] = a;
    }
    __syncthreads();
    // softmax backward per sentence -> d_score in S.score
    for (int s = warp; s < ns; s += 8) {
      float dot = 0.f;
      for (int l = lane; l < L; l += 32) dot += S.soft[s * L + l] * S.score[s * L + l];
      dot = warp_sum(dot);
      for (int l = lane; l < L; l += 32) S.score[s * L + l] = S.soft[s * L + l] * (S.score[s * L + l] - dot);
    }
    for (int r = rows + tid; r < ROWS; r += 256) S.score[r] = 0.f;
    __syncthreads();
    // through Ws and tanh: ths <- d(pre-tanh);  dWs += sum_r d_score[r] th[r]
    {
      const int a = tid & 63, rg = tid >> 6;
      const float wsa = S.ws[a];
      for (int r = rg; r < rows; r += 4) {
        const float t = S.ths[r * TLD + a], dsc = S.score[r];
        dws_acc += dsc * t;
        S.ths[r * TLD + a] = dsc * wsa * (1.f - t * t);
      }
    }
    __syncthreads();
    // dx[r][c] = soft[r] d_sa[s][c] + sum_a dpre[r][a] Ms[a][c]
    {
      float acc[RPT][8];
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
      for (int k4 = 0; k4 < ATT / 4; ++k4) {
        float4 a[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(&S.ths[(ty * RPT + i) * TLD + k4 * 4]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 b0 = *reinterpret_cast<const float4*>(&msn[(k4 * 4 + kk) * XLD + tx * 4]);
          const float4 b1 = *reinterpret_cast<const float4*>(&msn[(k4 * 4 + kk) * XLD + 64 + tx * 4]);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
            acc[i][0] += av * b0.x; acc[i][1] += av * b0.y; acc[i][2] += av * b0.z; acc[i][3] += av * b0.w;
            acc[i][4] += av * b1.x; acc[i][5] += av * b1.y; acc[i][6] += av * b1.z; acc[i][7] += av * b1.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = ty * RPT + i;
        if (r >= rows) continue;
        const int s = r / L;
        const float so = S.soft[r];
        const float4 d0 = *reinterpret_cast<const float4*>(&dsa[s * D + tx * 4]);
        const float4 d1 = *reinterpret_cast<const float4*>(&dsa[s * D + 64 + tx * 4]);
        float* o = dx + ((size_t)n0 * L + r) * D;
        *reinterpret_cast<float4*>(o + tx * 4) =
            make_float4(acc[i][0] + so * d0.x, acc[i][1] + so * d0.y, acc[i][2] + so * d0.z, acc[i][3] + so * d0.w);
        *reinterpret_cast<float4*>(o + 64 + tx * 4) =
            make_float4(acc[i][4] + so * d1.x, acc[i][5] + so * d1.y, acc[i][6] + so * d1.z, acc[i][7] + so * d1.w);
      }
    }
    // dMs[a][c] += sum_r dpre[r][a] x[r][c]   (zero rows beyond `rows` contribute nothing)
    for (int r = 0; r < ROWS; ++r) {
      if (r >= rows) break;
      const float4 a4 = *reinterpret_cast<const float4*>(&S.ths[r * TLD + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&S.xs[r * XLD + tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&S.xs[r * XLD + 64 + tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) dms_acc[i][c] += a[i] * bb[c];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      atomicAdd(&dMs[(ty * 4 + i) * D + (c < 4 ? tx * 4 + c : 64 + tx * 4 + c - 4)], dms_acc[i][c]);
  atomicAdd(&dWs[tid & 63], dws_acc);
}

}  // namespace umpr

using namespace umpr;

static int snet_group(int L, int rows) {
  int gs = rows / L;
  if (gs > SN_MAXG) gs = SN_MAXG;
  return gs;
}
// 64-row tiles (two CTAs per SM) when they waste no more rows than 128-row tiles do
static int snet_rows(int L) { return (L <= 64 && (64 / L) * 2 >= 128 / L) ? 64 : 128; }

extern "C" int umpr_snet_fwd(const float* x, const float* Ms, const float* Ws, int N, int L, float* self_atte, float* soft,
                             float* th, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (L < 1 || L > SN_ROWS) return fail_arg("snet_fwd: sentence length L=%d must be in [1, %d]", L, SN_ROWS);
  const int rows = 128;            // the forward has few barrier phases: 128-row tiles measured faster than 2 CTAs of 64 rows
  const int gs = snet_group(L, rows);
  const int n_groups = (N + gs - 1) / gs;
  const size_t sm = rows == 64 ? sizeof(SnetSmemT<64>) : sizeof(SnetSmemT<128>);
  auto kern = rows == 64 ? snet_fwd_kernel<64> : snet_fwd_kernel<128>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) { set_error("snet_fwd smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (rows == 64) n_ctas *= 2;
  const int grid = n_ctas > 0 && n_ctas < n_groups ? n_ctas : n_groups;
  kern<<<grid, 256, sm, (cudaStream_t)stream>>>(x, Ms, Ws, N, L, gs, self_atte, soft, th);
  return check_launch("snet_fwd");
}

extern "C" int umpr_snet_sentiment_fwd(const float* self_atte, const float* word_soft, int B, int S_, int Wd, float* wsum,
                                       float* sentiment, void* stream) {
  if (B <= 0) return 0;
  snet_sentiment_fwd_kernel<<<B, 128, sizeof(float) * S_, (cudaStream_t)stream>>>(self_atte, word_soft, S_, Wd, wsum, sentiment);
  return check_launch("snet_sentiment_fwd");
}

extern "C" int umpr_snet_sentiment_bwd(const float* self_atte, const float* wsum, const float* d_sentiment, const float* d_sa_up,
                                       int B, int S_, float* d_sa, float* d_wsum, void* stream) {
  if (B <= 0) return 0;
  snet_sentiment_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(self_atte, wsum, d_sentiment, d_sa_up, S_, d_sa, d_wsum);
  return check_launch("snet_sentiment_bwd");
}

extern "C" int umpr_snet_bwd(const float* x, const float* th, const float* soft, const float* d_sa, const float* Ms, const float* Ws,
                             int N, int L, float* dx, float* dMs, float* dWs, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (L < 1 || L > SN_ROWS) return fail_arg("snet_bwd: sentence length L=%d must be in [1, %d]", L, SN_ROWS);
  const int rows = snet_rows(L);
  const int gs = snet_group(L, rows);
  const int n_groups = (N + gs - 1) / gs;
  const size_t sm = (rows == 64 ? sizeof(SnetSmemT<64>) : sizeof(SnetSmemT<128>)) + sizeof(float) * SN_MAXG * D;
  auto kern = rows == 64 ? snet_bwd_kernel<64> : snet_bwd_kernel<128>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) { set_error("snet_bwd smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (rows == 64) n_ctas *= 2;
  const int grid = n_ctas > 0 && n_ctas < n_groups ? n_ctas : n_groups;
  kern<<<grid, 256, sm, (cudaStream_t)stream>>>(x, th, soft, d_sa, Ms, Ws, N, L, gs, dx, dMs, dWs);
  return check_launch("snet_bwd");
}
