// C-Net tail (reference src/model.py:118-125): Conv1d(128 -> K, k=3, pad=1) + ReLU + max over L, then
// Linear(K -> V) + Sigmoid, threshold, sum of squares over sentences.  Forward and backward.
// The convolution runs over ALL L positions (padding positions see zeros and produce relu(bias), which competes
// in the max exactly as in the reference); it is computed as an implicit GEMM (K' = 3*128) fused with the max-pool,
// so the (N, K, L) activation never reaches HBM.  The backward is sparse: one arg-max position per (sentence, filter).
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

constexpr int CK = 3;              // kernel_size (config.py:37)
constexpr int CKP = 128;           // filters padded to 128 columns
constexpr int CX_LD = D + 4;       // 132
constexpr int CROWS = 128;         // MAC tile rows (sentences incl. one zero guard row before and after each)
constexpr int CCH = 32;            // k' chunk streamed through shared memory

// wt[(dt*128 + c)][kf] = W[kf][c][dt]  (kf >= KC -> 0)
__global__ void cnet_prep_kernel(const float* __restrict__ w, int KC, float* __restrict__ wt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= CK * D * CKP) return;
  const int kf = idx & (CKP - 1), kp = idx >> 7;
  const int dt = kp >> 7, c = kp & 127;
  wt[idx] = kf < KC ? w[((size_t)kf * D + c) * CK + dt] : 0.f;
}

__global__ void __launch_bounds__(256, 1) cnet_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                               const float* __restrict__ bias, int N, int L, int KC, int gs,
                                                               float* __restrict__ cfeat, int* __restrict__ cidx) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                          // [(CROWS + 2)][132]; tile row r lives at xs row r + 1
  float* ys = xs + (CROWS + 2) * CX_LD;      // [CROWS][132]  relu(conv)
  float* wc = ys + CROWS * CX_LD;            // [2][CCH][128] weight chunks
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int Lg = L + 2;
  const int n_groups = (N + gs - 1) / gs;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int n0 = grp * gs;
    const int ns = min(gs, N - n0);
    __syncthreads();
    // x rows with zero guards
    for (int idx = tid; idx < (CROWS + 2) * (D / 4); idx += 256) {
      const int rr = idx >> 5, c4 = idx & 31;          // rr: xs row
      const int r = rr - 1;                            // tile row
      const int s = r >= 0 ? r / Lg : -1;
      const int l = r >= 0 ? r - s * Lg - 1 : -1;      // position inside the sentence, -1 / L are guards
      if (s >= 0 && s < ns && l >= 0 && l < L)
        cp_async16(&xs[rr * CX_LD + c4 * 4], x + ((size_t)(n0 + s) * L + l) * D + c4 * 4);
      else
        *reinterpret_cast<float4*>(&xs[rr * CX_LD + c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    auto load_w = [&](int ch) {
      const float4* src = reinterpret_cast<const float4*>(wt + (size_t)ch * CCH * CKP);
      float* dst = wc + (ch & 1) * CCH * CKP;
      for (int idx = tid; idx < CCH * CKP / 4; idx += 256) cp_async16(dst + idx * 4, src + idx);
      cp_async_commit();
    };
    load_w(0);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
    constexpr int NCH = CK * D / CCH;   // 12
    for (int ch = 0; ch < NCH; ++ch) {
      cp_async_wait_all();
      __syncthreads();                  // chunk ch (and, first time, xs) visible; chunk ch-1 fully consumed
      if (ch + 1 < NCH) load_w(ch + 1);
      const float* wcur = wc + (ch & 1) * CCH * CKP;
      const int dt = ch >> 2, c0 = (ch & 3) * CCH;
#pragma unroll 2
      for (int k4 = 0; k4 < CCH / 4; ++k4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)   // tile row r needs x row (r - 1 + dt) -> xs row (r + dt)
          a[i] = *reinterpret_cast<const float4*>(&xs[(ty * 8 + i + dt) * CX_LD + c0 + k4 * 4]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 b0 = *reinterpret_cast<const float4*>(&wcur[(k4 * 4 + kk) * CKP + tx * 4]);
          const float4 b1 = *reinterpret_cast<const float4*>(&wcur[(k4 * 4 + kk) * CKP + 64 + tx * 4]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
            acc[i][0] += av * b0.x; acc[i][1] += av * b0.y; acc[i][2] += av * b0.z; acc[i][3] += av * b0.w;
            acc[i][4] += av * b1.x; acc[i][5] += av * b1.y; acc[i][6] += av * b1.z; acc[i][7] += av * b1.w;
          }
        }
      }
    }
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int kf = c < 4 ? tx * 4 + c : 64 + tx * 4 + c - 4;
      bv[c] = kf < KC ? bias[kf] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* o = &ys[(ty * 8 + i) * CX_LD];
      *reinterpret_cast<float4*>(o + tx * 4) = make_float4(fmaxf(acc[i][0] + bv[0], 0.f), fmaxf(acc[i][1] + bv[1], 0.f),
                                                           fmaxf(acc[i][2] + bv[2], 0.f), fmaxf(acc[i][3] + bv[3], 0.f));
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(fmaxf(acc[i][4] + bv[4], 0.f), fmaxf(acc[i][5] + bv[5], 0.f),
                                                                fmaxf(acc[i][6] + bv[6], 0.f), fmaxf(acc[i][7] + bv[7], 0.f));
    }
    __syncthreads();
    // max over the L positions of each sentence (model.py:120); first maximum wins, <= 0 carries no gradient
    for (int idx = tid; idx < ns * KC; idx += 256) {
      const int s = idx / KC, kf = idx - s * KC;
      float best = -1.f;
      int arg = -1;
      for (int l = 0; l < L; ++l) {
        const float v = ys[(s * Lg + 1 + l) * CX_LD + kf];
        if (v > best) { best = v; arg = l; }
      }
      cfeat[(size_t)(n0 + s) * KC + kf] = best;
      cidx[(size_t)(n0 + s) * KC + kf] = best > 0.f ? arg : -1;
    }
  }
}

// view_p = where(sigmoid(Wc cfeat + bc) < thr, 0, .) (model.py:123-124);  final = sum_s view_p^2 (model.py:125)
__global__ void __launch_bounds__(128) cnet_head_fwd_kernel(const float* __restrict__ cfeat, const float* __restrict__ lw,
                                                            const float* __restrict__ lb, float thr, int S_, int V, int KC,
                                                            float* __restrict__ view_p, float* __restrict__ final_) {
  extern __shared__ float ps[];   // [S_*V]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int pr = warp; pr < S_ * V; pr += 4) {
    const int s = pr / V, v = pr - s * V;
    const float* f = cfeat + ((size_t)b * S_ + s) * KC;
    float a = 0.f;
    for (int k = lane; k < KC; k += 32) a += f[k] * lw[v * KC + k];
    a = warp_sum(a);
    if (lane == 0) {
      float p = sigmoidf_acc(a + lb[v]);
      if (p < thr) p = 0.f;
      ps[pr] = p;
      view_p[(size_t)b * S_ * V + pr] = p;
    }
  }
  __syncthreads();
  if (tid < V) {
    float a = 0.f;
    for (int s = 0; s < S_; ++s) { const float p = ps[s * V + tid]; a += p * p; }
    final_[(size_t)b * V + tid] = a;
  }
}

__global__ void __launch_bounds__(128) cnet_head_bwd_kernel(const float* __restrict__ cfeat, const int* __restrict__ cidx,
                                                            const float* __restrict__ view_p, const float* __restrict__ lw,
                                                            const float* __restrict__ d_view_p, const float* __restrict__ d_final,
                                                            int S_, int V, int KC, float* __restrict__ dcfeat,
                                                            float* __restrict__ d_lw, float* __restrict__ d_lb,
                                                            float* __restrict__ d_cb) {
  extern __shared__ float sm[];   // dpre [S_*V]
  float* dpre = sm;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int pr = tid; pr < S_ * V; pr += 128) {
    const int v = pr % V;
    const float p = view_p[(size_t)b * S_ * V + pr];
    float dp = d_view_p ? d_view_p[(size_t)b * S_ * V + pr] : 0.f;
    if (d_final) dp += 2.f * p * d_final[(size_t)b * V + v];
    dpre[pr] = p > 0.f ? dp * p * (1.f - p) : 0.f;
  }
  __syncthreads();
  if (tid < KC) {
    float dcb = 0.f;
    for (int s = 0; s < S_; ++s) {
      const size_t n = (size_t)b * S_ + s;
      float a = 0.f;
      for (int v = 0; v < V; ++v) a += dpre[s * V + v] * lw[v * KC + tid];
      if (cidx[n * KC + tid] < 0) a = 0.f;     // ReLU clipped the maximum: no gradient reaches the conv
      dcfeat[n * KC + tid] = a;
      dcb += a;
    }
    atomicAdd(&d_cb[tid], dcb);
    for (int v = 0; v < V; ++v) {
      float a = 0.f;
      for (int s = 0; s < S_; ++s) a += dpre[s * V + v] * cfeat[((size_t)b * S_ + s) * KC + tid];
      atomicAdd(&d_lw[v * KC + tid], a);
    }
  }
  if (tid < V) {
    float a = 0.f;
    for (int s = 0; s < S_; ++s) a += dpre[s * V + tid];
    atomicAdd(&d_lb[tid], a);
  }
}

// dx[n][l][c] = sum_{kf, dt : arg[kf] + dt - 1 == l} g[kf] W[kf][c][dt].   Thread (q, c) keeps its 32x3 weights in registers and
// owns column c of its own shared-memory copy q of the sentence gradient (plain read-modify-write, no atomics: shared-memory
// float atomics are CAS loops on this architecture); the NQ copies are summed on the way out.
template <int NQ>
__global__ void __launch_bounds__(128 * NQ, 1) cnet_conv_bwd_dx_kernel(const float* __restrict__ dcfeat, const int* __restrict__ cidx,
                                                                       const float* __restrict__ w, int N, int L, int KC,
                                                                       float* __restrict__ dx) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = 128 * NQ, FPQ = 128 / NQ;     // filters per group
  const int rows = L + 2;
  float* dxs = smem;                               // [NQ][(L+2)][128]
  float* gsm = dxs + NQ * rows * D;                // [128]
  int* tsm = reinterpret_cast<int*>(gsm + CKP);
  const int tid = threadIdx.x, q = tid >> 7, c = tid & 127;
  float wr[FPQ][CK];
#pragma unroll
  for (int i = 0; i < FPQ; ++i) {
    const int kf = q * FPQ + i;
#pragma unroll
    for (int dt = 0; dt < CK; ++dt) wr[i][dt] = kf < KC ? w[((size_t)kf * D + c) * CK + dt] : 0.f;
  }
  float* mine = dxs + q * rows * D + c;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    for (int r = 0; r < rows; ++r) mine[r * D] = 0.f;
    if (tid < CKP) {
      gsm[tid] = tid < KC ? dcfeat[(size_t)n * KC + tid] : 0.f;
      tsm[tid] = tid < KC ? cidx[(size_t)n * KC + tid] : -1;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < FPQ; ++i) {
      const float g = gsm[q * FPQ + i];
      const int t = tsm[q * FPQ + i];
      if (g != 0.f && t >= 0) {          // warp-uniform
#pragma unroll
        for (int dt = 0; dt < CK; ++dt) mine[(t + dt) * D] += g * wr[i][dt];     // x position t+dt-1 -> guard row +1
      }
    }
    __syncthreads();
    for (int idx = tid; idx < L * (D / 4); idx += NT) {
      const int l = idx >> 5, c4 = idx & 31;
      float4 a = *reinterpret_cast<const float4*>(&dxs[(l + 1) * D + c4 * 4]);
#pragma unroll
      for (int qq = 1; qq < NQ; ++qq) {
        const float4 b = *reinterpret_cast<const float4*>(&dxs[(qq * rows + l + 1) * D + c4 * 4]);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      *reinterpret_cast<float4*>(dx + ((size_t)n * L + l) * D + c4 * 4) = a;
    }
  }
}

// w[kf][c][dt] -> wt[kf][dt][c]: a warp's 128 channels of one (filter, tap) become one 512-byte row
__global__ void cnet_wt_kernel(const float* __restrict__ w, int KC, float* __restrict__ wt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= KC * CK * D) return;
  const int c = idx & (D - 1), dt = (idx >> 7) % CK, kf = idx / (CK * D);
  wt[idx] = w[((size_t)kf * D + c) * CK + dt];
}

// dx by a sorted sweep (replaces the shared-memory read-modify-write scatter above for the hot path).  ONE WARP per sentence,
// each lane owns 4 channels (float4).  The sentence's <= 128 (arg-max position, gradient, filter) triples are counting-sorted
// by position with warp ballots (deterministic: ties in filter order), then one pass over them keeps a sliding window of THREE
// float4 accumulators per lane - rows p-1, p, p+1 - and emits each output row exactly once, in order, as a 512-byte store.
// No atomics, no shared-memory accumulation; weights come from the tap-major copy wt[filter][tap][channel] (L1/L2 resident).
constexpr int SW_WARPS = 8;
__global__ void __launch_bounds__(SW_WARPS * 32, 4) cnet_conv_bwd_dx_sweep_kernel(const float* __restrict__ dcfeat, const int* __restrict__ cidx,
                                                                                 const float* __restrict__ wt, int N, int L, int KC,
                                                                                 const int* __restrict__ cst, float* __restrict__ dx) {
  __shared__ float2 s_sorted[SW_WARPS][CKP];     // (gradient, position | filter << 16) in sweep order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (int n = blockIdx.x * SW_WARPS + warp; n < N; n += gridDim.x * SW_WARPS) {
    // each lane holds filters lane, lane+32, lane+64, lane+96
    float g[4];
    int t[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kf = lane + 32 * q;
      g[q] = kf < KC ? dcfeat[(size_t)n * KC + kf] : 0.f;
      t[q] = kf < KC ? cidx[(size_t)n * KC + kf] : -1;
      if (g[q] == 0.f) t[q] = -1;                               // no gradient through this filter
    }
    __syncwarp();
    // counting sort by position: start[p] = number of valid filters with a smaller position; rank inside a position = filter order
    int base = 0;
    int rank[4] = {0, 0, 0, 0};
    for (int p = 0; p < L; ++p) {
      int cnt = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, t[q] == p);
        if (t[q] == p) rank[q] = base + cnt + __popc(m & lt);
        cnt += __popc(m);
      }
      base += cnt;
    }
    const int nvalid = base;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (t[q] >= 0) s_sorted[warp][rank[q]] = make_float2(g[q], __int_as_float(t[q] | ((lane + 32 * q) << 16)));
    __syncwarp();
    float* orow = dx + (size_t)n * L * D + lane * 4;
    // with a length table only the rows below the sentence's length are written (the others are never read: model.py:18)
    const int len = cst ? min(L, cst[n + 1] - cst[n]) : L;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;        // rows cur-1, cur, cur+1
    int cur = 0;
    for (int i = 0; i < nvalid; ++i) {
      const float2 e = s_sorted[warp][i];
      const int pk = __float_as_int(e.y), p = pk & 0xffff, kf = pk >> 16;
      if (p > len) break;                      // sorted by position: everything from here on lands on rows >= len
      while (cur < p) {                        // slide the window: row cur-1 is complete
        if (cur >= 1) *reinterpret_cast<float4*>(orow + (size_t)(cur - 1) * D) = a0;
        a0 = a1; a1 = a2; a2 = make_float4(0.f, 0.f, 0.f, 0.f); ++cur;
      }
      const float4* wk = reinterpret_cast<const float4*>(wt + (size_t)kf * CK * D) + lane;
      const float4 w0 = wk[0], w1 = wk[D / 4], w2 = wk[2 * D / 4];        // taps 0, 1, 2 -> x positions p-1, p, p+1
      a0.x = fmaf(e.x, w0.x, a0.x); a0.y = fmaf(e.x, w0.y, a0.y); a0.z = fmaf(e.x, w0.z, a0.z); a0.w = fmaf(e.x, w0.w, a0.w);
      a1.x = fmaf(e.x, w1.x, a1.x); a1.y = fmaf(e.x, w1.y, a1.y); a1.z = fmaf(e.x, w1.z, a1.z); a1.w = fmaf(e.x, w1.w, a1.w);
      a2.x = fmaf(e.x, w2.x, a2.x); a2.y = fmaf(e.x, w2.y, a2.y); a2.z = fmaf(e.x, w2.z, a2.z); a2.w = fmaf(e.x, w2.w, a2.w);
    }
    while (cur <= len) {                       // flush: every (valid) row of dx is written exactly once (zeros where nothing landed)
      if (cur >= 1) *reinterpret_cast<float4*>(orow + (size_t)(cur - 1) * D) = a0;
      a0 = a1; a1 = a2; a2 = make_float4(0.f, 0.f, 0.f, 0.f); ++cur;
    }
    __syncwarp();
  }
}

// dW[kf][c][dt] += sum_n g[n][kf] x[n][arg + dt - 1][c].  Thread (q, c) accumulates its 32x3 slice in registers.
// G sentences per pipeline stage (one barrier pair and one cp.async group per G sentences); with a length table only the rows
// below each sentence's length are staged - taps that fall on a zero row are skipped.
__global__ void __launch_bounds__(512, 1) cnet_conv_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ dcfeat,
                                                                  const int* __restrict__ cidx, int N, int L, int KC, int G,
                                                                  const int* __restrict__ cst, float* __restrict__ dw) {
  extern __shared__ __align__(16) float smem[];
  const int sent_f = L * D + 2 * CKP;          // per sentence: rows [L][128], gradients [128], arg-max positions [128]
  const int stage_f = G * sent_f;
  // thread = (warp w: filters 8w..8w+7, lane: channels 4*lane..+3): the filter loop is warp-uniform (no divergence on the
  // zero-gradient test), every tap is one conflict-free 128-bit shared-memory load feeding four FMAs
  const int tid = threadIdx.x, fg = tid >> 5, c4 = (tid & 31) * 4;
  const int n_groups = (N + G - 1) / G;
  float4 acc[8][CK];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int dt = 0; dt < CK; ++dt) acc[i][dt] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto sent_len = [&](int n) { return cst ? min(L, cst[n + 1] - cst[n]) : L; };
  auto issue = [&](int grp, int st) {
    for (int j = 0; j < G; ++j) {
      const int n = grp * G + j;
      if (n >= N) break;
      float* xs = smem + st * stage_f + j * sent_f;
      const int len = sent_len(n);
      const float4* src = reinterpret_cast<const float4*>(x + (size_t)n * L * D);
      for (int idx = tid; idx < len * (D / 4); idx += 512) cp_async16(&xs[idx * 4], src + idx);
      if (tid < CKP) {
        float* gsm = xs + L * D;
        int* tsm = reinterpret_cast<int*>(gsm + CKP);
        if (tid < KC) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&gsm[tid])), "l"(dcfeat + (size_t)n * KC + tid));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&tsm[tid])), "l"(cidx + (size_t)n * KC + tid));
        } else {
          gsm[tid] = 0.f; tsm[tid] = -1;
        }
      }
    }
    cp_async_commit();
  };
  int it = 0;
  if ((int)blockIdx.x < n_groups) issue(blockIdx.x, 0);
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
    const int st = it & 1;
    const int ng = grp + gridDim.x;
    if (ng < n_groups) { issue(ng, st ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else cp_async_wait_all();
    __syncthreads();                         // stage st has landed for every thread
    for (int j = 0; j < G; ++j) {
      const int n = grp * G + j;
      if (n >= N) break;
      const float* xs = smem + st * stage_f + j * sent_f;
      const float* gsm = xs + L * D;
      const int* tsm = reinterpret_cast<const int*>(gsm + CKP);
      const int len = sent_len(n);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float g = gsm[fg * 8 + i];
        const int t = tsm[fg * 8 + i];
        if (g != 0.f && t >= 0) {
#pragma unroll
          for (int dt = 0; dt < CK; ++dt) {
            const int row = t + dt - 1;
            if (row >= 0 && row < len) {
              const float4 v = *reinterpret_cast<const float4*>(&xs[row * D + c4]);
              acc[i][dt].x += g * v.x; acc[i][dt].y += g * v.y; acc[i][dt].z += g * v.z; acc[i][dt].w += g * v.w;
            }
          }
        }
      }
    }
    __syncthreads();                         // stage st may be refilled by the next iteration's prefetch
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kf = fg * 8 + i;
    if (kf < KC) {
#pragma unroll
      for (int dt = 0; dt < CK; ++dt) {
        float* d = dw + ((size_t)kf * D + c4) * CK + dt;
        atomicAdd(d, acc[i][dt].x); atomicAdd(d + CK, acc[i][dt].y); atomicAdd(d + 2 * CK, acc[i][dt].z); atomicAdd(d + 3 * CK, acc[i][dt].w);
      }
    }
  }
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_cnet_prep(const float* conv_w, int KC, int ksize, float* wt, void* stream) {
  if (ksize != CK) return fail_arg("cnet: kernel_size=%d (only %d is built)", ksize, CK);
  if (KC < 1 || KC > CKP) return fail_arg("cnet: kernel_count=%d must be in [1, %d]", KC, CKP);
  const int n = CK * D * CKP;
  cnet_prep_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conv_w, KC, wt);
  return check_launch("cnet_prep");
}

extern "C" int umpr_cnet_conv_fwd(const float* x, const float* wt, const float* conv_b, int N, int L, int KC, float* cfeat,
                                  int32_t* cidx, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (L < 1 || L + 2 > CROWS) return fail_arg("cnet_conv_fwd: sentence length L=%d must be in [1, %d]", L, CROWS - 2);
  if (KC < 1 || KC > CKP) return fail_arg("cnet: kernel_count=%d", KC);
  int gs = CROWS / (L + 2);
  if (gs > 16) gs = 16;
  const int n_groups = (N + gs - 1) / gs;
  const size_t sm = sizeof(float) * ((CROWS + 2) * CX_LD + CROWS * CX_LD + 2 * CCH * CKP);
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) { set_error("cnet_conv_fwd smem: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = n_ctas > 0 && n_ctas < n_groups ? n_ctas : n_groups;
  cnet_conv_fwd_kernel<<<grid, 256, sm, (cudaStream_t)stream>>>(x, wt, conv_b, N, L, KC, gs, cfeat, cidx);
  return check_launch("cnet_conv_fwd");
}

extern "C" int umpr_cnet_head_fwd(const float* cfeat, const float* lin_w, const float* lin_b, float threshold, int B, int S_, int V,
                                  int KC, float* view_p, float* final_, void* stream) {
  if (B <= 0) return 0;
  if (V < 1 || V > 128) return fail_arg("cnet_head: view_size=%d must be in [1,128]", V);
  cnet_head_fwd_kernel<<<B, 128, sizeof(float) * S_ * V, (cudaStream_t)stream>>>(cfeat, lin_w, lin_b, threshold, S_, V, KC, view_p, final_);
  return check_launch("cnet_head_fwd");
}

extern "C" int umpr_cnet_head_bwd(const float* cfeat, const int32_t* cidx, const float* view_p, const float* lin_w,
                                  const float* d_view_p, const float* d_final, int B, int S_, int V, int KC, float* dcfeat,
                                  float* d_lin_w, float* d_lin_b, float* d_conv_b, void* stream) {
  if (B <= 0) return 0;
  if (V < 1 || V > 128 || KC > 128) return fail_arg("cnet_head_bwd: V=%d KC=%d", V, KC);
  cnet_head_bwd_kernel<<<B, 128, sizeof(float) * S_ * V, (cudaStream_t)stream>>>(cfeat, cidx, view_p, lin_w, d_view_p, d_final, S_, V,
                                                                               KC, dcfeat, d_lin_w, d_lin_b, d_conv_b);
  return check_launch("cnet_head_bwd");
}

// cst (optional): exclusive prefix sum (N+1, int32, device) of the sentence lengths for an x produced by ImprovedRnn - rows at or
// beyond a sentence's length are exactly zero, are not staged, and their dx rows are not written.
extern "C" int umpr_cnet_conv_bwd_dx(const float* dcfeat, const int32_t* cidx, const float* conv_w, int N, int L, int KC,
                                     const int32_t* cst, float* wt_scratch, float* dx, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (KC < 1 || KC > CKP) return fail_arg("cnet: kernel_count=%d", KC);
  const int grid = n_ctas > 0 && n_ctas < N ? n_ctas : N;
  const size_t sm4 = sizeof(float) * (4 * (L + 2) * D + CKP) + sizeof(int) * CKP;
  const size_t sm2 = sizeof(float) * (2 * (L + 2) * D + CKP) + sizeof(int) * CKP;
  if (wt_scratch && KC <= CKP && L < 0x7fff) {
    cnet_wt_kernel<<<(KC * CK * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conv_w, KC, wt_scratch);
    const int g4 = (N + SW_WARPS - 1) / SW_WARPS;
    const int want = 8 * (n_ctas > 0 ? n_ctas : 148);
    cnet_conv_bwd_dx_sweep_kernel<<<g4 < want ? g4 : want, SW_WARPS * 32, 0, (cudaStream_t)stream>>>(dcfeat, cidx, wt_scratch, N, L, KC, cst, dx);
  } else if (sm4 <= 200 * 1024) {
    cudaFuncSetAttribute(cnet_conv_bwd_dx_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
    cnet_conv_bwd_dx_kernel<4><<<grid, 512, sm4, (cudaStream_t)stream>>>(dcfeat, cidx, conv_w, N, L, KC, dx);
  } else if (sm2 <= 200 * 1024) {
    cudaFuncSetAttribute(cnet_conv_bwd_dx_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    cnet_conv_bwd_dx_kernel<2><<<grid, 256, sm2, (cudaStream_t)stream>>>(dcfeat, cidx, conv_w, N, L, KC, dx);
  } else {
    return fail_arg("cnet_conv_bwd_dx: L=%d too large", L);
  }
  return check_launch("cnet_conv_bwd_dx");
}

extern "C" int umpr_cnet_conv_bwd_dw(const float* x, const float* dcfeat, const int32_t* cidx, int N, int L, int KC, const int32_t* cst,
                                     float* d_conv_w, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (KC < 1 || KC > CKP) return fail_arg("cnet: kernel_count=%d", KC);
  const size_t sent_b = sizeof(float) * ((size_t)L * D + 2 * CKP);
  int G = (int)((90 * 1024) / sent_b);
  if (G > 2) G = 2;
  if (G < 1) G = 1;
  const size_t sm = 2 * G * sent_b;
  if (sm > 200 * 1024) return fail_arg("cnet_conv_bwd_dw: L=%d too large", L);
  cudaFuncSetAttribute(cnet_conv_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int n_groups = (N + G - 1) / G;
  const int grid_w = n_ctas > 0 && n_ctas < n_groups ? n_ctas : n_groups;
  cnet_conv_bwd_dw_kernel<<<grid_w, 512, sm, (cudaStream_t)stream>>>(x, dcfeat, cidx, N, L, KC, G, cst, d_conv_w);
  return check_launch("cnet_conv_bwd_dw");
}
