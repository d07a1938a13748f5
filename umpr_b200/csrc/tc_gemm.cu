// tcgen05 GEMM, "NT" form: C[M][N] = A[M][K] · B[N][K]^T with fp32 operands split into 3xBF16 on the fly.
// One 128 x BN output tile per CTA; accumulators in TMEM; operands staged by four loader warps (global fp32 ->
// registers -> bf16 hi/lo -> SWIZZLE_128B shared memory), one elected thread issues tcgen05.mma, the loader warps
// then turn into the epilogue (tcgen05.ld -> bias/activation -> global).  2-stage mbarrier ring over K blocks of 64.
//
// Modes: PLAIN  - strided A, B, C (gi·M of model.py:50, its gradient products, text matching model.py:168)
//        INPROJ - GRU input projection for every packed token (model.py:19 input half): A = xp slabs,
//                 B = [W_ih | b_ih(+b_hh for r,z)] assembled on the fly, C = G[slab][dir][R][192]
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int TG_BM = 128, TG_BK = 64, TG_STAGES = 2, TG_THREADS = 160;

struct TcGemmArgs {
  const float* A; long lda;
  const float* B; long ldb;
  float* C; long ldc;
  int M, N, K;
  int accumulate, act;
  int b_kn;              // B given as [K][N] row-major (element (n,k) at B[k*ldb + n]) instead of [N][K]
  const float* bias;
  // INPROJ
  const float* w_ih[2]; const float* b_ih[2]; const float* b_hh[2];
  int E, R;
};

template <int BN> struct TgSmem {
  static constexpr int A_BYTES = TG_BM * 128, B_BYTES = BN * 128;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;        // A_hi, A_lo, B_hi, B_lo
  static constexpr int TOTAL = TG_STAGES * STAGE + 1024;         // + alignment slack
};

template <int BN, int MODE>
__global__ void __launch_bounds__(TG_THREADS, 1) tc_gemm_nt_kernel(const TcGemmArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  using SM = TgSmem<BN>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TG_BM, n0 = blockIdx.y * BN;
  const int dir = blockIdx.z;
  constexpr uint32_t TCOLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  const int nkb = (a.K + TG_BK - 1) / TG_BK;

  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TG_STAGES;
      if (kb >= TG_STAGES) mbar_wait(&empty_bar[s], ((kb / TG_STAGES) - 1) & 1);
      unsigned char* st = base + s * SM::STAGE;
      unsigned char* a_hi = st, *a_lo = st + SM::A_BYTES, *b_hi = st + 2 * SM::A_BYTES, *b_lo = b_hi + SM::B_BYTES;
      const int k0 = kb * TG_BK;
      // A tile: 128 rows x 64 k  (half-warp per row, float4 per lane)
      float4 va[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = i * 128 + tid, r = idx >> 4, k = (idx & 15) * 4;
        const int m = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M) {
          const float* p = a.A + (long)m * a.lda + k0 + k;
          if (k0 + k + 3 < a.K) v = *reinterpret_cast<const float4*>(p);
          else {
            if (k0 + k < a.K) v.x = p[0];
            if (k0 + k + 1 < a.K) v.y = p[1];
            if (k0 + k + 2 < a.K) v.z = p[2];
          }
        }
        va[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = i * 128 + tid;
        store_split4(a_hi, a_lo, idx >> 4, (idx & 15) * 4, va[i]);
      }
      // B tile: BN rows x 64 k
#pragma unroll
      for (int h = 0; h < BN / 64; ++h) {
        float4 vb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int idx = (h * 8 + i) * 128 + tid, r = idx >> 4, k = (idx & 15) * 4;
          if (MODE == 0 && a.b_kn) { r = h * 64 + (tid & 63); k = ((tid >> 6) * 8 + i) * 4; }   // lanes along n: coalesced
          const int n = n0 + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < a.N) {
            if (MODE == 1) {
              const float* w = a.w_ih[dir] + (long)n * a.E;
              float t[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int kk = k + q;
                t[q] = kk < a.E ? w[kk] : (kk == a.E ? a.b_ih[dir][n] + (n < 2 * H ? a.b_hh[dir][n] : 0.f) : 0.f);
              }
              v = make_float4(t[0], t[1], t[2], t[3]);
            } else if (a.b_kn) {
              const float* p = a.B + (long)(k0 + k) * a.ldb + n;
              if (k0 + k < a.K) v.x = p[0];
              if (k0 + k + 1 < a.K) v.y = p[a.ldb];
              if (k0 + k + 2 < a.K) v.z = p[2 * a.ldb];
              if (k0 + k + 3 < a.K) v.w = p[3 * a.ldb];
            } else {
              const float* p = a.B + (long)n * a.ldb + k0 + k;
              if (k0 + k + 3 < a.K) v = *reinterpret_cast<const float4*>(p);
              else {
                if (k0 + k < a.K) v.x = p[0];
                if (k0 + k + 1 < a.K) v.y = p[1];
                if (k0 + k + 2 < a.K) v.z = p[2];
              }
            }
          }
          vb[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int idx = (h * 8 + i) * 128 + tid, r = idx >> 4, k = (idx & 15) * 4;
          if (MODE == 0 && a.b_kn) { r = h * 64 + (tid & 63); k = ((tid >> 6) * 8 + i) * 4; }
          store_split4(b_hi, b_lo, r, k, vb[i]);
        }
      }
      fence_async_smem();            // make the generic-proxy stores visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[s]);
    }
    // ------------------------------------------------------------------ epilogue
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane, m = m0 + r;
    float* crow = nullptr;
    if (m < a.M) {
      if (MODE == 1) {
        const int slab = m / a.R, rr = m - slab * a.R;
        crow = a.C + (((long)slab * 2 + dir) * a.R + rr) * G3;
      } else {
        crow = a.C + (long)m * a.ldc;
      }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (crow) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = n0 + c0 + q * 4;
          float o[4] = {v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
          if (n + 3 < a.N) {
            if (MODE == 0) {
              if (a.accumulate) { const float4 c = *reinterpret_cast<const float4*>(crow + n); o[0] += c.x; o[1] += c.y; o[2] += c.z; o[3] += c.w; }
              if (a.bias) { o[0] += a.bias[n]; o[1] += a.bias[n + 1]; o[2] += a.bias[n + 2]; o[3] += a.bias[n + 3]; }
              if (a.act == 1) { o[0] = tanhf(o[0]); o[1] = tanhf(o[1]); o[2] = tanhf(o[2]); o[3] = tanhf(o[3]); }
              else if (a.act == 2) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
            }
            *reinterpret_cast<float4*>(crow + n) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            for (int e = 0; e < 4; ++e) {
              if (n + e < a.N) {
                float x = o[e];
                if (MODE == 0) {
                  if (a.accumulate) x += crow[n + e];
                  if (a.bias) x += a.bias[n + e];
                  if (a.act == 1) x = tanhf(x); else if (a.act == 2) x = fmaxf(x, 0.f);
                }
                crow[n + e] = x;
              }
            }
          }
        }
      }
    }
  } else {
    const uint32_t el = elect_one_sync();      // whole warp converged, the elected lane issues
    // ------------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = idesc_bf16(TG_BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TG_STAGES;
      mbar_wait(&full_bar[s], (kb / TG_STAGES) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(base + s * SM::STAGE);
      const uint64_t ah = smem_desc_sw128(st), al = smem_desc_sw128(st + SM::A_BYTES);
      const uint64_t bh = smem_desc_sw128(st + 2 * SM::A_BYTES), bl = smem_desc_sw128(st + 2 * SM::A_BYTES + SM::B_BYTES);
#pragma unroll
      for (int kk = 0; kk < TG_BK / 16; ++kk) {
        const uint64_t o = (uint64_t)(kk * 2);      // +32 bytes (>>4) inside the 128-byte swizzled row
        umma_bf16_e(el, tmem, ah + o, bh + o, idesc, (kb | kk) != 0);
        umma_bf16_e(el, tmem, ah + o, bl + o, idesc, 1);
        umma_bf16_e(el, tmem, al + o, bh + o, idesc, 1);
      }
      umma_commit_e(el, &empty_bar[s]);     // frees the stage when these MMAs have read it
    }
    umma_commit_e(el, &accum_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TCOLS);
}

template <int BN, int MODE> static int launch_tc(const TcGemmArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = TgSmem<BN>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_nt_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("tc_gemm smem: %s", cudaGetErrorString(e)); return (int)e; }
  tc_gemm_nt_kernel<BN, MODE><<<grid, TG_THREADS, smem, st>>>(a);
  return check_launch("tc_gemm_nt");
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_tc_gemm_nt(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K,
                               int accumulate, const float* bias, int act, int b_kn, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda & 3) || (ldb & 3) || (ldc & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(C) & 15))
    return fail_arg("tc_gemm_nt: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  if (act < 0 || act > 2) return fail_arg("tc_gemm_nt: act=%d", act);
  TcGemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.accumulate = accumulate; a.act = act; a.bias = bias; a.b_kn = b_kn;
  const int gm = (M + TG_BM - 1) / TG_BM;
  if (N <= 64) return launch_tc<64, 0>(a, dim3(gm, 1, 1), (cudaStream_t)stream);
  if (N <= 128) return launch_tc<128, 0>(a, dim3(gm, 1, 1), (cudaStream_t)stream);
  return launch_tc<128, 0>(a, dim3(gm, (N + 127) / 128, 1), (cudaStream_t)stream);
}

