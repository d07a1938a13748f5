// GRU weight gradients on the tensor cores (backward of reference src/model.py:19; main.py:36):
//   C[256][128] = sum over all packed token slots of dG[slot][256]^T · [xp[slot][64] | h_prev[slot][64]]
// Both operands are "MN-major" (the reduction index - the token slot - is the slow index in memory), which tcgen05 consumes
// directly: shared-memory tiles hold, per slot k, 64 contiguous m (or n) elements per 128-byte row, 8 slots per 1024-byte
// SWIZZLE_128B atom, atoms of consecutive slots SBO = 1024 B apart, the next 64 elements of m/n LBO = 8192 B apart.
// Persistent CTAs own a contiguous range of slots; 64 slots per pipeline stage; the two 128-row halves of C accumulate in
// TMEM over the whole range and are flushed once with atomics into the eight nn.GRU gradient tensors.
//   warps 0-7 loaders (fp32 -> bf16 hi/lo, h_prev gathered from the GRU output), warp 8 MMA issuer, warps 0-3 epilogue.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int WT_THREADS = 288;
constexpr int WT_TILE = 128 * 128;              // bytes: [2 chunks][8 atoms][8 slots][128 B] = 64 slots x 128 elements of bf16
constexpr int WT_STAGE = 6 * WT_TILE;           // A_top hi/lo, A_bot hi/lo, B hi/lo
constexpr int WT_NSTAGE = 2;

// byte offset of elements (mn..mn+3, k) inside a 64-slot x 128-element MN-major tile
__device__ __forceinline__ uint32_t mn_off(int mn, int k) {
  const int i2 = mn >> 6, mi = mn & 63, i1 = mi >> 3, i0 = mi & 7, j1 = k >> 3, j0 = k & 7;
  return (uint32_t)(i2 * 8192 + j1 * 1024 + j0 * 128 + (((i1 ^ j0) << 4) | (i0 << 1)));
}
__device__ __forceinline__ void store_split4_mn(unsigned char* hi, unsigned char* lo, int mn, int k, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  const uint32_t off = mn_off(mn, k);
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
}

__global__ void __launch_bounds__(WT_THREADS, 1) gru_wgrad_tc_kernel(const float* __restrict__ dG, const float* __restrict__ xp,
                                                                     const float* __restrict__ out, Plan p, int L, int E, int slots_per_cta,
                                                                     float* __restrict__ dw_ih_f, float* __restrict__ dw_hh_f,
                                                                     float* __restrict__ db_ih_f, float* __restrict__ db_hh_f,
                                                                     float* __restrict__ dw_ih_b, float* __restrict__ dw_hh_b,
                                                                     float* __restrict__ db_ih_b, float* __restrict__ db_hh_b) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t full_bar[WT_NSTAGE], empty_bar[WT_NSTAGE], acc_bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y, R = p.R;
  const int n_flat = p.n_slabs * R;
  const int f_beg = blockIdx.x * slots_per_cta, f_end = min(n_flat, f_beg + slots_per_cta);
  const int n_steps = (f_end - f_beg + 63) / 64;

  if (tid == 0) {
    for (int s = 0; s < WT_NSTAGE; ++s) { mbar_init(&full_bar[s], 256); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders
    for (int st = 0; st < n_steps; ++st) {
      const int s = st % WT_NSTAGE;
      if (st >= WT_NSTAGE) mbar_wait(&empty_bar[s], ((st / WT_NSTAGE) - 1) & 1);
      unsigned char* sb = base + s * WT_STAGE;
      const int f0 = f_beg + st * 64;
      // thread (warp w, lane): slots k = i*8 + w (i = 0..7), elements lane*4..+3 of each 128-wide half row; the swizzled
      // destination is a per-thread constant plus i*1024 (k>>3 == i, k&7 == w)
      const int m = lane * 4;
      const uint32_t off0 = mn_off(m, warp);
      const float* gsrc[8];
      const float* xsrc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = f0 + i * 8 + warp;
        gsrc[i] = nullptr; xsrc[i] = nullptr;
        if (f < f_end) {
          const int sl = f / R, r = f - sl * R;
          gsrc[i] = dG + (((size_t)sl * 2 + dir) * R + r) * SV + m;
          if (m < KP) {
            xsrc[i] = xp + ((size_t)sl * R + r) * KP + m;
          } else {
            const int j = p.slab_tile[sl], t = sl - p.tile_off[j], kj = j * R + r;
            const int rw = p.row_of[kj], tp = dir ? t + 1 : t - 1;
            if (rw >= 0 && tp >= 0 && tp < p.len_of[kj]) xsrc[i] = out + ((size_t)rw * L + tp) * D + dir * H + (m - KP);
          }
        }
      }
      float4 v[8];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {                 // dG rows: top half (dr, dz) then bottom half (dn, dn*r)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gsrc[i] ? *reinterpret_cast<const float4*>(gsrc[i] + half * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char* hi = sb + (half * 2) * WT_TILE + off0, *lo = hi + WT_TILE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t h0, l0, h1, l1;
          split2(v[i].x, v[i].y, h0, l0);
          split2(v[i].z, v[i].w, h1, l1);
          *reinterpret_cast<uint2*>(hi + i * 1024) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(lo + i * 1024) = make_uint2(l0, l1);
        }
      }
      {                                                       // B rows: [xp (64) | h_prev (64)]
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xsrc[i] ? *reinterpret_cast<const float4*>(xsrc[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char* hi = sb + 4 * WT_TILE + off0, *lo = hi + WT_TILE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t h0, l0, h1, l1;
          split2(v[i].x, v[i].y, h0, l0);
          split2(v[i].z, v[i].w, h1, l1);
          *reinterpret_cast<uint2*>(hi + i * 1024) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(lo + i * 1024) = make_uint2(l0, l1);
        }
      }
      fence_async_smem();
      mbar_arrive(&full_bar[s]);
    }
  } else {
    const uint32_t el = elect_one_sync();      // whole warp converged, the elected lane issues
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);     // A and B are MN-major
    for (int st = 0; st < n_steps; ++st) {
      const int s = st % WT_NSTAGE;
      mbar_wait(&full_bar[s], (st / WT_NSTAGE) & 1);
      tc_fence_after();
      const uint32_t sb = smem_u32(base + s * WT_STAGE);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t ko = kk * 2048;           // 16 slots = two 1024-byte atoms
        const uint64_t bh = smem_desc_mn_sw128(sb + 4 * WT_TILE + ko), bl = smem_desc_mn_sw128(sb + 5 * WT_TILE + ko);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t ah = smem_desc_mn_sw128(sb + (half * 2) * WT_TILE + ko), al = smem_desc_mn_sw128(sb + (half * 2 + 1) * WT_TILE + ko);
          const uint32_t d = tmem + half * 128;
          umma_bf16_e(el, d, ah, bh, idesc, (st | kk) != 0);
          umma_bf16_e(el, d, ah, bl, idesc, 1);
          umma_bf16_e(el, d, al, bh, idesc, 1);
        }
      }
      umma_commit_e(el, &empty_bar[s]);
    }
    umma_commit_e(el, &acc_bar);
  }
  __syncwarp();
  if (warp < 4 && n_steps > 0) {
    // ------------------------------------------------------------------ epilogue: flush the two accumulator halves
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    float* dw_ih = dir ? dw_ih_b : dw_ih_f;
    float* dw_hh = dir ? dw_hh_b : dw_hh_f;
    float* db_ih = dir ? db_ih_b : db_ih_f;
    float* db_hh = dir ? db_hh_b : db_hh_f;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int g = half * 128 + warp * 32 + lane;          // 0..255 : dr, dz, dn, dn*r
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 128 + c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int col = c0 + c;
          if (col < KP) {
            if (g < G3) {
              if (col < E) atomicAdd(&dw_ih[g * E + col], v[c]);
              else if (col == E) { atomicAdd(&db_ih[g], v[c]); if (g < 2 * H) atomicAdd(&db_hh[g], v[c]); }
            } else if (col == E) {
              atomicAdd(&db_hh[g - H], v[c]);
            }
          } else {
            const int hc = col - KP;
            if (g < 2 * H) atomicAdd(&dw_hh[g * H + hc], v[c]);
            else if (g >= G3) atomicAdd(&dw_hh[(g - H) * H + hc], v[c]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------------------------
// Generic "TN" reduction GEMM on the tensor cores: C[M][N] += sum_k A[k][M]^T · B[k][N]   (M, N <= 128, K huge; both operands
// row-major with the reduction index slowest = MN-major for tcgen05, converted fp32 -> bf16 hi/lo on the fly).
// Used for dM = gi^T · dgiM of the co-attention backward (model.py:50; K = B*P).  Persistent CTAs own contiguous K ranges,
// accumulate in TMEM and flush once with atomics.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int TN_STAGE = 4 * WT_TILE;            // A hi/lo, B hi/lo (64 k x 128 elements each)
constexpr int TN_NSTAGE = 3;

__global__ void __launch_bounds__(WT_THREADS, 1) tc_gemm_tn_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                                                                   float* __restrict__ C, long ldc, int M, int N, long K, long k_per_cta) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t full_bar[TN_NSTAGE], empty_bar[TN_NSTAGE], acc_bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long k_beg = (long)blockIdx.x * k_per_cta, k_end = min(K, k_beg + k_per_cta);
  const int n_steps = (int)((k_end - k_beg + 63) / 64);

  if (tid == 0) {
    for (int s = 0; s < TN_NSTAGE; ++s) { mbar_init(&full_bar[s], 256); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // loaders: thread (warp w, lane) owns k = i*8 + w (i = 0..7) and elements lane*4..+3 of both 128-wide rows
    const int m = lane * 4;
    const uint32_t off0 = mn_off(m, warp);
    const bool a_ok = m < M, b_ok = m < N;          // M, N are multiples of 4
    // register double buffer: the 16 row loads of step st+1 are in flight while step st is split and stored
    float4 va[8], vb[8], na[8], nb[8];
    auto fetch = [&](int st, float4* xa, float4* xb) {
      const long k0 = k_beg + (long)st * 64 + warp;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long k = k0 + i * 8;
        const bool in = st < n_steps && k < k_end;
        xa[i] = (in && a_ok) ? *reinterpret_cast<const float4*>(A + k * lda + m) : make_float4(0.f, 0.f, 0.f, 0.f);
        xb[i] = (in && b_ok) ? *reinterpret_cast<const float4*>(B + k * ldb + m) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0, va, vb);
    for (int st = 0; st < n_steps; ++st) {
      const int s = st % TN_NSTAGE;
      fetch(st + 1, na, nb);
      if (st >= TN_NSTAGE) mbar_wait(&empty_bar[s], ((st / TN_NSTAGE) - 1) & 1);
      unsigned char* sb = base + s * TN_STAGE + off0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t h0, l0, h1, l1;
        split2(va[i].x, va[i].y, h0, l0);
        split2(va[i].z, va[i].w, h1, l1);
        *reinterpret_cast<uint2*>(sb + i * 1024) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(sb + WT_TILE + i * 1024) = make_uint2(l0, l1);
        split2(vb[i].x, vb[i].y, h0, l0);
        split2(vb[i].z, vb[i].w, h1, l1);
        *reinterpret_cast<uint2*>(sb + 2 * WT_TILE + i * 1024) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(sb + 3 * WT_TILE + i * 1024) = make_uint2(l0, l1);
      }
      fence_async_smem();
      mbar_arrive(&full_bar[s]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { va[i] = na[i]; vb[i] = nb[i]; }
    }
  } else {
    const uint32_t el = elect_one_sync();      // whole warp converged, the elected lane issues
    constexpr uint32_t idesc = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);     // A and B are MN-major
    for (int st = 0; st < n_steps; ++st) {
      const int s = st % TN_NSTAGE;
      mbar_wait(&full_bar[s], (st / TN_NSTAGE) & 1);
      tc_fence_after();
      const uint32_t sb = smem_u32(base + s * TN_STAGE);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t ko = kk * 2048;           // 16 k = two 1024-byte atoms
        const uint64_t ah = smem_desc_mn_sw128(sb + ko), al = smem_desc_mn_sw128(sb + WT_TILE + ko);
        const uint64_t bh = smem_desc_mn_sw128(sb + 2 * WT_TILE + ko), bl = smem_desc_mn_sw128(sb + 3 * WT_TILE + ko);
        umma_bf16_e(el, tmem, ah, bh, idesc, (st | kk) != 0);
        umma_bf16_e(el, tmem, ah, bl, idesc, 1);
        umma_bf16_e(el, tmem, al, bh, idesc, 1);
      }
      umma_commit_e(el, &empty_bar[s]);
    }
    umma_commit_e(el, &acc_bar);
  }
  __syncwarp();
  if (warp < 4 && n_steps > 0) {
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    const int mrow = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (mrow < M) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c0 + c < N) atomicAdd(&C[(long)mrow * ldc + c0 + c], v[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_gru_wgrad_tc(const float* dG, const float* xp, const float* out, const int32_t* plan, int n_tiles, int n_slabs,
                                 int R, int L, int E, float* const* dw, int n_ctas, void* stream) {
  if (n_slabs == 0) return 0;
  if (R != 32 && R != 64 && R != 128) return fail_arg("gru_wgrad_tc: R=%d", R);
  Plan p = make_plan(plan, n_tiles, n_slabs, R);
  const int n_flat = n_slabs * R;
  if (n_ctas < 2) n_ctas = 2;
  int per = (n_flat + n_ctas / 2 - 1) / (n_ctas / 2);
  per = ((per + 63) / 64) * 64;
  const int grid = (n_flat + per - 1) / per;
  const int smem = WT_NSTAGE * WT_STAGE + 1024;
  cudaError_t e = cudaFuncSetAttribute(gru_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("gru_wgrad_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  gru_wgrad_tc_kernel<<<dim3(grid, 2), WT_THREADS, smem, (cudaStream_t)stream>>>(dG, xp, out, p, L, E, per, dw[0], dw[1], dw[2], dw[3], dw[4],
                                                                                dw[5], dw[6], dw[7]);
  return check_launch("gru_wgrad_tc");
}

// C[M][N] (+=) sum_k A[k*lda + m] * B[k*ldb + n]: tensor-core reduction over a huge K (M, N <= 128, multiples of 4; 16-byte aligned rows)
extern "C" int umpr_tc_gemm_tn(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, long K, int n_ctas,
                               void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M > 128 || N > 128 || (M & 3) || (N & 3)) return fail_arg("tc_gemm_tn: M=%d N=%d (need <= 128, multiples of 4)", M, N);
  if ((lda & 3) || (ldb & 3) || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15))
    return fail_arg("tc_gemm_tn: A and B must be 16-byte aligned with leading dimensions that are multiples of 4");
  if (n_ctas < 1) n_ctas = 148;
  long per = (K + n_ctas - 1) / n_ctas;
  per = ((per + 63) / 64) * 64;
  const int grid = (int)((K + per - 1) / per);
  const int smem = TN_NSTAGE * TN_STAGE + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("tc_gemm_tn smem: %s", cudaGetErrorString(e)); return (int)e; }
  tc_gemm_tn_kernel<<<grid, WT_THREADS, smem, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, M, N, K, per);
  return check_launch("tc_gemm_tn");
}
