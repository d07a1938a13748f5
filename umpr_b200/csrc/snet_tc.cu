// S-Net sentence self-attention (reference src/model.py:71-81) on tcgen05, forward and backward, for inputs that come out of
// ImprovedRnn: rows at or beyond a sentence's length are exactly zero (pad_packed_sequence, model.py:20), so tanh(Ms·0) = 0 and
// their score is exactly 0.  Only the VALID rows are loaded and multiplied: consecutive sentences are compacted into 128-row
// tiles (host table, plan.py:snet_table), the padded positions enter the softmax analytically ((L - len)·exp(0 - max)).
//   forward : e = x·Ms^T (3xBF16, 24 MMAs per tile) -> tanh, score, per-sentence softmax, pooling          -> self_atte
//   backward: recomputes e (nothing but x is saved), then per tile  dx = dpre·Ms (12 MMAs)  and  dMs^T += x^T·dpre (24 MMAs,
//             accumulated in tensor memory over the CTA's whole queue); the x operand image is used K-major for e and
//             MN-major for dMs^T, the dpre image K-major for dx and MN-major for dMs^T, the Ms image K-major / MN-major alike.
//             dx is written for valid rows only: the rows beyond a sentence's length are never read downstream (the packed GRU
//             drops them, model.py:18).
// Roles (288 threads): warps 0-3 loaders (global fp32 -> bf16 hi/lo SWIZZLE_128B images), warp 4 MMA issuer, warps 5-8 row
// threads (TMEM lane = compact row).
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int ST_THREADS = 288;
constexpr int ST_NSTAGE = 2;
constexpr int ST_XIMG = 65536;        // [kb 2][hi|lo][128 rows][128 B]
constexpr int ST_MSIMG = 32768;       // [kb 2][hi|lo][64 rows][128 B]
constexpr int ST_PIMG = 32768;        // [hi|lo][128 rows][128 B]
constexpr int ST_STG_LD = 36;
constexpr int ST_NMETA = 4;

struct StMeta {
  int s0, ns, rows, pad;
  int sbase[132];                     // first compact row of each sentence of the tile (+ end)
  int rowmap[128];                    // compact row -> global row (n*L + l)
  float dsoft[128];                   // backward: <x[row], d_self_atte[sentence]>
  short rsent[128];                   // compact row -> sentence index inside the tile
};

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// tanh(x) = 2 / (1 + 2^(-2x log2 e)) - 1 with ex2.approx / rcp.approx (absolute error ~2e-7, the same cell math as the GRU kernels):
// 6 instructions instead of the ~45 of tanhf, which dominated the row threads' serial chain
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.885390082f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(2.f, r, -1.f);
}

// Ms [64][128] fp32 -> resident images (all threads of the CTA)
__device__ __forceinline__ void load_ms_images(unsigned char* msi, const float* __restrict__ Ms, int tid) {
  for (int idx = tid; idx < ATT * 32; idx += ST_THREADS) {
    const int a = idx >> 5, c = (idx & 31) * 4, kb = c >> 6;
    const float4 v = *reinterpret_cast<const float4*>(Ms + a * D + c);
    unsigned char* t = msi + kb * 16384;
    store_split4(t, t + 8192, a, c & 63, v);
  }
}

// tile bookkeeping by the loader threads: sentence bases, row maps (table values from the pipeline above)
__device__ __forceinline__ int build_meta(StMeta& m, const TabPipe& tp, int L, int tid) {
  const int s0 = tp.s0c, ns = tp.s1c - s0, c0 = tp.c0c;
  if (tid < ns) {
    const int b = tp.cbc - c0, e = tp.cec - c0;
    m.sbase[tid] = b;
    if (tid == ns - 1) m.sbase[ns] = e;
    const int g0 = (s0 + tid) * L - b;
    for (int r = b; r < e; ++r) { m.rowmap[r] = g0 + r; m.rsent[r] = (short)tid; }
  }
  if (tid == 0) { m.s0 = s0; m.ns = ns; }
  return tp.cendc - c0;
}

// the e = x·Ms^T product of one tile (24 MMAs)
__device__ __forceinline__ void issue_scores(uint32_t el, uint32_t d, uint32_t a0, uint32_t b0) {
  constexpr uint32_t idesc = idesc_bf16(128, ATT);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    const uint64_t ah = smem_desc_sw128(a0 + kb * 32768), al = smem_desc_sw128(a0 + kb * 32768 + 16384);
    const uint64_t bh = smem_desc_sw128(b0 + kb * 16384), bl = smem_desc_sw128(b0 + kb * 16384 + 8192);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t o = (uint64_t)(kk * 2);
      umma_bf16_e(el, d, ah + o, bh + o, idesc, (kb | kk) != 0);
      umma_bf16_e(el, d, ah + o, bl + o, idesc, 1);
      umma_bf16_e(el, d, al + o, bh + o, idesc, 1);
    }
  }
}

// softmax over the L positions of sentence j of the tile; positions beyond its length have score 0 exactly
__device__ __forceinline__ void sentence_softmax(const StMeta& m, int j, int L, const float* score, float* soft, float& pad_soft) {
  const int b = m.sbase[j], e = m.sbase[j + 1];
  const int npad = L - (e - b);
  float mx = npad > 0 ? 0.f : -INFINITY;
  for (int r = b; r < e; ++r) mx = fmaxf(mx, score[r]);
  float sum = npad > 0 ? (float)npad * expf(-mx) : 0.f;
  for (int r = b; r < e; ++r) { const float ex = expf(score[r] - mx); soft[r] = ex; sum += ex; }
  const float inv = 1.f / sum;
  for (int r = b; r < e; ++r) soft[r] *= inv;
  pad_soft = expf(-mx) * inv;
}

// ======================================================================================================= forward
// 416 threads: warps 0-7 loaders (a warp per row and pass: the 16 row loads of a thread - the whole tile - are in flight together),
// warp 8 MMA issuer, warps 9-12 row threads.  The pooling reads the tile's operand image in shared memory (x = hi + lo to 2^-18
// relative) instead of fetching every row from L2 a second time, so the image stage is released by the row threads.
constexpr int SF_THREADS = 416;

__global__ void __launch_bounds__(SF_THREADS, 1) snet_fwd_tc_kernel(const float* __restrict__ x, const int* __restrict__ tso,
                                                                    const int* __restrict__ cst, const float* __restrict__ Ms,
                                                                    const float* __restrict__ Ws, int n_tiles, int L,
                                                                    float* __restrict__ self_atte, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t a_full[ST_NSTAGE], a_empty[ST_NSTAGE], acc_full[2], acc_empty[2], m_full[ST_NMETA];
  __shared__ uint32_t tmem_slot;
  __shared__ float ws_s[ATT], score[128], soft[128];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* msi = base;
  unsigned char* xim = base + ST_MSIMG;
  StMeta* meta = reinterpret_cast<StMeta*>(base + ST_MSIMG + ST_NSTAGE * ST_XIMG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < ST_NSTAGE; ++s) { mbar_init(&a_full[s], 256); mbar_init(&a_empty[s], 128); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    for (int s = 0; s < ST_NMETA; ++s) mbar_init(&m_full[s], 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 128);
  for (int idx = tid; idx < ATT * 32; idx += SF_THREADS) {      // Ms [64][128] fp32 -> resident images
    const int a = idx >> 5, c = (idx & 31) * 4, kb = c >> 6;
    const float4 v = *reinterpret_cast<const float4*>(Ms + a * D + c);
    unsigned char* t = msi + kb * 16384;
    store_split4(t, t + 8192, a, c & 63, v);
  }
  if (tid < ATT) ws_s[tid] = Ws[tid];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    TabPipe tp;
    tp.init(tso, cst, blockIdx.x, gridDim.x, n_tiles, tid);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % ST_NSTAGE;
      tp.prefetch(tso, cst, tile, gridDim.x, n_tiles, tid);
      if (it >= ST_NSTAGE) mbar_wait(&a_empty[s], ((it / ST_NSTAGE) - 1) & 1);
      StMeta& m = meta[it % ST_NMETA];
      const int rows = build_meta(m, tp, L, tid);
      bar_sync(1, 256);
      unsigned char* st = xim + s * ST_XIMG + (lane >> 4) * 32768;      // channel half of this lane
      float4 va[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = i * 8 + warp;
        va[i] = (r < rows && !(dbg & 2)) ? *reinterpret_cast<const float4*>(x + (size_t)m.rowmap[r] * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!(dbg & 4)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) store_split4(st, st + 16384, i * 8 + warp, (lane & 15) * 4, va[i]);
      }
      fence_async_smem();
      mbar_arrive(&a_full[s]);
      mbar_arrive(&m_full[it % ST_NMETA]);
      tp.rotate();
    }
  } else if (warp == 8) {
    const uint32_t el = elect_one_sync();
    const uint32_t b0 = smem_u32(msi);
    constexpr uint32_t idesc = idesc_bf16(128, ATT);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % ST_NSTAGE, acc = it & 1;
      if (it >= 2) mbar_wait(&acc_empty[acc], ((it >> 1) - 1) & 1);
      mbar_wait(&a_full[s], (it / ST_NSTAGE) & 1);
      tc_fence_after();
      const uint32_t a0 = smem_u32(xim + s * ST_XIMG), d = tmem + acc * ATT;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t ah = smem_desc_sw128(a0 + kb * 32768), al = smem_desc_sw128(a0 + kb * 32768 + 16384);
        const uint64_t bh = smem_desc_sw128(b0 + kb * 16384), bl = smem_desc_sw128(b0 + kb * 16384 + 8192);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          if (dbg & 8) break;
          umma_bf16_e(el, d, ah + o, bh + o, idesc, (kb | kk) != 0);
          umma_bf16_e(el, d, ah + o, bl + o, idesc, 1);
          umma_bf16_e(el, d, al + o, bh + o, idesc, 1);
        }
      }
      umma_commit_e(el, &acc_full[acc]);
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, et = tid - 288;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % ST_NSTAGE, acc = it & 1;
      // the loaders' bookkeeping of this tile is visible (its own barrier: a_full may already be a phase ahead by now)
      mbar_wait(&m_full[it % ST_NMETA], (it / ST_NMETA) & 1);
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const StMeta& m = meta[it % ST_NMETA];
      float sc = 0.f;
      if (!(dbg & 1)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * ATT + h * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) sc += ws_s[h * 32 + i] * tanh_fast(v[i]);
        }
      }
      score[r] = sc;
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      bar_sync(2, 128);
      // softmax over the L positions of each sentence, a thread per ROW (the rows of a sentence recompute its max and sum in the same
      // order - a thread per sentence would leave all but ~15 of the 128 threads idle behind the longest sentence)
      {
        const int rows = m.sbase[m.ns];
        float ex = 0.f, mx = 0.f;
        int b = 0, e = 0, npad = 0;
        if (r < rows && !(dbg & 16)) {
          const int j = m.rsent[r];
          b = m.sbase[j]; e = m.sbase[j + 1]; npad = L - (e - b);
          mx = npad > 0 ? 0.f : -INFINITY;
          for (int rr = b; rr < e; ++rr) mx = fmaxf(mx, score[rr]);
          ex = expf(score[r] - mx);
          soft[r] = ex;
        }
        bar_sync(2, 128);
        if (r < rows && !(dbg & 16)) {
          float sum = npad > 0 ? (float)npad * expf(-mx) : 0.f;
          for (int rr = b; rr < e; ++rr) sum += soft[rr];
          ex *= 1.f / sum;
        }
        bar_sync(2, 128);
        if (r < rows) soft[r] = ex;
      }
      bar_sync(2, 128);
      // pooling: self_atte[n] = sum_l soft[l] x[n,l]  (a warp per sentence, lanes = channel quads, rows from the operand image)
      const unsigned char* img = xim + s * ST_XIMG + (lane >> 4) * 32768;
      const int k = (lane & 15) * 4;
      for (int j = (warp - 9); j < m.ns && !(dbg & 32); j += 4) {
        const int b = m.sbase[j], e = m.sbase[j + 1];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int rr = b; rr < e; ++rr) {
          const uint32_t off = sw128_off(rr, k);
          const uint2 hi = *reinterpret_cast<const uint2*>(img + off), lo = *reinterpret_cast<const uint2*>(img + 16384 + off);
          const float w = soft[rr];
          a.x += w * (__uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16));
          a.y += w * (__uint_as_float(hi.x & 0xffff0000u) + __uint_as_float(lo.x & 0xffff0000u));
          a.z += w * (__uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16));
          a.w += w * (__uint_as_float(hi.y & 0xffff0000u) + __uint_as_float(lo.y & 0xffff0000u));
        }
        *reinterpret_cast<float4*>(self_atte + (size_t)(m.s0 + j) * D + lane * 4) = a;
      }
      mbar_arrive(&a_empty[s]);            // (the generic-proxy reads of the image are ordered before the loaders' next writes by the barrier)
      bar_sync(2, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

// ======================================================================================================= backward
struct StBwdBars {
  uint64_t a_full[ST_NSTAGE], a_empty[ST_NSTAGE], m_full[ST_NMETA], e_full[2], e_empty[2], p_ready, d_full, w_full;
};

__global__ void __launch_bounds__(ST_THREADS, 1) snet_bwd_tc_kernel(const float* __restrict__ x, const int* __restrict__ tso,
                                                                    const int* __restrict__ cst, const float* __restrict__ d_sa,
                                                                    const float* __restrict__ Ms, const float* __restrict__ Ws,
                                                                    int n_tiles, int L, float* __restrict__ dx,
                                                                    float* __restrict__ dMs, float* __restrict__ dWs, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ StBwdBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float ws_s[ATT], score[2][128], soft[2][128], dsc[2][128];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* msi = base;
  unsigned char* pim = base + ST_MSIMG;
  unsigned char* xim = base + ST_MSIMG + ST_PIMG;
  float* stg = reinterpret_cast<float*>(base + ST_MSIMG + ST_PIMG + ST_NSTAGE * ST_XIMG);
  StMeta* meta = reinterpret_cast<StMeta*>(reinterpret_cast<unsigned char*>(stg) + 4 * 32 * ST_STG_LD * 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t T_E = 0, T_DX = 128, T_W = 384;      // e: 2 x 64 columns, dx: 2 x 128, dMs^T: 64
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;

  if (tid == 0) {
    for (int s = 0; s < ST_NSTAGE; ++s) { mbar_init(&bar.a_full[s], 128); mbar_init(&bar.a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.e_full[s], 1); mbar_init(&bar.e_empty[s], 128); }
    for (int s = 0; s < ST_NMETA; ++s) mbar_init(&bar.m_full[s], 128);
    mbar_init(&bar.p_ready, 128);
    mbar_init(&bar.d_full, 1);
    mbar_init(&bar.w_full, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, 512);
  load_ms_images(msi, Ms, tid);
  if (tid < ATT) ws_s[tid] = Ws[tid];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (+ d_soft = <x row, d_self_atte of its sentence>)
    TabPipe tp;
    tp.init(tso, cst, blockIdx.x, gridDim.x, n_tiles, tid);
    for (int it = 0; it < n_mine; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it % ST_NSTAGE;
      tp.prefetch(tso, cst, tile, gridDim.x, n_tiles, tid);
      if (it >= ST_NSTAGE) mbar_wait(&bar.a_empty[s], ((it / ST_NSTAGE) - 1) & 1);
      StMeta& m = meta[it % ST_NMETA];
      const int rows = build_meta(m, tp, L, tid);
      bar_sync(1, 128);
      const int s0 = tp.s0c;
      unsigned char* st = xim + s * ST_XIMG;
      float dot[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) dot[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        unsigned char* a_hi = st + kb * 32768, *a_lo = a_hi + 16384;
        float4 va[16], vd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid, r = idx >> 4, k = kb * 64 + (idx & 15) * 4;
          const bool ok = r < rows && !(dbg & 2);
          va[i] = ok ? *reinterpret_cast<const float4*>(x + (size_t)m.rowmap[r] * D + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          vd[i] = ok ? *reinterpret_cast<const float4*>(d_sa + (size_t)(s0 + m.rsent[r]) * D + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid;
          store_split4(a_hi, a_lo, idx >> 4, (idx & 15) * 4, va[i]);
          dot[i] += va[i].x * vd[i].x + va[i].y * vd[i].y + va[i].z * vd[i].z + va[i].w * vd[i].w;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = dot[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 15) == 0) m.dsoft[i * 8 + (tid >> 4)] = v;
      }
      fence_async_smem();
      mbar_arrive(&bar.a_full[s]);
      mbar_arrive(&bar.m_full[it % ST_NMETA]);
      tp.rotate();
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    if (n_mine > 0) {       // whole warp converged, the elected lane issues
      const uint32_t el = elect_one_sync();
      const uint32_t b0 = smem_u32(msi), p0 = smem_u32(pim);
      constexpr uint32_t idesc_dx = idesc_bf16(128, 128) | (1u << 16);                 // A K-major (dpre), B MN-major (Ms)
      constexpr uint32_t idesc_w = idesc_bf16(128, ATT) | (1u << 15) | (1u << 16);    // A MN-major (x), B MN-major (dpre)
      mbar_wait(&bar.a_full[0], 0);
      tc_fence_after();
      issue_scores(el, tmem + T_E, smem_u32(xim), b0);
      umma_commit_e(el, &bar.e_full[0]);
      for (int it = 0; it < n_mine; ++it) {
        const int s = it % ST_NSTAGE;
        if (it + 1 < n_mine) {
          const int s1 = (it + 1) % ST_NSTAGE, a1 = (it + 1) & 1;
          if (it + 1 >= 2) mbar_wait(&bar.e_empty[a1], (((it + 1) >> 1) - 1) & 1);
          mbar_wait(&bar.a_full[s1], ((it + 1) / ST_NSTAGE) & 1);
          tc_fence_after();
          issue_scores(el, tmem + T_E + a1 * ATT, smem_u32(xim + s1 * ST_XIMG), b0);
          umma_commit_e(el, &bar.e_full[a1]);
        }
        mbar_wait(&bar.p_ready, it & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(xim + s * ST_XIMG);
        // dx part = dpre [128 x 64] · Ms [64 x 128]
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ph = smem_desc_sw128(p0) + (uint64_t)(kk * 2), pl = smem_desc_sw128(p0 + 16384) + (uint64_t)(kk * 2);
          const uint64_t mh = desc_mn(b0 + kk * 2048, 16384), ml = desc_mn(b0 + 8192 + kk * 2048, 16384);
          umma_bf16_e(el, tmem + T_DX + (it & 1) * 128, ph, mh, idesc_dx, kk != 0);
          umma_bf16_e(el, tmem + T_DX + (it & 1) * 128, ph, ml, idesc_dx, 1);
          umma_bf16_e(el, tmem + T_DX + (it & 1) * 128, pl, mh, idesc_dx, 1);
        }
        // dMs^T [128 c x 64 a] += x^T [128 c x 128 rows] · dpre [128 rows x 64 a]
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t xh = desc_mn(a0 + ks * 2048, 32768), xl = desc_mn(a0 + 16384 + ks * 2048, 32768);
          const uint64_t ph = desc_mn(p0 + ks * 2048, 8192), pl = desc_mn(p0 + 16384 + ks * 2048, 8192);
          umma_bf16_e(el, tmem + T_W, xh, ph, idesc_w, (it | ks) != 0);
          umma_bf16_e(el, tmem + T_W, xh, pl, idesc_w, 1);
          umma_bf16_e(el, tmem + T_W, xl, ph, idesc_w, 1);
        }
        umma_commit_e(el, &bar.a_empty[s]);
        umma_commit_e(el, &bar.d_full);
      }
      umma_commit_e(el, &bar.w_full);
    }
  } else {
    // ------------------------------------------------------------------ row threads
    const int q = warp & 3, r = q * 32 + lane, et = tid - 160;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float* sw = stg + q * 32 * ST_STG_LD;
    float dws0 = 0.f, dws1 = 0.f;
    // Software pipeline over the tiles: A(t) = scores -> softmax backward -> d(pre-tanh) image of tile t, B(t) = dx rows of tile t.  The
    // order is A(0), A(1), B(0), A(2), B(1), ...: the MMAs of tile t (dx, dMs^T) run while these threads are busy with A(t+1), so
    // they never wait for the tensor pipe; the dx accumulator is double-buffered for that.
    for (int it = 0; it <= n_mine; ++it) {
     if (it < n_mine) {
      const int acc = it & 1, pb = it & 1;
      mbar_wait(&bar.m_full[it % ST_NMETA], (it / ST_NMETA) & 1);      // own barrier: a_full may already be a phase ahead by now
      mbar_wait(&bar.e_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const StMeta& m = meta[it % ST_NMETA];
      const int rows = m.sbase[m.ns];
      float th[ATT];
      float sc = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(trow + T_E + acc * ATT + h * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) { th[h * 32 + i] = tanh_fast(v[i]); sc += ws_s[h * 32 + i] * th[h * 32 + i]; }
      }
      score[pb][r] = sc;
      tc_fence_before();
      mbar_arrive(&bar.e_empty[acc]);
      bar_sync(2, 128);
      // softmax over each sentence's positions and its backward, a thread per ROW: the rows of a sentence recompute its max, sum and
      // <soft, d_soft> in the same order (a thread per sentence leaves most of the 128 threads idle behind the longest sentence)
      if (!(dbg & 8)) {
        float ex = 0.f, mx = 0.f;
        int b = 0, e = 0, npad = 0;
        if (r < rows) {
          const int j = m.rsent[r];
          b = m.sbase[j]; e = m.sbase[j + 1]; npad = L - (e - b);
          mx = npad > 0 ? 0.f : -INFINITY;
          for (int rr = b; rr < e; ++rr) mx = fmaxf(mx, score[pb][rr]);
          ex = expf(score[pb][r] - mx);
          soft[pb][r] = ex;
        }
        bar_sync(2, 128);
        if (r < rows) {
          float sum = npad > 0 ? (float)npad * expf(-mx) : 0.f;
          for (int rr = b; rr < e; ++rr) sum += soft[pb][rr];
          ex *= 1.f / sum;
        }
        bar_sync(2, 128);
        if (r < rows) soft[pb][r] = ex;
        bar_sync(2, 128);
        if (r < rows) {
          float dot = 0.f;                               // padded positions: d_soft = <0, .> = 0
          for (int rr = b; rr < e; ++rr) dot += soft[pb][rr] * m.dsoft[rr];
          dsc[pb][r] = ex * (m.dsoft[r] - dot);
        }
      }
      bar_sync(2, 128);
      // d(pre-tanh) row -> bf16 hi/lo image; dWs partial sums
      const float ds = r < rows ? dsc[pb][r] : 0.f;
      if (it >= 1) { mbar_wait(&bar.d_full, (it - 1) & 1); tc_fence_after(); }      // the MMAs of tile it-1 have read the image
      {
        unsigned char* prow = pim + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int a = ch * 8 + 2 * i;
            split2(ds * ws_s[a] * (1.f - th[a] * th[a]), ds * ws_s[a + 1] * (1.f - th[a + 1] * th[a + 1]), hi[i], lo[i]);
          }
          const uint32_t off = (uint32_t)((ch ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(prow + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_async_smem();
      mbar_arrive(&bar.p_ready);
      // dWs[a] += sum_rows d_score th[a]: transposing butterfly over the warp's 32 rows; lane ends up with a = 2*lane + {0,1}
#pragma unroll
      for (int i = 0; i < ATT; ++i) th[i] *= ds;
#pragma unroll
      for (int w = 32, o = 16; o > 0 && !(dbg & 16); w >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
          const float keep = up ? th[i + w] : th[i];
          const float send = up ? th[i] : th[i + w];
          th[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      dws0 += th[0];
      dws1 += th[1];
     }
     if (it >= 1) {
      const int jt = it - 1, pb = jt & 1;
      const StMeta& m = meta[jt % ST_NMETA];
      const int rows = m.sbase[m.ns];
      if (it == n_mine) { mbar_wait(&bar.d_full, jt & 1); tc_fence_after(); }      // (otherwise waited for in A(it) above)
      // dx rows: soft · d_self_atte + dpre · Ms.  The 8 rows a thread stores (row j*4 + lane/8 of the warp's 32, 4 channels at
      // (lane%8)*4) are resolved once per tile - output pointer, d_self_atte row of the sentence, soft - and the d_self_atte values of a
      // 32-channel chunk are requested one chunk ahead (the first before the accumulator is ready)
      const int s0 = m.s0;
      const int c4 = (lane & 7) * 4;
      float* orow[8];
      const float* drow[8];
      float so[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = q * 32 + j * 4 + (lane >> 3);
        orow[j] = nullptr; drow[j] = d_sa; so[j] = 0.f;
        if (row < rows) {
          orow[j] = dx + (size_t)m.rowmap[row] * D + c4;
          drow[j] = d_sa + (size_t)(s0 + m.rsent[row]) * D + c4;
          so[j] = soft[pb][row];
        }
      }
      float4 dn[8];
      auto fetch_d = [&](int c0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dn[j] = orow[j] ? *reinterpret_cast<const float4*>(drow[j] + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      fetch_d(0);
#pragma unroll 1
      for (int c0 = 0; c0 < D && !(dbg & 1); c0 += 32) {
        float v[32];
        tmem_ld32(trow + T_DX + pb * 128 + c0, v);
        float4 dc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dc[j] = dn[j];
        if (c0 + 32 < D) fetch_d(c0 + 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(&sw[lane * ST_STG_LD + j * 4]) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (orow[j]) {
            float4 o = *reinterpret_cast<const float4*>(&sw[(j * 4 + (lane >> 3)) * ST_STG_LD + c4]);
            o.x += so[j] * dc[j].x; o.y += so[j] * dc[j].y; o.z += so[j] * dc[j].z; o.w += so[j] * dc[j].w;
            *reinterpret_cast<float4*>(orow[j] + c0) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
     }
    }
    if (n_mine > 0) {
      // flush: dMs[a][c] += W^T[c][a] (TMEM lane = c), dWs
      mbar_wait(&bar.w_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(trow + T_W + h * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(&dMs[(h * 32 + i) * D + r], v[i]);
      }
      const int a = 2 * lane;
      atomicAdd(&dWs[a], dws0);
      atomicAdd(&dWs[a + 1], dws1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace umpr

using namespace umpr;

static int snet_tc_check(const char* what, const void* x, const int* table, int N, int L, int n_tiles) {
  if (L < 1 || L > 128) return fail_arg("%s: sentence length L=%d must be in [1, 128]", what, L);
  if (n_tiles < 1 || n_tiles > N || !table) return fail_arg("%s: tile table missing or inconsistent (n_tiles=%d, N=%d)", what, n_tiles, N);
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail_arg("%s: x must be 16-byte aligned", what);
  return 0;
}

// table = [tile_sent_off (n_tiles+1) | cstart (N+1)]: sentences tile_sent_off[k] .. tile_sent_off[k+1]-1 form tile k, cstart is the
// exclusive prefix sum of the sentence lengths (plan.py:snet_table); every tile holds at most 128 valid rows.
extern "C" int umpr_snet_fwd_tc(const float* x, const int* table, int n_tiles, const float* Ms, const float* Ws, int N, int L,
                                float* self_atte, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (int e = snet_tc_check("snet_fwd_tc", x, table, N, L, n_tiles)) return e;
  constexpr int smem = ST_MSIMG + ST_NSTAGE * ST_XIMG + ST_NMETA * (int)sizeof(StMeta) + 1024;
  cudaError_t e = cudaFuncSetAttribute(snet_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("snet_fwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  snet_fwd_tc_kernel<<<grid, SF_THREADS, smem, (cudaStream_t)stream>>>(x, table, table + n_tiles + 1, Ms, Ws, n_tiles, L, self_atte, dbg_flags());
  return check_launch("snet_fwd_tc");
}

// dx is written for rows below each sentence's length only; dMs / dWs are accumulated (+=)
extern "C" int umpr_snet_bwd_tc(const float* x, const int* table, int n_tiles, const float* d_sa, const float* Ms, const float* Ws,
                                int N, int L, float* dx, float* dMs, float* dWs, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (int e = snet_tc_check("snet_bwd_tc", x, table, N, L, n_tiles)) return e;
  constexpr int smem = ST_MSIMG + ST_PIMG + ST_NSTAGE * ST_XIMG + 4 * 32 * ST_STG_LD * 4 + ST_NMETA * (int)sizeof(StMeta) + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(snet_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("snet_bwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  snet_bwd_tc_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(x, table, table + n_tiles + 1, d_sa, Ms, Ws, n_tiles, L, dx, dMs, dWs, dbg_flags());
  return check_launch("snet_bwd_tc");
}
