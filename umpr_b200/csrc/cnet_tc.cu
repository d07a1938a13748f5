// C-Net convolution on the tensor cores (reference src/model.py:118-120): Conv1d(128 -> K, k=3, pad=1) + ReLU + max over L
// as an implicit GEMM, persistent CTAs, 3xBF16 split, accumulators in TMEM.
//   M tile = whole sentences, each with one zero guard row before and after.  Without a length table: gs sentences of L+2 rows.
//            With one (inputs from ImprovedRnn: rows at or beyond a sentence's length are exactly zero, model.py:20): only the
//            len valid rows plus the two guards are laid out (positions 0..len are computed; every later position is an all-zero
//            window whose value is exactly the bias and enters the max analytically), so a tile holds ~2.4x more sentences.
//   K      = 3*128 = 6 blocks of 64: block kb covers tap dt = kb/2, channels (kb%2)*64..+64; the A rows of tap dt are the
//            x rows shifted by dt-1 (the im2col happens in the loader's addressing, nothing is materialised)
//   B      = the weights, pre-split once per step into the exact shared-memory image (bf16 hi/lo, SWIZZLE_128B) by
//            cnet_tc_prep_kernel and streamed into the stage ring by ONE cp.async.bulk per k-block (TMA bulk copy)
//   warps 0-7 loaders (A tiles: global fp32 -> bf16 hi/lo -> swizzled smem), warp 8 MMA issuer, warps 9-12 epilogue
//   (tcgen05.ld -> bias + ReLU -> shared staging -> max / arg-max over each sentence's rows -> cfeat, cidx)
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int CT_THREADS = 416;          // 13 warps
constexpr int CT_NSTAGE = 3;
constexpr int CT_TILE = 128 * 128;       // bytes of one [128][64 bf16] swizzled tile
constexpr int CT_STAGE = 4 * CT_TILE;    // A_hi, A_lo, B_hi, B_lo
constexpr int CT_KB = 6;
constexpr int CT_STG_LD = 33;
constexpr int CT_IMG_BYTES = CT_KB * 2 * CT_TILE;     // 196608
constexpr int CT_HDR_BYTES = CT_IMG_BYTES + 1024;      // image | wnorm[128] floats | counter | pad   then int4 worklist[cap]
// |3xBF16 dot - exact| <= 1.2e-5 |a||b|; two values can swap order if closer than twice that; |window| <= sqrt(384) since |x| < 1
constexpr float CT_TAU = 2.f * 1.2e-5f * 19.6f;

// wimg[kb][hi|lo][128 n][128 B]: W[n][c][dt] with dt = kb/2, c = (kb%2)*64 + k
__global__ void cnet_tc_prep_kernel(const float* __restrict__ w, int KC, unsigned char* __restrict__ wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (kb, n, k4)
  if (idx >= CT_KB * 128 * 16) return;
  const int kb = idx / (128 * 16), n = (idx >> 4) & 127, k = (idx & 15) * 4;
  const int dt = kb >> 1, c = (kb & 1) * 64 + k;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < KC) {
#pragma unroll
    for (int q = 0; q < 4; ++q) t[q] = w[((size_t)n * D + c + q) * 3 + dt];
  }
  unsigned char* img = wimg + (size_t)kb * 2 * CT_TILE;
  store_split4(img, img + CT_TILE, n, k, make_float4(t[0], t[1], t[2], t[3]));
  // |W_kf| for the near-tie tolerance: one warp per filter; thread 0 also resets the worklist counter
  const int f = idx >> 5, lane = idx & 31;
  if (f < 128) {
    float a = 0.f;
    if (f < KC) for (int e = lane; e < D * 3; e += 32) { const float v = w[(size_t)f * D * 3 + e]; a += v * v; }
    a = warp_sum(a);
    if (lane == 0) reinterpret_cast<float*>(wimg + CT_IMG_BYTES)[f] = sqrtf(a);
  }
  if (idx == 0) *reinterpret_cast<int*>(wimg + CT_IMG_BYTES + 512) = 0;
}

constexpr int CT_NMETA = 4;
constexpr int CT_MAXS = 44;              // sentences per tile: each takes at least 3 rows
struct CtMeta {
  int s0, ns;
  int sb[CT_MAXS];                       // tile row of each sentence's leading guard row
  int len[CT_MAXS];                      // valid rows of each sentence
  int rowsrc[128];                       // tile row -> global x row, -1 = zero row
};

__global__ void __launch_bounds__(CT_THREADS, 1) cnet_conv_fwd_tc_kernel(const float* __restrict__ x, const unsigned char* __restrict__ wimg,
                                                                         const float* __restrict__ bias, int N, int L, int KC, int gs,
                                                                         const int* __restrict__ tso, const int* __restrict__ cstc, int n_tiles,
                                                                         float* __restrict__ cfeat, int* __restrict__ cidx,
                                                                         int4* __restrict__ worklist, int cap) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t full_bar[CT_NSTAGE], empty_bar[CT_NSTAGE], acc_full[2], acc_empty[2], m_full[CT_NMETA];
  __shared__ uint32_t tmem_slot;
  __shared__ CtMeta meta[CT_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* stg = reinterpret_cast<float*>(base + CT_NSTAGE * CT_STAGE);      // [128][33]
  const float* wnorm = reinterpret_cast<const float*>(wimg + CT_IMG_BYTES);
  int* counter = reinterpret_cast<int*>(const_cast<unsigned char*>(wimg) + CT_IMG_BYTES + 512);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lg = L + 2;

  if (tid == 0) {
    for (int s = 0; s < CT_NSTAGE; ++s) { mbar_init(&full_bar[s], 256); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    for (int s = 0; s < CT_NMETA; ++s) mbar_init(&m_full[s], 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders
    int it = 0;       // global k-block counter (ring position)
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      // tile bookkeeping (slot t % 4: the epilogue of tile t-4 has finished before the ring lets this tile's stages through)
      CtMeta& m = meta[t % CT_NMETA];
      if (t >= CT_NMETA) mbar_wait(&empty_bar[it % CT_NSTAGE], ((it / CT_NSTAGE) - 1) & 1);
      if (tid < 128) m.rowsrc[tid] = -1;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      {
        int s0, ns;
        if (tso) { s0 = tso[tile]; ns = tso[tile + 1] - s0; } else { s0 = tile * gs; ns = min(gs, N - s0); }
        if (tid < ns) {
          int base, len;
          if (tso) { const int c0 = cstc[s0]; base = cstc[s0 + tid] - c0; len = cstc[s0 + tid + 1] - cstc[s0 + tid] - 2; }
          else { base = tid * Lg; len = L; }
          m.sb[tid] = base; m.len[tid] = len;
          const int g0 = (s0 + tid) * L;
          for (int l = 0; l < len; ++l) m.rowsrc[base + 1 + l] = g0 + l;
        }
        if (tid == 0) { m.s0 = s0; m.ns = ns; }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      mbar_arrive(&m_full[t % CT_NMETA]);
#pragma unroll 1
      for (int kb = 0; kb < CT_KB; ++kb, ++it) {
        const int s = it % CT_NSTAGE;
        if (it >= CT_NSTAGE) mbar_wait(&empty_bar[s], ((it / CT_NSTAGE) - 1) & 1);
        unsigned char* st = base + s * CT_STAGE;
        if (tid == 0) {      // weights: one bulk copy of the pre-split image (hi + lo = 32 KB) per k-block
          mbar_expect_tx(&full_bar[s], 2 * CT_TILE);
          bulk_copy_g2s(st + 2 * CT_TILE, wimg + (size_t)kb * 2 * CT_TILE, 2 * CT_TILE, &full_bar[s]);
        }
        const int dt = kb >> 1, c0 = (kb & 1) * 64;
        float4 va[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 256 + tid, r = idx >> 4, k = (idx & 15) * 4;
          const int rr = r - 1 + dt;                 // tile row whose x feeds output row r at tap dt
          const int src = (rr >= 0 && rr < 128) ? m.rowsrc[rr] : -1;
          va[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + (size_t)src * D + c0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 256 + tid;
          store_split4(st, st + CT_TILE, idx >> 4, (idx & 15) * 4, va[i]);
        }
        fence_async_smem();
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, 128);
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        if (t >= 2) mbar_wait(&acc_empty[acc], ((t >> 1) - 1) & 1);
        const uint32_t d = tmem + acc * 128;
#pragma unroll 1
        for (int kb = 0; kb < CT_KB; ++kb, ++it) {
          const int s = it % CT_NSTAGE;
          mbar_wait(&full_bar[s], (it / CT_NSTAGE) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(base + s * CT_STAGE);
          const uint64_t ah = smem_desc_sw128(st), al = smem_desc_sw128(st + CT_TILE);
          const uint64_t bh = smem_desc_sw128(st + 2 * CT_TILE), bl = smem_desc_sw128(st + 3 * CT_TILE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            umma_bf16(d, ah + o, bh + o, idesc, (kb | kk) != 0);
            umma_bf16(d, ah + o, bl + o, idesc, 1);
            umma_bf16(d, al + o, bh + o, idesc, 1);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 9..12 -> TMEM lane quarter warp%4)
    const int q = warp & 3, row = q * 32 + lane, etid = (warp - 9) * 32 + lane;
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      mbar_wait(&m_full[t % CT_NMETA], (t / CT_NMETA) & 1);
      const CtMeta& m = meta[t % CT_NMETA];
      const int n0 = m.s0, ns = m.ns;
      mbar_wait(&acc_full[acc], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * 128 + c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float bv = c0 + c < KC ? bias[c0 + c] : 0.f;
          stg[row * CT_STG_LD + c] = v[c] + bv;        // pre-activation; ReLU is applied after the max (monotone)
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // max over the L positions of each sentence (model.py:120); first maximum wins, <= 0 carries no gradient
        for (int idx = etid; idx < ns * 32; idx += 128) {
          const int sn = idx >> 5, c = idx & 31;
          if (c0 + c < KC) {
            float best = -INFINITY, second = -INFINITY;
            int arg = -1, arg2 = -1;
            const float bz = bias[c0 + c];
            const int len = m.len[sn];
            const int np = min(len + 1, L);            // positions whose window touches a valid row
            const float* col = stg + (m.sb[sn] + 1) * CT_STG_LD + c;
            for (int l = 0; l < np; ++l) {
              const float y = col[l * CT_STG_LD];
              if (y > best) { second = best; arg2 = arg; best = y; arg = l; }
              else if (y > second && !(y == best && y == bz)) { second = y; arg2 = l; }   // all-zero windows give exactly the bias: a true tie, first wins
            }
            if (np < L) {                              // the remaining positions are all-zero windows: exactly the bias, first one at np
              if (bz > best) { second = best; arg2 = arg; best = bz; arg = np; }
              else if (bz > second && !(bz == best)) { second = bz; arg2 = np; }
            }
            const size_t o = (size_t)(n0 + sn) * KC + c0 + c;
            cfeat[o] = fmaxf(best, 0.f);
            cidx[o] = best > 0.f ? arg : -1;
            // the arg-max routes the gradient: near-ties (and maxima next to the ReLU threshold) are re-scored exactly in fp32
            const float tau = CT_TAU * wnorm[c0 + c];
            if ((arg2 >= 0 && best - second <= tau && best > -tau) || fabsf(best) <= tau) {
              const int slot = atomicAdd(counter, 1);
              if (slot < cap) worklist[slot] = make_int4(n0 + sn, c0 + c, arg, (arg2 >= 0 && best - second <= tau) ? arg2 : -1);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// exact fp32 re-scoring of the uncertain (sentence, filter) pairs: one warp per record
__global__ void __launch_bounds__(256) cnet_conv_fix_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            const int* __restrict__ counter, const int4* __restrict__ worklist, int cap, int L,
                                                            int KC, float* __restrict__ cfeat, int* __restrict__ cidx) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int n_rec = min(*counter, cap);
  for (int rec = wid; rec < n_rec; rec += nw) {
    const int4 r = worklist[rec];
    const int n = r.x, kf = r.y;
    float best = 0.f;
    int arg = -1;
    for (int cand = 0; cand < 2; ++cand) {
      const int l = cand ? r.w : r.z;
      if (l < 0) continue;
      float a = 0.f;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ll = l + dt - 1;
        if (ll < 0 || ll >= L) continue;
        const float4 xv = *reinterpret_cast<const float4*>(x + ((size_t)n * L + ll) * D + lane * 4);
        const float* wp = w + ((size_t)kf * D + lane * 4) * 3 + dt;
        a += xv.x * wp[0] + xv.y * wp[3] + xv.z * wp[6] + xv.w * wp[9];
      }
      a = warp_sum(a) + bias[kf];
      a = fmaxf(a, 0.f);
      if (arg < 0 || a > best || (a == best && l < arg)) { best = a; arg = l; }
    }
    if (lane == 0) {
      const size_t o = (size_t)n * KC + kf;
      cfeat[o] = best;
      cidx[o] = best > 0.f ? arg : -1;
    }
  }
}

}  // namespace umpr

using namespace umpr;

// scratch: 197632 + 16*cap bytes, 16-byte aligned (weight image, filter norms, worklist of near-tied maxima).
// table (optional) = [tile_sent_off (n_tiles+1) | cstart (N+1)] with cstart the exclusive prefix sum of (len + 2) per sentence
// (plan.py:cnet_table): sentences tile_sent_off[k] .. tile_sent_off[k+1]-1 form tile k, at most 128 rows including the guards.
extern "C" int umpr_cnet_conv_fwd_tc(const float* x, const float* conv_w, const float* conv_b, int N, int L, int KC, int ksize,
                                     const int32_t* table, int table_tiles, void* wimg, int cap, float* cfeat, int32_t* cidx,
                                     int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (ksize != 3) return fail_arg("cnet: kernel_size=%d (only 3 is built)", ksize);
  if (KC < 1 || KC > 128) return fail_arg("cnet: kernel_count=%d must be in [1, 128]", KC);
  if (L < 1 || L + 2 > 128) return fail_arg("cnet_conv_fwd: sentence length L=%d must be in [1, 126]", L);
  if (cap < 1) return fail_arg("cnet_conv_fwd_tc: cap=%d", cap);
  if (reinterpret_cast<uintptr_t>(wimg) & 15) return fail_arg("cnet_conv_fwd_tc: wimg must be 16-byte aligned");
  cnet_tc_prep_kernel<<<(CT_KB * 128 * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conv_w, KC, reinterpret_cast<unsigned char*>(wimg));
  if (int e = check_launch("cnet_tc_prep")) return e;
  int gs = 128 / (L + 2);
  if (gs > 16) gs = 16;
  if (table && (table_tiles < 1 || table_tiles > N)) return fail_arg("cnet_conv_fwd_tc: tile table inconsistent (n_tiles=%d, N=%d)", table_tiles, N);
  const int n_tiles = table ? table_tiles : (N + gs - 1) / gs;
  const int smem = CT_NSTAGE * CT_STAGE + 128 * CT_STG_LD * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_fwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_fwd_tc_kernel<<<grid, CT_THREADS, smem, (cudaStream_t)stream>>>(x, reinterpret_cast<const unsigned char*>(wimg), conv_b, N, L, KC,
                                                                          gs, table, table ? table + n_tiles + 1 : nullptr, n_tiles, cfeat, cidx, reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(wimg) + CT_HDR_BYTES), cap);
  if (int rc = check_launch("cnet_conv_fwd_tc")) return rc;
  cnet_conv_fix_kernel<<<n_ctas * 2, 256, 0, (cudaStream_t)stream>>>(x, conv_w, conv_b, reinterpret_cast<const int*>(reinterpret_cast<unsigned char*>(wimg) + CT_IMG_BYTES + 512),
                                                                      reinterpret_cast<const int4*>(reinterpret_cast<unsigned char*>(wimg) + CT_HDR_BYTES), cap, L, KC, cfeat, cidx);
  return check_launch("cnet_conv_fix");
}
