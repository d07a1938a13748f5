// C-Net convolution on the tensor cores (reference src/model.py:118-120): Conv1d(128 -> K, k=3, pad=1) + ReLU + max over L
// as an implicit GEMM, persistent CTAs, 3xBF16 split, accumulators in TMEM.
//   tile   = whole sentences, each with one zero guard row before and after.  Without a length table: gs sentences of L+2 rows.
//            With one (inputs from ImprovedRnn: rows at or beyond a sentence's length are exactly zero, model.py:20): only the
//            len valid rows plus the two guards are laid out (positions 0..len are computed; every later position is an all-zero
//            window whose value is exactly the bias and enters the max analytically), so a tile holds ~2.4x more sentences.
//   D      = W · x^T: M = filters (TMEM lanes), N = the 128 positions of a tile (TMEM columns)
//   K      = 3*128 = 6 blocks of 64: block kb covers tap dt = kb/2, channels (kb%2)*64..+64
//   A      = the weights, pre-split once per step into the exact shared-memory image (bf16 hi/lo, SWIZZLE_128B) by
//            cnet_tc_prep_kernel and streamed through a ring by ONE cp.async.bulk per k-block (TMA bulk copy), each stage used
//            for two tiles
//   B      = ONE x image per tile (every x row read and split once); tap dt reads it one row further down (descriptor start
//            address + dt*128 B) - the im2col is descriptor arithmetic, nothing is materialised
//   epilogue: thread = filter, tcgen05.ld of its own lane -> bias, max / arg-max over each sentence's columns -> cfeat, cidx
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int CT_TILE = 128 * 128;       // bytes of one [128][64 bf16] swizzled tile
constexpr int CT_KB = 6;
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

constexpr int CT_NMETA = 6;              // three tile pairs: the loaders of pair p reuse the slots of pair p-3, whose epilogue has finished
constexpr int CT_MAXS = 44;              // sentences per tile: each takes at least 3 rows
constexpr int CT_XSUB = 17 * 1024;       // one [136 rows][64 bf16] swizzled x sub-image (130 rows used)
constexpr int CT_XIMG = 4 * CT_XSUB;     // [channel half][hi|lo]
constexpr int CT_WSTAGE = 2 * CT_TILE;   // one k-block of the weights, hi + lo
constexpr int CT_NW = 2;                 // weight ring stages
struct CtMeta {
  int s0, ns;
  int sb[CT_MAXS];                       // tile row of each sentence's leading guard row
  int len[CT_MAXS];                      // valid rows of each sentence
  int rowsrc[128];                       // tile row -> global x row, -1 = zero row
};

// Roles (576 threads): warps 0-7 loaders, warp 8 weight producer (TMA bulk copies), warp 9 MMA issuer, warps 10-13 / 14-17 the
// epilogue of the first / second tile of a pair.
//   * Every x row is read ONCE per tile: the loaders write one image [channel half][hi|lo][130 rows][64 bf16] (image row = tile row
//     + 1, zero rows for guards) and the three taps are the SAME image read through descriptors whose start address is shifted by
//     0 / 128 / 256 bytes (one row; the 128-byte swizzle is a function of the shared-memory address, so a row shift keeps it intact).
//   * The weights stream through a 2-stage ring, one k-block (tap, channel half) per stage, and every stage is used for TWO tiles.
//   * D = W · x^T: filters are the TMEM lanes, positions the columns, so the max over a sentence's positions is a scan by ONE
//     thread over its own lane - no staging, no barriers in the epilogue; thread f writes cfeat / cidx of filter f.
constexpr int CT2_THREADS = 576;

__global__ void __launch_bounds__(CT2_THREADS, 1) cnet_conv_fwd_tc_kernel(const float* __restrict__ x, const unsigned char* __restrict__ wimg,
                                                                          const float* __restrict__ bias, int N, int L, int KC, int gs,
                                                                          const int* __restrict__ tso, const int* __restrict__ cstc, int n_tiles,
                                                                          float* __restrict__ cfeat, int* __restrict__ cidx,
                                                                          int4* __restrict__ worklist, int cap) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t a_full[2], a_empty, w_full[CT_NW], w_empty[CT_NW], acc_full[2], acc_empty[2], m_full[CT_NMETA];
  __shared__ uint32_t tmem_slot;
  __shared__ CtMeta meta[CT_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* ximg = base;                              // [tile of the pair 2][CT_XIMG]
  unsigned char* wring = base + 2 * CT_XIMG;               // [CT_NW][CT_WSTAGE]
  const float* wnorm = reinterpret_cast<const float*>(wimg + CT_IMG_BYTES);
  int* counter = reinterpret_cast<int*>(const_cast<unsigned char*>(wimg) + CT_IMG_BYTES + 512);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lg = L + 2;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;
  const int n_pairs = (n_mine + 1) >> 1;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 256); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 256); }
    mbar_init(&a_empty, 1);
    for (int s = 0; s < CT_NW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < CT_NMETA; ++s) mbar_init(&m_full[s], 256);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders
    for (int p = 0; p < n_pairs; ++p) {
      if (p >= 1) mbar_wait(&a_empty, (p - 1) & 1);        // the MMAs of the previous pair have read the images
      for (int j = 0; j < 2; ++j) {
        const int t = 2 * p + j;
        if (t >= n_mine) break;
        const int tile = blockIdx.x + t * gridDim.x;
        CtMeta& m = meta[t % CT_NMETA];
        if (tid < 128) m.rowsrc[tid] = -1;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        {
          int s0, ns;
          if (tso) { s0 = tso[tile]; ns = tso[tile + 1] - s0; } else { s0 = tile * gs; ns = min(gs, N - s0); }
          if (tid < ns) {
            int rb, len;
            if (tso) { const int c0 = cstc[s0]; rb = cstc[s0 + tid] - c0; len = cstc[s0 + tid + 1] - cstc[s0 + tid] - 2; }
            else { rb = tid * Lg; len = L; }
            m.sb[tid] = rb; m.len[tid] = len;
            const int g0 = (s0 + tid) * L;
            for (int l = 0; l < len; ++l) m.rowsrc[rb + 1 + l] = g0 + l;
          }
          if (tid == 0) { m.s0 = s0; m.ns = ns; }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        mbar_arrive(&m_full[t % CT_NMETA]);
        // image rows 0..135 (row r = tile row r-1): a warp per row and pass, 32 lanes x 4 channels; all 136 rows are (re)written
        unsigned char* img = ximg + j * CT_XIMG;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int i0 = h ? 9 : 0, ni = h ? 8 : 9;
          float4 va[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int r = (i0 + i) * 8 + warp;
            const int tr = r - 1;
            const int src = (i < ni && tr >= 0 && tr < 128) ? m.rowsrc[tr] : -1;
            va[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + (size_t)src * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            if (i < ni) {
              const int r = (i0 + i) * 8 + warp;
              unsigned char* sub = img + (lane >> 4) * 2 * CT_XSUB;
              store_split4(sub, sub + CT_XSUB, r, (lane & 15) * 4, va[i]);
            }
          }
        }
        fence_async_smem();
        mbar_arrive(&a_full[j]);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ weight producer: one 32 KB bulk copy per k-block and pair
    const uint32_t el = elect_one_sync();
    const int n_it = n_pairs * CT_KB;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % CT_NW;
      if (it >= CT_NW) mbar_wait(&w_empty[s], ((it / CT_NW) - 1) & 1);
      mbar_arrive_expect_tx_e(el, &w_full[s], CT_WSTAGE);
      bulk_copy_g2s_e(el, wring + s * CT_WSTAGE, wimg + (size_t)(it % CT_KB) * CT_WSTAGE, CT_WSTAGE, &w_full[s]);
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, elected lane issues)
    const uint32_t el = elect_one_sync();
    constexpr uint32_t idesc = idesc_bf16(128, 128);
    int wit = 0;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1;
      const bool two = 2 * p + 1 < n_mine;
      if (p >= 2) mbar_wait(&acc_empty[buf], ((p >> 1) - 1) & 1);
#pragma unroll 1
      for (int kb = 0; kb < CT_KB; ++kb, ++wit) {
        const int s = wit % CT_NW;
        mbar_wait(&w_full[s], (wit / CT_NW) & 1);
        if (kb == 0) {
          mbar_wait(&a_full[0], p & 1);
          if (two) mbar_wait(&a_full[1], p & 1);
        }
        tc_fence_after();
        const int dt = kb >> 1, ch = kb & 1;
        const uint32_t wst = smem_u32(wring + s * CT_WSTAGE);
        const uint64_t wh = smem_desc_sw128(wst), wl = smem_desc_sw128(wst + CT_TILE);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j == 1 && !two) break;
          const uint32_t xa = smem_u32(ximg + j * CT_XIMG + ch * 2 * CT_XSUB) + dt * 128;      // tap dt = the image one row further
          const uint64_t xh = smem_desc_sw128(xa), xl = smem_desc_sw128(xa + CT_XSUB);
          const uint32_t d = tmem + buf * 256 + j * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            umma_bf16_e(el, d, wh + o, xh + o, idesc, (kb | kk) != 0);
            umma_bf16_e(el, d, wh + o, xl + o, idesc, 1);
            umma_bf16_e(el, d, wl + o, xh + o, idesc, 1);
          }
        }
        umma_commit_e(el, &w_empty[s]);
      }
      umma_commit_e(el, &a_empty);
      umma_commit_e(el, &acc_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group g takes tile g of every pair; thread = filter
    const int g = (warp - 10) >> 2, q = warp & 3, f = q * 32 + lane;
    const bool act = f < KC;
    const float bz = act ? bias[f] : 0.f;
    const float tau = act ? CT_TAU * wnorm[f] : 0.f;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1, t = 2 * p + g;
      // (also when this group has no tile in the last pair: its arrival below must not run ahead into an earlier phase of acc_empty)
      mbar_wait(&acc_full[buf], (p >> 1) & 1);
      if (t < n_mine) {
        mbar_wait(&m_full[t % CT_NMETA], (t / CT_NMETA) & 1);
        tc_fence_after();
        const CtMeta& m = meta[t % CT_NMETA];
        const int n0 = m.s0, ns = m.ns;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + g * 128;
#pragma unroll 1
        for (int sn = 0; sn < ns; ++sn) {
          // max over the L positions of the sentence (model.py:120); first maximum wins, <= 0 carries no gradient
          const int sb = m.sb[sn], len = m.len[sn];
          const int np = min(len + 1, L);              // positions whose window touches a valid row
          float best = -INFINITY, second = -INFINITY;
          int arg = -1, arg2 = -1;
          const int first = sb + 1;                    // accumulator columns [first, first + np) in 16-column aligned chunks
#pragma unroll 1
          for (int c0 = first & ~15; c0 < first + np; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int l = c0 + i - first;
              if (l >= 0 && l < np) {
                const float y = v[i] + bz;             // pre-activation; ReLU is applied after the max (monotone)
                if (y > best) { second = best; arg2 = arg; best = y; arg = l; }
                else if (y > second && !(y == best && y == bz)) { second = y; arg2 = l; }   // all-zero windows give exactly the bias: a true tie, first wins
              }
            }
          }
          if (np < L) {                                // the remaining positions are all-zero windows: exactly the bias, first one at np
            if (bz > best) { second = best; arg2 = arg; best = bz; arg = np; }
            else if (bz > second && !(bz == best)) { second = bz; arg2 = np; }
          }
          if (act) {
            const size_t o = (size_t)(n0 + sn) * KC + f;
            cfeat[o] = fmaxf(best, 0.f);
            cidx[o] = best > 0.f ? arg : -1;
            // the arg-max routes the gradient: near-ties (and maxima next to the ReLU threshold) are re-scored exactly in fp32
            if ((arg2 >= 0 && best - second <= tau && best > -tau) || fabsf(best) <= tau) {
              const int slot = atomicAdd(counter, 1);
              if (slot < cap) worklist[slot] = make_int4(n0 + sn, f, arg, (arg2 >= 0 && best - second <= tau) ? arg2 : -1);
            }
          }
        }
        tc_fence_before();
      }
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
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
  const int smem = 2 * CT_XIMG + CT_NW * CT_WSTAGE + 1024;
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_fwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_fwd_tc_kernel<<<grid, CT2_THREADS, smem, (cudaStream_t)stream>>>(x, reinterpret_cast<const unsigned char*>(wimg), conv_b, N, L, KC,
                                                                          gs, table, table ? table + n_tiles + 1 : nullptr, n_tiles, cfeat, cidx, reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(wimg) + CT_HDR_BYTES), cap);
  if (int rc = check_launch("cnet_conv_fwd_tc")) return rc;
  cnet_conv_fix_kernel<<<n_ctas * 2, 256, 0, (cudaStream_t)stream>>>(x, conv_w, conv_b, reinterpret_cast<const int*>(reinterpret_cast<unsigned char*>(wimg) + CT_IMG_BYTES + 512),
                                                                      reinterpret_cast<const int4*>(reinterpret_cast<unsigned char*>(wimg) + CT_HDR_BYTES), cap, L, KC, cfeat, cidx);
  return check_launch("cnet_conv_fix");
}
