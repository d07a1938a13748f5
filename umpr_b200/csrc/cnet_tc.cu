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
#include <stdlib.h>
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int CT_TILE = 128 * 128;       // bytes of one [128][64 bf16] swizzled tile
constexpr int CT_KB = 6;
constexpr int CT_IMG_BYTES = CT_KB * 2 * CT_TILE;     // 196608
constexpr int CT_HDR_BYTES = CT_IMG_BYTES + 1024;      // image | wnorm[128] floats | pad   then the 2-byte re-scoring records (N*KC)
// |3xBF16 dot - exact| <= 1.2e-5 |a||b|; two values can swap order if closer than twice that; |window| <= sqrt(384) since |x| < 1
constexpr float CT_TAU = 2.f * 1.2e-5f * 19.6f;

// wimg[kb][hi|lo][128 n][128 B]: W[n][c][dt] with dt = kb%3, c = (kb/3)*64 + k
__global__ void cnet_tc_prep_kernel(const float* __restrict__ w, int KC, unsigned char* __restrict__ wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (kb, n, k4)
  if (idx >= CT_KB * 128 * 16) return;
  const int kb = idx / (128 * 16), n = (idx >> 4) & 127, k = (idx & 15) * 4;
  const int dt = kb % 3, c = (kb / 3) * 64 + k;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < KC) {
#pragma unroll
    for (int q = 0; q < 4; ++q) t[q] = w[((size_t)n * D + c + q) * 3 + dt];
  }
  unsigned char* img = wimg + (size_t)kb * 2 * CT_TILE;
  store_split4(img, img + CT_TILE, n, k, make_float4(t[0], t[1], t[2], t[3]));
  // |W_kf| for the near-tie tolerance: one warp per filter
  const int f = idx >> 5, lane = idx & 31;
  if (f < 128) {
    float a = 0.f;
    if (f < KC) for (int e = lane; e < D * 3; e += 32) { const float v = w[(size_t)f * D * 3 + e]; a += v * v; }
    a = warp_sum(a);
    if (lane == 0) reinterpret_cast<float*>(wimg + CT_IMG_BYTES)[f] = sqrtf(a);
  }
}

constexpr int CT_NMETA = 8;              // four tile pairs: the loaders build pair p+1's bookkeeping while pair p is in flight; its slots
                                         // are those of pair p-3, whose epilogue has finished by then
constexpr int CT_MAXS = 44;              // sentences per tile: each takes at least 3 rows
constexpr int CT_XSUB = 33 * 1024;       // one [264 rows][64 bf16] swizzled x sub-image of a tile PAIR (258 rows used)
constexpr int CT_XIMG = 4 * CT_XSUB;     // [channel half][hi|lo]
constexpr int CT_WSTAGE = 2 * CT_TILE;   // one k-block of the weights, hi + lo
constexpr int CT_NW = 2;                 // weight ring stages
struct CtMeta {
  int s0, ns;
  int sb[CT_MAXS];                       // tile row of each sentence's leading guard row
  int len[CT_MAXS];                      // valid rows of each sentence
  int rowsrc[128];                       // tile row -> global x row, -1 = zero row
};

// Roles (576 threads): warps 0-3 / 4-7 loaders of the first / second tile of a pair, warp 8 weight producer (TMA bulk copies),
// warp 9 MMA issuer, warps 10-13 / 14-17 the epilogue of the even / odd sentences of a tile.
//   * Every x row is read ONCE: the loaders write one image [channel half][hi|lo][258 rows][64 bf16] per tile PAIR (image row =
//     128*j + tile row + 1; rows 128 and 129 - the last tile row of the first tile and the row before the second tile - are zero for
//     both) and the three taps are the SAME image read through descriptors whose start address is shifted by 0 / 128 / 256 bytes
//     (one row; the 128-byte swizzle is a function of the shared-memory address, so a row shift keeps it intact).
//   * The weights stream through a 2-stage ring, one k-block (channel half, tap) per stage; one N=256 MMA covers both tiles.
//   * K runs over the first channel half (3 taps), then the second: the two half-images are a double buffer - while the MMAs of a pair
//     work on the second half the loaders already store the first half of the NEXT pair, whose row loads were issued (into registers)
//     a half earlier, so neither the load latency nor the fp32 -> bf16 hi/lo split is on the tensor pipe's critical path.
//   * D = W · x^T: filters are the TMEM lanes, positions the columns, so the max over a sentence's positions is a scan by ONE
//     thread over its own lane - no staging, no barriers in the epilogue; thread f writes cfeat / cidx of filter f.
constexpr int CT2_THREADS = 576;
constexpr int CT_LROWS = 17;             // loads per loader thread and pass (one channel half of the tile's 136 image rows, 4 warps x 2 rows)

__global__ void __launch_bounds__(CT2_THREADS, 1) cnet_conv_fwd_tc_kernel(const float* __restrict__ x, const unsigned char* __restrict__ wimg,
                                                                          const float* __restrict__ bias, int N, int L, int KC, int gs,
                                                                          const int* __restrict__ tso, const int* __restrict__ cstc, int n_tiles,
                                                                          float* __restrict__ cfeat, int* __restrict__ cidx,
                                                                          short* __restrict__ fixrec, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t a_full[2], a_empty[2], w_full[CT_NW], w_empty[CT_NW], acc_full[2], acc_empty[2], m_full[CT_NMETA];
  __shared__ uint32_t tmem_slot;
  __shared__ CtMeta meta[CT_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* ximg = base;                              // [CT_XIMG]
  unsigned char* wring = base + CT_XIMG;                   // [CT_NW][CT_WSTAGE]
  const float* wnorm = reinterpret_cast<const float*>(wimg + CT_IMG_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lg = L + 2;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;
  const int n_pairs = (n_mine + 1) >> 1;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 256);
      mbar_init(&a_full[s], 256); mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < CT_NW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < CT_NMETA; ++s) mbar_init(&m_full[s], 128);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders: group j (4 warps) owns tile j of every pair
    const int j = warp >> 2, w4 = warp & 3, gt = tid & 127;
    const int kq = (lane & 15) * 4, rsub = lane >> 4;               // a half-warp per row: 16 lanes x 4 channels = one channel half
    float4 va[CT_LROWS];
    // the group's tiles are blockIdx.x + (j + 2i) * gridDim.x; their tables are read one / two tiles ahead (TabPipe)
    TabPipe tp;
    if (tso) {
      tp.init(tso, cstc, blockIdx.x + j * gridDim.x, 2 * gridDim.x, n_tiles, gt);
      tp.prefetch(tso, cstc, blockIdx.x + j * gridDim.x, 2 * gridDim.x, n_tiles, gt);
    }
    auto build = [&](int t) {                                       // tile bookkeeping by the group's 128 threads
      const int tile = blockIdx.x + t * gridDim.x;
      CtMeta& m = meta[t % CT_NMETA];
      m.rowsrc[gt] = -1;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + j) : "memory");
      int s0, ns;
      if (tso) { s0 = tp.s0c; ns = tp.s1c - s0; } else { s0 = tile * gs; ns = min(gs, N - s0); }
      if (gt < ns) {
        int rb, len;
        if (tso) { rb = tp.cbc - tp.c0c; len = tp.cec - tp.cbc - 2; }
        else { rb = gt * Lg; len = L; }
        m.sb[gt] = rb; m.len[gt] = len;
        const int g0 = (s0 + gt) * L;
        for (int l = 0; l < len; ++l) m.rowsrc[rb + 1 + l] = g0 + l;
      }
      if (gt == 0) { m.s0 = s0; m.ns = ns; }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + j) : "memory");
      mbar_arrive(&m_full[t % CT_NMETA]);
    };
    // image rows of this tile: r = 0..135 (tile row r-1) at image row 128*j + r; the first tile stops at r = 129 (the rows behind
    // belong to the second tile)
    const int rmax = j ? 136 : 130;
    auto load = [&](const CtMeta& m, int h) {                       // channel half h of every row, into registers
#pragma unroll
      for (int i = 0; i < CT_LROWS; ++i) {
        const int r = (i * 4 + w4) * 2 + rsub, tr = r - 1;
        const int src = (tr >= 0 && tr < 128) ? m.rowsrc[tr] : -1;
        va[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + (size_t)src * D + h * 64 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store = [&](int h) {
      unsigned char* sub = ximg + h * 2 * CT_XSUB;
#pragma unroll
      for (int i = 0; i < CT_LROWS; ++i) {
        const int r = (i * 4 + w4) * 2 + rsub;
        if (r < rmax) store_split4(sub, sub + CT_XSUB, 128 * j + r, kq, va[i]);
      }
    };
    const bool off = (dbg & 2) != 0;
    if (j < n_mine) { build(j); if (!off) load(meta[j % CT_NMETA], 0); }
    for (int p = 0; p < n_pairs; ++p) {
      const int t = 2 * p + j;
      const bool has = t < n_mine && !off;
      if (p >= 1) mbar_wait(&a_empty[0], (p - 1) & 1);     // the MMAs of the previous pair are through with the first half-image
      if (has) { store(0); fence_async_smem(); }
      mbar_arrive(&a_full[0]);
      if (has) load(meta[t % CT_NMETA], 1);
      if (p >= 1) mbar_wait(&a_empty[1], (p - 1) & 1);
      if (has) { store(1); fence_async_smem(); }
      mbar_arrive(&a_full[1]);
      if (t + 2 < n_mine) {                                        // next pair: bookkeeping, first half in flight
        if (tso) { tp.rotate(); tp.prefetch(tso, cstc, blockIdx.x + (t + 2) * gridDim.x, 2 * gridDim.x, n_tiles, gt); }
        build(t + 2);
        if (!off) load(meta[(t + 2) % CT_NMETA], 0);
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
    int wit = 0;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1;
      const uint32_t idesc = 2 * p + 1 < n_mine ? idesc_bf16(128, 256) : idesc_bf16(128, 128);      // one MMA covers both tiles
      if (p >= 2) mbar_wait(&acc_empty[buf], ((p >> 1) - 1) & 1);
      const uint32_t d = tmem + buf * 256;
#pragma unroll 1
      for (int kb = 0; kb < CT_KB; ++kb, ++wit) {
        const int s = wit % CT_NW;
        const int ch = kb >= 3, dt = kb - 3 * ch;
        mbar_wait(&w_full[s], (wit / CT_NW) & 1);
        if (dt == 0) mbar_wait(&a_full[ch], p & 1);
        tc_fence_after();
        const uint32_t wst = smem_u32(wring + s * CT_WSTAGE);
        const uint64_t wh = smem_desc_sw128(wst), wl = smem_desc_sw128(wst + CT_TILE);
        const uint32_t xa = smem_u32(ximg + ch * 2 * CT_XSUB) + dt * 128;      // tap dt = the image one row further
        const uint64_t xh = smem_desc_sw128(xa), xl = smem_desc_sw128(xa + CT_XSUB);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (dbg & 4) break;
          const uint64_t o = (uint64_t)(kk * 2);
          umma_bf16_e(el, d, wh + o, xh + o, idesc, (kb | kk) != 0);
          umma_bf16_e(el, d, wh + o, xl + o, idesc, 1);
          umma_bf16_e(el, d, wl + o, xh + o, idesc, 1);
        }
        umma_commit_e(el, &w_empty[s]);
        if (dt == 2) umma_commit_e(el, &a_empty[ch]);      // this half-image may be overwritten
      }
      umma_commit_e(el, &acc_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = filter; warps 10-13 the even, 14-17 the odd sentences
    const int half = (warp - 10) >> 2, q = warp & 3, f = q * 32 + lane;
    const bool act = f < KC;
    const float bz = act ? bias[f] : 0.f;
    const float tau = act ? CT_TAU * wnorm[f] : 0.f;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1;
      mbar_wait(&acc_full[buf], (p >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int t = 2 * p + g;
        if (t >= n_mine) break;
        mbar_wait(&m_full[t % CT_NMETA], (t / CT_NMETA) & 1);
        if (dbg & 1) continue;
        const CtMeta& m = meta[t % CT_NMETA];
        const int n0 = m.s0, ns = m.ns;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + buf * 256 + g * 128;
#pragma unroll 1
        for (int sn = half; sn < ns; sn += 2) {
          // max over the L positions of the sentence (model.py:120); first maximum wins, <= 0 carries no gradient
          const int sb = m.sb[sn], len = m.len[sn];
          const int np = min(len + 1, L);              // positions whose window touches a valid row
          float best = -INFINITY, second = -INFINITY;
          int arg = -1, arg2 = -1;
          const int first = sb + 1;                    // accumulator columns [first, first + np) in 8-column aligned chunks
#pragma unroll 1
          for (int c0 = first & ~7; c0 < first + np; c0 += 8) {
            float v[8];
            tmem_ld8(trow + c0, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int l = c0 + i - first;
              // pre-activation (ReLU is applied after the max: monotone); columns outside the sentence count as -inf.  Selects, not
              // branches: every thread of the warp scans another filter.  All-zero windows give exactly the bias: a true tie, first wins
              const float y = (l >= 0 && l < np) ? v[i] + bz : -INFINITY;
              const bool gt = y > best;
              const bool g2 = !gt && y > second && !(y == best && y == bz);
              second = gt ? best : (g2 ? y : second);
              arg2 = gt ? arg : (g2 ? l : arg2);
              best = gt ? y : best;
              arg = gt ? l : arg;
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
            // the arg-max routes the gradient: near-ties (and maxima next to the ReLU threshold) are re-scored exactly in fp32 by
            // cnet_conv_fix_kernel; one 2-byte record per (sentence, filter): -1 = nothing to do, else arg | (arg2 + 1) << 7
            const bool tie = arg2 >= 0 && best - second <= tau;
            fixrec[o] = ((tie && best > -tau) || fabsf(best) <= tau) ? (short)(arg | ((tie ? arg2 + 1 : 0) << 7)) : (short)-1;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// exact fp32 re-scoring of the uncertain (sentence, filter) pairs: every lane scans 8 records (one 16-byte load), the whole warp
// re-scores each flagged one - the filter's 384 weights as three coalesced float4 per lane, the (up to) six x rows of the two
// candidate windows all in flight together
__global__ void __launch_bounds__(256) cnet_conv_fix_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            const short* __restrict__ fixrec, long n_rec, int L, int KC,
                                                            float* __restrict__ cfeat, int* __restrict__ cidx) {
  const int lane = threadIdx.x & 31;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long i0 = wid * 256; i0 < n_rec; i0 += nw * 256) {
    const long mine0 = i0 + lane * 8;
    __align__(16) short r8[8];
    if (mine0 + 8 <= n_rec) {
      *reinterpret_cast<uint4*>(r8) = *reinterpret_cast<const uint4*>(fixrec + mine0);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) r8[e] = mine0 + e < n_rec ? fixrec[mine0 + e] : (short)-1;
    }
    unsigned flags = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) flags |= (r8[e] >= 0 ? 1u : 0u) << e;
    unsigned todo = __ballot_sync(0xffffffffu, flags != 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      unsigned fl = __shfl_sync(0xffffffffu, flags, src);
      while (fl) {
        const int e = __ffs(fl) - 1;
        fl &= fl - 1;
        int recv = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) recv = q == e ? (int)r8[q] : recv;
        const int rec = __shfl_sync(0xffffffffu, recv, src);
        const long idx = i0 + src * 8 + e;
        const int n = (int)(idx / KC), kf = (int)(idx - (long)n * KC);
        const int c0 = rec & 127, c1 = (rec >> 7) - 1;             // the maximum's position, the runner-up's (-1: none)
        const float4* wf = reinterpret_cast<const float4*>(w + (size_t)kf * D * 3) + lane * 3;      // channels 4*lane..+3, 3 taps each
        const float4 w0 = wf[0], w1 = wf[1], w2 = wf[2];
        const float wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
        float4 xv[2][3];
#pragma unroll
        for (int cand = 0; cand < 2; ++cand) {
          const int l = cand ? c1 : c0;
#pragma unroll
          for (int dt = 0; dt < 3; ++dt) {
            const int ll = l + dt - 1;
            xv[cand][dt] = (l >= 0 && ll >= 0 && ll < L) ? *reinterpret_cast<const float4*>(x + ((size_t)n * L + ll) * D + lane * 4) : zero4;
          }
        }
        float best = 0.f;
        int arg = -1;
#pragma unroll
        for (int cand = 0; cand < 2; ++cand) {
          const int l = cand ? c1 : c0;
          if (l < 0) continue;
          float a = 0.f;
#pragma unroll
          for (int dt = 0; dt < 3; ++dt)                           // same order of additions as before: taps outer, channels inner
            a += xv[cand][dt].x * wv[dt] + xv[cand][dt].y * wv[3 + dt] + xv[cand][dt].z * wv[6 + dt] + xv[cand][dt].w * wv[9 + dt];
          a = warp_sum(a) + bias[kf];
          a = fmaxf(a, 0.f);
          if (arg < 0 || a > best || (a == best && l < arg)) { best = a; arg = l; }
        }
        if (lane == 0) {
          cfeat[idx] = best;
          cidx[idx] = best > 0.f ? arg : -1;
        }
      }
    }
  }
}

}  // namespace umpr

using namespace umpr;

namespace umpr {
// prep = 0: the weight image (and filter norms) in `wimg` are those of a previous call with the same conv_w (the three C-Net calls of
// one step share them, csrc/step.cu); prep = 2: ONLY build the image
int cnet_conv_fwd_tc_impl(const float* x, const float* conv_w, const float* conv_b, int N, int L, int KC, int ksize,
                          const int32_t* table, int table_tiles, void* wimg, int cap, float* cfeat, int32_t* cidx,
                          int n_ctas, int prep, void* fix_records, void* stream) {
  if (N <= 0) return 0;
  if (ksize != 3) return fail_arg("cnet: kernel_size=%d (only 3 is built)", ksize);
  if (KC < 1 || KC > 128) return fail_arg("cnet: kernel_count=%d must be in [1, 128]", KC);
  if (L < 1 || L + 2 > 128) return fail_arg("cnet_conv_fwd: sentence length L=%d must be in [1, 126]", L);
  if ((long)cap * 8 < (long)N * KC) return fail_arg("cnet_conv_fwd_tc: cap=%d is below N*KC/8 = %ld (one 2-byte record per sentence and filter)", cap, ((long)N * KC + 7) / 8);
  if (reinterpret_cast<uintptr_t>(wimg) & 15) return fail_arg("cnet_conv_fwd_tc: wimg must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(conv_w)) & 15) return fail_arg("cnet_conv_fwd_tc: x and conv_w must be 16-byte aligned");
  if (prep) {
    cnet_tc_prep_kernel<<<(CT_KB * 128 * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conv_w, KC, reinterpret_cast<unsigned char*>(wimg));
    if (int e = check_launch("cnet_tc_prep")) return e;
    if (prep == 2) return 0;
  }
  int gs = 128 / (L + 2);
  if (gs > 16) gs = 16;
  if (table && (table_tiles < 1 || table_tiles > N)) return fail_arg("cnet_conv_fwd_tc: tile table inconsistent (n_tiles=%d, N=%d)", table_tiles, N);
  const int n_tiles = table ? table_tiles : (N + gs - 1) / gs;
  const int smem = CT_XIMG + CT_NW * CT_WSTAGE + 1024;
  {
    cudaError_t e = cudaFuncSetAttribute(cnet_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // (per device: every call)
    if (e != cudaSuccess) { set_error("cnet_conv_fwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  // the 2-byte records behind the weight image by default; callers that run several sides concurrently pass one buffer per side
  short* fixrec = fix_records ? reinterpret_cast<short*>(fix_records) : reinterpret_cast<short*>(reinterpret_cast<unsigned char*>(wimg) + CT_HDR_BYTES);
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("UMPR_CONV_DBG"); dbg = e ? atoi(e) : 0; }
  cnet_conv_fwd_tc_kernel<<<grid, CT2_THREADS, smem, (cudaStream_t)stream>>>(x, reinterpret_cast<const unsigned char*>(wimg), conv_b, N, L, KC,
                                                                           gs, table, table ? table + n_tiles + 1 : nullptr, n_tiles, cfeat, cidx, fixrec, dbg);
  if (int rc = check_launch("cnet_conv_fwd_tc")) return rc;
  const long n_rec = (long)N * KC;
  long blocks = (n_rec + 8 * 256 - 1) / (8 * 256);               // one 256-record pass per warp
  if (blocks > n_ctas * 8) blocks = n_ctas * 8;
  cnet_conv_fix_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, conv_w, conv_b, fixrec, n_rec, L, KC, cfeat, cidx);
  return check_launch("cnet_conv_fix");
}
}  // namespace umpr

// scratch: 197632 + 16*cap bytes, 16-byte aligned (weight image, filter norms, one 2-byte re-scoring record per (sentence, filter):
// cap >= ceil(N*KC / 8)).
// table (optional) = [tile_sent_off (n_tiles+1) | cstart (N+1)] with cstart the exclusive prefix sum of (len + 2) per sentence
// (plan.py:cnet_table): sentences tile_sent_off[k] .. tile_sent_off[k+1]-1 form tile k, at most 128 rows including the guards.
extern "C" int umpr_cnet_conv_fwd_tc(const float* x, const float* conv_w, const float* conv_b, int N, int L, int KC, int ksize,
                                     const int32_t* table, int table_tiles, void* wimg, int cap, float* cfeat, int32_t* cidx,
                                     int n_ctas, void* stream) {
  return cnet_conv_fwd_tc_impl(x, conv_w, conv_b, N, L, KC, ksize, table, table_tiles, wimg, cap, cfeat, cidx, n_ctas, 1, nullptr, stream);
}
