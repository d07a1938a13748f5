// C-Net convolution weight gradient on tcgen05 (backward of reference src/model.py:118-120).
//   dW[k][c][j] = sum_n g[n][k] x[n][arg(n,k) + j - 1][c]        (g = gradient of the pooled feature, arg = its max position)
// is sparse per (sentence, filter) - 1.5 KB of row gathers per pair on CUDA cores - but dense per tap once the gradients are
// scattered into a one-hot tile:  dW_j^T [c x k] = x^T [c x rows] . G_j [rows x k],  G_j[r][k] = g[n][k] at the row r that tap j of
// the winning window reads.  Tiles are the convolution's own (plan.py:cnet_table: valid rows of whole sentences between zero guard
// rows, <= 128 rows); the x image is loaded once per tile and used MN-major; G_j is rebuilt in shared memory for each tap
// (zero-fill + scatter of <= ns*K bf16 hi/lo pairs); three accumulators (one per tap) live in tensor memory for the CTA's whole
// queue and are flushed once.  warps 0-3: x loaders, warp 4: MMA issuer, warps 5-8: gradient scatter + flush.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int CB_THREADS = 288;
constexpr int CB_XIMG = 65536;          // [c block 2][hi|lo][128 rows][128 B]
constexpr int CB_GIMG = 65536;          // [k block 2][hi|lo][128 rows][128 B]
constexpr int CB_NMETA = 4;
constexpr int CB_MAXS = 44;
constexpr int CB_RC = 16;             // (gradient, position) pairs a scatter thread keeps in registers per tile

struct CbMeta {
  int s0, ns;
  int sb[CB_MAXS];                      // tile row of each sentence's leading guard row
  int len[CB_MAXS];
  int rowsrc[128];                      // tile row -> global x row, -1 = zero row
};

// ------------------------------------------------------------------------------------------------------------------------
// Input gradient on tcgen05:  dX [rows x c] = sum_j G_j [rows x k] . W_j [k x c]  with the same one-hot gradient tiles (K-major A
// operand this time) and the tap's weights W_j^T as a K-major image streamed from L2 by one TMA bulk copy per tap.
//   warps 0-3: scatter, warp 4: MMA issuer + TMA producer, warps 5-8: epilogue (valid rows only, coalesced row stores)
// ------------------------------------------------------------------------------------------------------------------------
constexpr int CX_WIMG = 65536;          // per tap: [k block 2][hi|lo][128 c rows][128 B]
constexpr int CX_STG_LD = 36;

// wimg[tap][kb][hi|lo][c][kk] = W[kb*64 + kk][c][tap]
__global__ void cnet_bwd_wimg_kernel(const float* __restrict__ w, int KC, unsigned char* __restrict__ wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (tap, kb, c, k4)
  if (idx >= 3 * 2 * 128 * 16) return;
  const int tap = idx / (2 * 128 * 16), rem = idx - tap * 2 * 128 * 16, kb = rem / (128 * 16), c = (rem >> 4) & 127, kk = (rem & 15) * 4;
  float t[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { const int k = kb * 64 + kk + q; t[q] = k < KC ? w[((size_t)k * D + c) * 3 + tap] : 0.f; }
  unsigned char* img = wimg + (size_t)tap * CX_WIMG + kb * 32768;
  store_split4(img, img + 16384, c, kk, make_float4(t[0], t[1], t[2], t[3]));
}

// ---- v2 ----------------------------------------------------------------------------------------------------------------
// ONE gradient image per tile instead of one one-hot tile per (tap, filter half):  dY[r][k] = g[n][k] at the tile row r of the
// winning position (zero elsewhere; positions beyond the sentence's trailing guard row touch no valid x row and are dropped), laid
// out like the forward's x image - [filter half][hi|lo][130 rows][64 bf16], image row = tile row + 1 - and
//     dX[r][c] = sum_j sum_k dY[r + 1 - j][k] W[k][c][j]
// reads it through descriptors shifted by (2 - j) rows: zero-fill + scatter ONCE per tile, 72 MMAs back to back.  Two tiles share
// every weight stage (tap, filter half: 32 KB by TMA bulk copy); K runs over the first filter half (3 taps), then the second, so the
// scatter of the NEXT pair's first half runs under the MMAs of this pair's second half.
//   warps 0-3 / 4-7: bookkeeping + scatter of the first / second tile of a pair, warp 8: weight producer, warp 9: MMA issuer,
//   warps 10-13: epilogue (valid rows only, row pointers resolved once per tile, 64-byte row pieces)
constexpr int DX_THREADS = 448;
constexpr int DX_GSUB = 17 * 1024;       // [136 rows][128 B]
constexpr int DX_GIMG = 4 * DX_GSUB;     // [filter half 2][hi|lo]
constexpr int DX_WST = 32768;            // one weight stage: [hi|lo][128 c rows][128 B] of (tap, filter half)
constexpr int DX_NMETA = 8;
constexpr int DX_STG_LD = 20;            // 16 staged columns + 4 floats of padding
constexpr int DX_RC = 8;                 // (gradient, position) pairs per filter half a scatter thread prefetches into registers

struct DxBars { uint64_t m_full[DX_NMETA], g_full[2], g_empty[2], w_full[2], w_empty[2], acc_full[2], acc_empty[2]; };
struct DxCache { float g[2][DX_RC]; int t[2][DX_RC]; };

__global__ void __launch_bounds__(DX_THREADS, 1) cnet_conv_bwd_dx_tc_kernel(const float* __restrict__ dcfeat, const int* __restrict__ cidx,
                                                                            const unsigned char* __restrict__ wimg, const int* __restrict__ tso,
                                                                            const int* __restrict__ cstc, int n_tiles, int L, int KC,
                                                                            float* __restrict__ dx, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ DxBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ CbMeta meta[DX_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* gim = base;                               // [tile of the pair 2][DX_GIMG]
  unsigned char* wsm = base + 2 * DX_GIMG;                 // [2][DX_WST]
  float* stg = reinterpret_cast<float*>(base + 2 * DX_GIMG + 2 * DX_WST);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;
  const int n_pairs = (n_mine + 1) >> 1;

  if (tid == 0) {
    for (int s = 0; s < DX_NMETA; ++s) mbar_init(&bar.m_full[s], 128);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar.g_full[s], 256); mbar_init(&bar.g_empty[s], 1);
      mbar_init(&bar.w_full[s], 1); mbar_init(&bar.w_empty[s], 1);
      mbar_init(&bar.acc_full[s], 1); mbar_init(&bar.acc_empty[s], 128);
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ tile bookkeeping + gradient scatter: group jt owns tile jt of a pair
    const int jt = warp >> 2, et = tid & 127, kk = et & 63, sg = et >> 6;
    unsigned char* img = gim + jt * DX_GIMG;
    // the group's tiles are blockIdx.x + (jt + 2i) * gridDim.x; their tables are read one / two tiles ahead (TabPipe)
    TabPipe tp;
    tp.init(tso, cstc, blockIdx.x + jt * gridDim.x, 2 * gridDim.x, n_tiles, et);
    tp.prefetch(tso, cstc, blockIdx.x + jt * gridDim.x, 2 * gridDim.x, n_tiles, et);
    auto build = [&](int t) {                                       // bookkeeping of tile t from the pipeline's current values
      CbMeta& m = meta[t % DX_NMETA];
      m.rowsrc[et] = -1;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + jt) : "memory");
      const int s0 = tp.s0c, ns = tp.s1c - s0;
      if (et < ns) {
        const int b = tp.cbc - tp.c0c, len = tp.cec - tp.cbc - 2;
        m.sb[et] = b; m.len[et] = len;
        const int g0 = (s0 + et) * L;
        for (int l = 0; l < len; ++l) m.rowsrc[b + 1 + l] = g0 + l;
      }
      if (et == 0) { m.s0 = s0; m.ns = ns; }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + jt) : "memory");
      mbar_arrive(&bar.m_full[t % DX_NMETA]);
    };
    // thread et owns filter kk of each half and every second sentence (sg, sg+2, ...): the first DX_RC of them are fetched ahead
    auto prefetch = [&](const CbMeta& m, DxCache& c) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < DX_RC; ++i) {
          const int sn = 2 * i + sg, k = h * 64 + kk;
          c.g[h][i] = 0.f; c.t[h][i] = -1;
          if (sn < m.ns && k < KC) {
            const size_t o = (size_t)(m.s0 + sn) * KC + k;
            c.g[h][i] = dcfeat[o];
            c.t[h][i] = cidx[o];
          }
        }
    };
    auto put = [&](unsigned char* half, const CbMeta& m, int sn, float g, int t) {
      if (g == 0.f || t < 0 || t > m.len[sn]) return;            // positions past the trailing guard row touch no valid x row
      const int r = m.sb[sn] + t + 2;                             // image row = tile row (sb + 1 + t) + 1
      const __nv_bfloat16 hi = __float2bfloat16_rn(g);
      const __nv_bfloat16 lo = __float2bfloat16_rn(g - __bfloat162float(hi));
      const uint32_t off = (uint32_t)(r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2);
      *reinterpret_cast<__nv_bfloat16*>(half + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(half + DX_GSUB + off) = lo;
    };
    DxCache cur, nxt;
    if (jt < n_mine) { build(jt); prefetch(meta[jt % DX_NMETA], cur); }
    for (int p = 0; p < n_pairs; ++p) {
      const int t = 2 * p + jt;
      const bool has = t < n_mine && !(dbg & 2);
      const CbMeta& m = meta[t % DX_NMETA];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (p >= 1) mbar_wait(&bar.g_empty[h], (p - 1) & 1);      // the MMAs of the previous pair are through with this filter half
        if (has) {
          unsigned char* half = img + h * 2 * DX_GSUB;
          for (int i = et; i < 2 * DX_GSUB / 16; i += 128) reinterpret_cast<uint4*>(half)[i] = make_uint4(0u, 0u, 0u, 0u);
          asm volatile("bar.sync %0, 128;" ::"r"(2 + jt) : "memory");
#pragma unroll
          for (int i = 0; i < DX_RC; ++i) put(half, m, 2 * i + sg, cur.g[h][i], cur.t[h][i]);
          const int k = h * 64 + kk;
          if (k < KC)
            for (int sn = 2 * DX_RC + sg; sn < m.ns; sn += 2) {
              const size_t o = (size_t)(m.s0 + sn) * KC + k;
              put(half, m, sn, dcfeat[o], cidx[o]);
            }
          fence_async_smem();
        }
        mbar_arrive(&bar.g_full[h]);
        if (h == 0 && t + 2 < n_mine) {                            // next pair, early
          tp.rotate();
          tp.prefetch(tso, cstc, blockIdx.x + (t + 2) * gridDim.x, 2 * gridDim.x, n_tiles, et);
          build(t + 2);
          prefetch(meta[(t + 2) % DX_NMETA], nxt);
        }
      }
      cur = nxt;
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ weight producer: stage (filter half kb, tap j), 6 per pair
    const uint32_t el = elect_one_sync();
    const int n_it = n_pairs * 6;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1, st = it % 6, kb = st / 3, j = st - 3 * kb;
      if (it >= 2) mbar_wait(&bar.w_empty[s], ((it >> 1) - 1) & 1);
      mbar_arrive_expect_tx_e(el, &bar.w_full[s], DX_WST);
      bulk_copy_g2s_e(el, wsm + s * DX_WST, wimg + (size_t)j * CX_WIMG + kb * DX_WST, DX_WST, &bar.w_full[s]);
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, the elected lane issues)
    const uint32_t el = elect_one_sync();
    constexpr uint32_t idesc = idesc_bf16(128, 128);              // A (dY) and B (W_j^T) K-major
    int wit = 0;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1;
      const bool two = 2 * p + 1 < n_mine;
      if (p >= 2) mbar_wait(&bar.acc_empty[buf], ((p >> 1) - 1) & 1);
#pragma unroll 1
      for (int st = 0; st < 6; ++st, ++wit) {
        const int s = wit & 1, kb = st / 3, j = st - 3 * kb;
        mbar_wait(&bar.w_full[s], (wit >> 1) & 1);
        if (j == 0) mbar_wait(&bar.g_full[kb], p & 1);
        tc_fence_after();
        const uint32_t w0 = smem_u32(wsm + s * DX_WST);
        const uint64_t wh = smem_desc_sw128(w0), wl = smem_desc_sw128(w0 + 16384);
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          if (t2 == 1 && !two) break;
          const uint32_t g0 = smem_u32(gim + t2 * DX_GIMG + kb * 2 * DX_GSUB) + (2 - j) * 128;      // dY rows r + 1 - j
          const uint64_t gh = smem_desc_sw128(g0), gl = smem_desc_sw128(g0 + DX_GSUB);
          const uint32_t d = tmem + buf * 256 + t2 * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (dbg & 4) break;
            const uint64_t o = (uint64_t)(q * 2);
            umma_bf16_e(el, d, gh + o, wh + o, idesc, (st | q) != 0);
            umma_bf16_e(el, d, gh + o, wl + o, idesc, 1);
            umma_bf16_e(el, d, gl + o, wh + o, idesc, 1);
          }
        }
        umma_commit_e(el, &bar.w_empty[s]);
        if (j == 2) umma_commit_e(el, &bar.g_empty[kb]);
      }
      umma_commit_e(el, &bar.acc_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: valid rows of dx
    const int q4 = warp & 3;
    float* sw = stg + q4 * 32 * DX_STG_LD;
    const int c4 = (lane & 3) * 4;
    for (int p = 0; p < n_pairs; ++p) {
      const int buf = p & 1;
      mbar_wait(&bar.acc_full[buf], (p >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t2 = 0; t2 < 2; ++t2) {
        const int t = 2 * p + t2;
        if (t >= n_mine) break;
        mbar_wait(&bar.m_full[t % DX_NMETA], (t / DX_NMETA) & 1);
        if (dbg & 1) continue;
        const CbMeta& m = meta[t % DX_NMETA];
        float* orow[4];                                             // the 4 output rows this thread stores (row i*8 + lane/4 of the warp's 32)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int src = m.rowsrc[q4 * 32 + i * 8 + (lane >> 2)];
          orow[i] = src >= 0 ? dx + (size_t)src * D + c4 : nullptr;
        }
        const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16) + buf * 256 + t2 * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < D; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<float4*>(&sw[lane * DX_STG_LD + jj * 4]) = make_float4(v[jj * 4], v[jj * 4 + 1], v[jj * 4 + 2], v[jj * 4 + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 o = *reinterpret_cast<const float4*>(&sw[(i * 8 + (lane >> 2)) * DX_STG_LD + c4]);
            if (orow[i]) *reinterpret_cast<float4*>(orow[i] + c0) = o;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(&bar.acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------------
// Weight gradient, same gradient image:  dW_j^T [c x k] = sum_r x[r][c] dY[r + 1 - j][k]  - the x image of the tile MN-major as the
// A operand (unshifted), the dY image MN-major as the B operand read (2 - j) rows further down; 72 MMAs (N = 128) per tile into three
// accumulators (one per tap) that live in tensor memory for the CTA's whole queue and are flushed once.
//   warps 0-7: bookkeeping + x loaders (the next tile's row loads are in flight in registers while the MMAs of this one run; the
//   image is single-buffered), warps 8-11: gradient scatter (dY double-buffered) and the final flush, warp 12: MMA issuer
// ------------------------------------------------------------------------------------------------------------------------
constexpr int DW_THREADS = 416;
struct DwBars { uint64_t a_full, a_empty, m_full[CB_NMETA], g_full[2], g_empty[2], w_full; };

__global__ void __launch_bounds__(DW_THREADS, 1) cnet_conv_bwd_dw_tc_kernel(const float* __restrict__ x, const float* __restrict__ dcfeat,
                                                                            const int* __restrict__ cidx, const int* __restrict__ tso,
                                                                            const int* __restrict__ cstc, int n_tiles, int L, int KC,
                                                                            float* __restrict__ dw, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ DwBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ CbMeta meta[CB_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* xim = base;                     // [c block 2][hi|lo][128 rows][128 B]
  unsigned char* gim = base + CB_XIMG;           // [2][DX_GIMG]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;

  if (tid == 0) {
    mbar_init(&bar.a_full, 256); mbar_init(&bar.a_empty, 1);
    for (int s = 0; s < CB_NMETA; ++s) mbar_init(&bar.m_full[s], 256);
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.g_full[s], 128); mbar_init(&bar.g_empty[s], 1); }
    mbar_init(&bar.w_full, 1);
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders: tile bookkeeping + x rows -> bf16 hi/lo image
    // (the tile tables are read one / two tiles ahead: no table load sits in front of a tile's row loads)
    TabPipe tp;
    tp.init(tso, cstc, blockIdx.x, gridDim.x, n_tiles, tid);
    auto build = [&](int it) {                                      // bookkeeping of tile `it` from the pipeline's current values
      CbMeta& m = meta[it % CB_NMETA];
      if (tid < 128) m.rowsrc[tid] = -1;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int s0 = tp.s0c, ns = tp.s1c - s0;
      if (tid < ns) {
        const int b = tp.cbc - tp.c0c, len = tp.cec - tp.cbc - 2;
        m.sb[tid] = b; m.len[tid] = len;
        const int g0 = (s0 + tid) * L;
        for (int l = 0; l < len; ++l) m.rowsrc[b + 1 + l] = g0 + l;
      }
      if (tid == 0) { m.s0 = s0; m.ns = ns; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_arrive(&bar.m_full[it % CB_NMETA]);
    };
    float4 va[16];                                                  // a warp per row and pass: rows i*8 + warp, 32 lanes x 4 channels
    auto load = [&](const CbMeta& m) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int src = (dbg & 2) ? -1 : m.rowsrc[i * 8 + warp];
        va[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + (size_t)src * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    tp.prefetch(tso, cstc, blockIdx.x, gridDim.x, n_tiles, tid);
    build(0);
    load(meta[0]);
    unsigned char* sub = xim + (lane >> 4) * 32768;
    for (int it = 0; it < n_mine; ++it) {
      if (it >= 1) mbar_wait(&bar.a_empty, (it - 1) & 1);          // the MMAs of the previous tile have read the image
#pragma unroll
      for (int i = 0; i < 16; ++i) store_split4(sub, sub + 16384, i * 8 + warp, (lane & 15) * 4, va[i]);
      fence_async_smem();
      mbar_arrive(&bar.a_full);
      if (it + 1 < n_mine) {
        tp.rotate();                                               // tile it+1 becomes current; its tables were fetched an iteration ago
        tp.prefetch(tso, cstc, blockIdx.x + (it + 1) * gridDim.x, gridDim.x, n_tiles, tid);
        build(it + 1);
        load(meta[(it + 1) % CB_NMETA]);
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, the elected lane issues)
    const uint32_t el = elect_one_sync();
    constexpr uint32_t idesc = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);       // A (x) and B (dY) MN-major
    const uint32_t a0 = smem_u32(xim);
    for (int it = 0; it < n_mine; ++it) {
      const int gb = it & 1;
      mbar_wait(&bar.a_full, it & 1);
      mbar_wait(&bar.g_full[gb], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t g0 = smem_u32(gim + gb * DX_GIMG);
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const uint32_t gj = g0 + (2 - j) * 128;                    // dY rows r + 1 - j (image row = tile row + 1)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (dbg & 4) break;
          const uint64_t xh = desc_mn(a0 + ks * 2048, 32768), xl = desc_mn(a0 + 16384 + ks * 2048, 32768);
          const uint64_t gh = desc_mn(gj + ks * 2048, 2 * DX_GSUB), gl = desc_mn(gj + DX_GSUB + ks * 2048, 2 * DX_GSUB);
          const uint32_t d = tmem + j * 128;
          umma_bf16_e(el, d, xh, gh, idesc, (it | ks) != 0);
          umma_bf16_e(el, d, xh, gl, idesc, 1);
          umma_bf16_e(el, d, xl, gh, idesc, 1);
        }
      }
      umma_commit_e(el, &bar.a_empty);
      umma_commit_e(el, &bar.g_empty[gb]);
    }
    umma_commit_e(el, &bar.w_full);
  } else {
    // ------------------------------------------------------------------ gradient scatter (one image per tile), final flush
    const int et = tid - 256, kk = et & 63, sg = et >> 6;
    auto put = [&](unsigned char* half, const CbMeta& m, int sn, float g, int t) {
      if (g == 0.f || t < 0 || t > m.len[sn]) return;            // positions past the trailing guard row touch no valid x row
      const int r = m.sb[sn] + t + 2;                             // image row = tile row (sb + 1 + t) + 1
      const __nv_bfloat16 hi = __float2bfloat16_rn(g);
      const __nv_bfloat16 lo = __float2bfloat16_rn(g - __bfloat162float(hi));
      const uint32_t off = (uint32_t)(r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2);
      *reinterpret_cast<__nv_bfloat16*>(half + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(half + DX_GSUB + off) = lo;
    };
    for (int it = 0; it < n_mine; ++it) {
      const int gb = it & 1;
      mbar_wait(&bar.m_full[it % CB_NMETA], (it / CB_NMETA) & 1);
      const CbMeta& m = meta[it % CB_NMETA];
      // the tile's (gradient, position) pairs: all loads in flight together, before the wait for the image buffer
      DxCache c;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < DX_RC; ++i) {
          const int sn = 2 * i + sg, k = h * 64 + kk;
          c.g[h][i] = 0.f; c.t[h][i] = -1;
          if (sn < m.ns && k < KC && !(dbg & 8)) {
            const size_t o = (size_t)(m.s0 + sn) * KC + k;
            c.g[h][i] = dcfeat[o];
            c.t[h][i] = cidx[o];
          }
        }
      if (it >= 2) mbar_wait(&bar.g_empty[gb], ((it >> 1) - 1) & 1);      // the MMAs of tile it-2 have read this buffer
      unsigned char* img = gim + gb * DX_GIMG;
      if (!(dbg & 8)) {
        for (int i = et; i < DX_GIMG / 16; i += 128) reinterpret_cast<uint4*>(img)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned char* half = img + h * 2 * DX_GSUB;
#pragma unroll
          for (int i = 0; i < DX_RC; ++i) put(half, m, 2 * i + sg, c.g[h][i], c.t[h][i]);
          const int k = h * 64 + kk;
          if (k < KC)
            for (int sn = 2 * DX_RC + sg; sn < m.ns; sn += 2) {
              const size_t o = (size_t)(m.s0 + sn) * KC + k;
              put(half, m, sn, dcfeat[o], cidx[o]);
            }
        }
        fence_async_smem();
      }
      mbar_arrive(&bar.g_full[gb]);
    }
    mbar_wait(&bar.w_full, 0);
    tc_fence_after();
    const int q4 = warp & 3, cch = q4 * 32 + lane;                // TMEM lane = channel
#pragma unroll 1
    for (int j = 0; j < 3 && !(dbg & 16); ++j)
#pragma unroll 1
      for (int k0 = 0; k0 < 128; k0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + j * 128 + k0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + i < KC) atomicAdd(&dw[((size_t)(k0 + i) * D + cch) * 3 + j], v[i]);
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

}  // namespace umpr

using namespace umpr;

// table = the convolution's tile table (plan.py:cnet_table): [tile_sent_off (n_tiles+1) | cstart (N+1)], cstart = prefix sum of len+2
extern "C" int umpr_cnet_conv_bwd_dw_tc(const float* x, const float* dcfeat, const int32_t* cidx, int N, int L, int KC, const int32_t* table,
                                        int n_tiles, float* d_conv_w, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (KC < 1 || KC > 128) return fail_arg("cnet: kernel_count=%d must be in [1, 128]", KC);
  if (L < 1 || L + 2 > 128) return fail_arg("cnet_conv_bwd_dw_tc: sentence length L=%d must be in [1, 126]", L);
  if (!table || n_tiles < 1 || n_tiles > N) return fail_arg("cnet_conv_bwd_dw_tc: tile table missing or inconsistent (n_tiles=%d, N=%d)", n_tiles, N);
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail_arg("cnet_conv_bwd_dw_tc: x must be 16-byte aligned");
  constexpr int smem = CB_XIMG + 2 * DX_GIMG + 1024;
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_bwd_dw_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_bwd_dw_tc_kernel<<<grid, DW_THREADS, smem, (cudaStream_t)stream>>>(x, dcfeat, cidx, table, table + n_tiles + 1, n_tiles, L, KC, d_conv_w, dbg_flags());
  return check_launch("cnet_conv_bwd_dw_tc");
}

// wimg_scratch: 3 * 65536 bytes, 16-byte aligned (per-tap weight images, rebuilt every call)
extern "C" int umpr_cnet_conv_bwd_dx_tc(const float* dcfeat, const int32_t* cidx, const float* conv_w, int N, int L, int KC,
                                        const int32_t* table, int n_tiles, void* wimg_scratch, float* dx, int n_ctas, void* stream) {
  if (N <= 0) return 0;
  if (KC < 1 || KC > 128) return fail_arg("cnet: kernel_count=%d must be in [1, 128]", KC);
  if (L < 1 || L + 2 > 128) return fail_arg("cnet_conv_bwd_dx_tc: sentence length L=%d must be in [1, 126]", L);
  if (!table || n_tiles < 1 || n_tiles > N) return fail_arg("cnet_conv_bwd_dx_tc: tile table missing or inconsistent (n_tiles=%d, N=%d)", n_tiles, N);
  if (!wimg_scratch || (reinterpret_cast<uintptr_t>(wimg_scratch) & 15)) return fail_arg("cnet_conv_bwd_dx_tc: scratch must be 16-byte aligned");
  cnet_bwd_wimg_kernel<<<(3 * 2 * 128 * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conv_w, KC, reinterpret_cast<unsigned char*>(wimg_scratch));
  if (int e = check_launch("cnet_bwd_wimg")) return e;
  constexpr int smem = 2 * DX_GIMG + 2 * DX_WST + 4 * 32 * DX_STG_LD * 4 + 1024;
  static_assert(smem <= 220 * 1024, "shared memory budget (the bookkeeping ring is static shared memory on top)");
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_bwd_dx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_bwd_dx_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_bwd_dx_tc_kernel<<<grid, DX_THREADS, smem, (cudaStream_t)stream>>>(dcfeat, cidx, reinterpret_cast<const unsigned char*>(wimg_scratch),
                                                                              table, table + n_tiles + 1, n_tiles, L, KC, dx, dbg_flags());
  return check_launch("cnet_conv_bwd_dx_tc");
}
