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

constexpr int CB_GHALF = 32768;         // one half (64 filters) of a one-hot gradient tile: [hi|lo][128 rows][128 B]

// The scatter side of both kernels.  A step = (tap j, filter half h); its one-hot tile G[r][k] (k in that half) holds g[n][k] at the
// row r that tap j of sentence n's winning window touches.  Two half-tile buffers ping-pong with the MMA issuer, so the scatter of
// step q+1 runs under the MMAs of step q.  Thread et owns filter column kk = et & 63 of the half and every second sentence.
struct GradCache {
  float g[2][CB_RC / 2];
  int t[2][CB_RC / 2];
};
__device__ __forceinline__ void grad_cache_load(GradCache& c, const float* __restrict__ dcfeat, const int* __restrict__ cidx, int s0, int ns,
                                                int KC, int et) {
  const int kk = et & 63, sg = et >> 6;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < CB_RC / 2; ++i) {
      const int sn = 2 * i + sg, k = h * 64 + kk;
      c.g[h][i] = 0.f; c.t[h][i] = -1;
      if (sn < ns && k < KC) {
        const size_t o = (size_t)(s0 + sn) * KC + k;
        c.g[h][i] = dcfeat[o];
        c.t[h][i] = cidx[o];
      }
    }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < CB_RC / 2; ++i)
      if (c.g[h][i] == 0.f) c.t[h][i] = -1;
}
__device__ __forceinline__ void grad_put(unsigned char* gbuf, const CbMeta& m, int sn, int kk, float g, int row) {
  if (row < 0 || row >= m.len[sn]) return;
  const int r = m.sb[sn] + 1 + row;
  const __nv_bfloat16 hi = __float2bfloat16_rn(g);
  const __nv_bfloat16 lo = __float2bfloat16_rn(g - __bfloat162float(hi));
  const uint32_t off = (uint32_t)(r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2);
  *reinterpret_cast<__nv_bfloat16*>(gbuf + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(gbuf + 16384 + off) = lo;
}
// one step: zero the half tile, scatter (register-cached pairs first, the rest re-read), publish
__device__ __forceinline__ void grad_scatter_step(unsigned char* gbuf, const CbMeta& m, const GradCache& c, const float* __restrict__ dcfeat,
                                                  const int* __restrict__ cidx, int KC, int j, int h, int et, int bar_id) {
  for (int i = et; i < CB_GHALF / 16; i += 128) reinterpret_cast<uint4*>(gbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
  const int kk = et & 63, sg = et >> 6, ns = m.ns;
#pragma unroll
  for (int i = 0; i < CB_RC / 2; ++i)
    if (c.t[h][i] >= 0) grad_put(gbuf, m, 2 * i + sg, kk, c.g[h][i], c.t[h][i] + j - 1);
  const int k = h * 64 + kk;
  if (k < KC)
    for (int sn = CB_RC + sg; sn < ns; sn += 2) {
      const size_t o = (size_t)(m.s0 + sn) * KC + k;
      const float g = dcfeat[o];
      const int t = cidx[o];
      if (g != 0.f && t >= 0) grad_put(gbuf, m, sn, kk, g, t + j - 1);
    }
  fence_async_smem();
}

struct CbBars { uint64_t a_full[2], a_empty[2], m_full[CB_NMETA], g_ready[2], g_free[2], w_full; };

__global__ void __launch_bounds__(CB_THREADS, 1) cnet_conv_bwd_dw_tc_kernel(const float* __restrict__ x, const float* __restrict__ dcfeat,
                                                                            const int* __restrict__ cidx, const int* __restrict__ tso,
                                                                            const int* __restrict__ cstc, int n_tiles, int L, int KC,
                                                                            float* __restrict__ dw) {
  extern __shared__ unsigned char raw[];
  __shared__ CbBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ CbMeta meta[CB_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* xim = base;                     // 2 stages
  unsigned char* gim = base + 2 * CB_XIMG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.a_full[s], 128); mbar_init(&bar.a_empty[s], 1); }
    for (int s = 0; s < CB_NMETA; ++s) mbar_init(&bar.m_full[s], 128);
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.g_ready[s], 128); mbar_init(&bar.g_free[s], 1); }
    mbar_init(&bar.w_full, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders: tile bookkeeping + x rows -> bf16 hi/lo image
    for (int it = 0; it < n_mine; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, s = it & 1;
      if (it >= 2) mbar_wait(&bar.a_empty[s], ((it >> 1) - 1) & 1);     // also frees meta slot it % 4 (tile it-4 is long finished)
      CbMeta& m = meta[it % CB_NMETA];
      m.rowsrc[tid] = -1;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int s0 = tso[tile], ns = tso[tile + 1] - s0;
      if (tid < ns) {
        const int c0 = cstc[s0], b = cstc[s0 + tid] - c0, len = cstc[s0 + tid + 1] - cstc[s0 + tid] - 2;
        m.sb[tid] = b; m.len[tid] = len;
        const int g0 = (s0 + tid) * L;
        for (int l = 0; l < len; ++l) m.rowsrc[b + 1 + l] = g0 + l;
      }
      if (tid == 0) { m.s0 = s0; m.ns = ns; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      unsigned char* st = xim + s * CB_XIMG;
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        unsigned char* a_hi = st + kb * 32768, *a_lo = a_hi + 16384;
        float4 va[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid, r = idx >> 4, k = kb * 64 + (idx & 15) * 4;
          const int src = m.rowsrc[r];
          va[i] = src >= 0 ? *reinterpret_cast<const float4*>(x + (size_t)src * D + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid;
          store_split4(a_hi, a_lo, idx >> 4, (idx & 15) * 4, va[i]);
        }
      }
      fence_async_smem();
      mbar_arrive(&bar.a_full[s]);
      mbar_arrive(&bar.m_full[it % CB_NMETA]);
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, the elected lane issues)
    if (n_mine > 0) {
      const uint32_t el = elect_one_sync();
      constexpr uint32_t idesc = idesc_bf16(128, 64) | (1u << 15) | (1u << 16);       // A (x) and B (G half tile) MN-major
      const uint32_t g0 = smem_u32(gim);
      int q = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int s = it & 1;
        mbar_wait(&bar.a_full[s], (it >> 1) & 1);
        const uint32_t a0 = smem_u32(xim + s * CB_XIMG);
        for (int j = 0; j < 3; ++j)
          for (int h = 0; h < 2; ++h, ++q) {
            const int gb = q & 1;
            mbar_wait(&bar.g_ready[gb], (q >> 1) & 1);
            tc_fence_after();
            const uint32_t gq0 = g0 + gb * CB_GHALF;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t xh = desc_mn(a0 + ks * 2048, 32768), xl = desc_mn(a0 + 16384 + ks * 2048, 32768);
              const uint64_t gh = desc_mn(gq0 + ks * 2048, 8192), gl = desc_mn(gq0 + 16384 + ks * 2048, 8192);
              const uint32_t d = tmem + j * 128 + h * 64;
              umma_bf16_e(el, d, xh, gh, idesc, (it | ks) != 0);
              umma_bf16_e(el, d, xh, gl, idesc, 1);
              umma_bf16_e(el, d, xl, gh, idesc, 1);
            }
            umma_commit_e(el, &bar.g_free[gb]);
            if (j == 2 && h == 1) umma_commit_e(el, &bar.a_empty[s]);
          }
      }
      umma_commit_e(el, &bar.w_full);
    }
  } else {
    // ------------------------------------------------------------------ gradient scatter (one-hot tile per tap), final flush
    const int et = tid - 160;
    int q = 0;
    for (int it = 0; it < n_mine; ++it) {
      mbar_wait(&bar.m_full[it % CB_NMETA], (it / CB_NMETA) & 1);
      const CbMeta& m = meta[it % CB_NMETA];
      // the tile's (gradient, position) pairs are fetched ONCE, all loads in flight together, and reused by the six steps
      GradCache gc;
      grad_cache_load(gc, dcfeat, cidx, m.s0, m.ns, KC, et);
      for (int j = 0; j < 3; ++j)
        for (int h = 0; h < 2; ++h, ++q) {
          const int gb = q & 1;
          if (q >= 2) mbar_wait(&bar.g_free[gb], ((q >> 1) - 1) & 1);      // the MMAs of step q-2 have read this buffer
          grad_scatter_step(gim + gb * CB_GHALF, m, gc, dcfeat, cidx, KC, j, h, et, 2);
          mbar_arrive(&bar.g_ready[gb]);
        }
    }
    if (n_mine > 0) {
      mbar_wait(&bar.w_full, 0);
      tc_fence_after();
      const int q4 = warp & 3, c = q4 * 32 + lane;                 // TMEM lane = channel
#pragma unroll 1
      for (int j = 0; j < 3; ++j)
#pragma unroll 1
        for (int k0 = 0; k0 < 128; k0 += 32) {
          float v[32];
          tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + j * 128 + k0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + i < KC) atomicAdd(&dw[((size_t)(k0 + i) * D + c) * 3 + j], v[i]);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

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

struct CxBars { uint64_t m_full[CB_NMETA], g_ready[2], g_free[2], w_full[2], w_empty[2], acc_full[2], acc_empty[2]; };

__global__ void __launch_bounds__(CB_THREADS, 1) cnet_conv_bwd_dx_tc_kernel(const float* __restrict__ dcfeat, const int* __restrict__ cidx,
                                                                            const unsigned char* __restrict__ wimg, const int* __restrict__ tso,
                                                                            const int* __restrict__ cstc, int n_tiles, int L, int KC,
                                                                            float* __restrict__ dx) {
  extern __shared__ unsigned char raw[];
  __shared__ CxBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ CbMeta meta[CB_NMETA];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* gim = base;
  unsigned char* wsm = base + CB_GIMG;             // 2 stages of one tap's weights
  float* stg = reinterpret_cast<float*>(base + CB_GIMG + 2 * CX_WIMG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n_mine = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_mine;

  if (tid == 0) {
    for (int s = 0; s < CB_NMETA; ++s) mbar_init(&bar.m_full[s], 128);
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.g_ready[s], 128); mbar_init(&bar.g_free[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar.w_full[s], 1); mbar_init(&bar.w_empty[s], 1); mbar_init(&bar.acc_full[s], 1); mbar_init(&bar.acc_empty[s], 128); }
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ tile bookkeeping + gradient scatter
    const int et = tid;
    int q = 0;
    for (int it = 0; it < n_mine; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      // meta slot it % 4: the epilogue of tile it-4 finished before the accumulator ring let tile it-2's MMAs start, and the last of
      // those has retired (g_free of step q-2) before the bookkeeping below is written - so the slot is free
      CbMeta& m = meta[it % CB_NMETA];
      if (q >= 2) mbar_wait(&bar.g_free[q & 1], ((q >> 1) - 1) & 1);
      m.rowsrc[tid] = -1;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int s0 = tso[tile], ns = tso[tile + 1] - s0;
      if (tid < ns) {
        const int c0 = cstc[s0], b = cstc[s0 + tid] - c0, len = cstc[s0 + tid + 1] - cstc[s0 + tid] - 2;
        m.sb[tid] = b; m.len[tid] = len;
        const int g0 = (s0 + tid) * L;
        for (int l = 0; l < len; ++l) m.rowsrc[b + 1 + l] = g0 + l;
      }
      if (tid == 0) { m.s0 = s0; m.ns = ns; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_arrive(&bar.m_full[it % CB_NMETA]);
      GradCache gc;
      grad_cache_load(gc, dcfeat, cidx, s0, ns, KC, et);
      for (int j = 0; j < 3; ++j)
        for (int h = 0; h < 2; ++h, ++q) {
          const int gb = q & 1;
          if (q >= 2) mbar_wait(&bar.g_free[gb], ((q >> 1) - 1) & 1);
          grad_scatter_step(gim + gb * CB_GHALF, m, gc, dcfeat, cidx, KC, j, h, et, 1);
          mbar_arrive(&bar.g_ready[gb]);
        }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer (tap weights) + MMA issuer
    if (n_mine > 0) {       // whole warp converged, the elected lane issues
      const uint32_t el = elect_one_sync();
      constexpr uint32_t idesc = idesc_bf16(128, 128);              // A (G) and B (W_j^T) K-major
      const uint32_t g0 = smem_u32(gim);
      const int n_taps = 3 * n_mine;
      auto fetch = [&](int p) {                                     // weights of tap p % 3 into stage p & 1
        const int ws = p & 1;
        if (p >= 2) mbar_wait(&bar.w_empty[ws], ((p >> 1) - 1) & 1);
        mbar_arrive_expect_tx_e(el, &bar.w_full[ws], CX_WIMG);
        bulk_copy_g2s_e(el, wsm + ws * CX_WIMG, wimg + (size_t)(p % 3) * CX_WIMG, CX_WIMG, &bar.w_full[ws]);
      };
      fetch(0);
      int q = 0, p = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int acc = it & 1;
        if (it >= 2) mbar_wait(&bar.acc_empty[acc], ((it >> 1) - 1) & 1);
        for (int j = 0; j < 3; ++j, ++p) {
          if (p + 1 < n_taps) fetch(p + 1);
          const int ws = p & 1;
          mbar_wait(&bar.w_full[ws], (p >> 1) & 1);
          const uint32_t w0 = smem_u32(wsm + ws * CX_WIMG);
          for (int h = 0; h < 2; ++h, ++q) {
            const int gb = q & 1;
            mbar_wait(&bar.g_ready[gb], (q >> 1) & 1);
            tc_fence_after();
            const uint32_t gq0 = g0 + gb * CB_GHALF;
            const uint64_t gh = smem_desc_sw128(gq0), gl = smem_desc_sw128(gq0 + 16384);
            const uint64_t wh = smem_desc_sw128(w0 + h * 32768), wl = smem_desc_sw128(w0 + h * 32768 + 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t o = (uint64_t)(kk * 2);
              umma_bf16_e(el, tmem + acc * 128, gh + o, wh + o, idesc, (j | h | kk) != 0);
              umma_bf16_e(el, tmem + acc * 128, gh + o, wl + o, idesc, 1);
              umma_bf16_e(el, tmem + acc * 128, gl + o, wh + o, idesc, 1);
            }
            umma_commit_e(el, &bar.g_free[gb]);
          }
          umma_commit_e(el, &bar.w_empty[ws]);
          if (j == 2) umma_commit_e(el, &bar.acc_full[acc]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: valid rows of dx, coalesced 128-byte row stores
    const int q4 = warp & 3;
    float* sw = stg + q4 * 32 * CX_STG_LD;
    for (int it = 0; it < n_mine; ++it) {
      const int acc = it & 1;
      mbar_wait(&bar.m_full[it % CB_NMETA], (it / CB_NMETA) & 1);
      const CbMeta& m = meta[it % CB_NMETA];
      mbar_wait(&bar.acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + acc * 128 + c0, v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(&sw[lane * CX_STG_LD + jj * 4]) = make_float4(v[jj * 4], v[jj * 4 + 1], v[jj * 4 + 2], v[jj * 4 + 3]);
        __syncwarp();
        const int c = c0 + (lane & 7) * 4;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int rr = jj * 4 + (lane >> 3), src = m.rowsrc[q4 * 32 + rr];
          if (src >= 0) *reinterpret_cast<float4*>(dx + (size_t)src * D + c) = *reinterpret_cast<const float4*>(&sw[rr * CX_STG_LD + (lane & 7) * 4]);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&bar.acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
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
  constexpr int smem = 2 * CB_XIMG + CB_GIMG + 1024;
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_bwd_dw_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_bwd_dw_tc_kernel<<<grid, CB_THREADS, smem, (cudaStream_t)stream>>>(x, dcfeat, cidx, table, table + n_tiles + 1, n_tiles, L, KC, d_conv_w);
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
  constexpr int smem = CB_GIMG + 2 * CX_WIMG + 4 * 32 * CX_STG_LD * 4 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(cnet_conv_bwd_dx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("cnet_conv_bwd_dx_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  if (n_ctas < 1) n_ctas = 148;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  cnet_conv_bwd_dx_tc_kernel<<<grid, CB_THREADS, smem, (cudaStream_t)stream>>>(dcfeat, cidx, reinterpret_cast<const unsigned char*>(wimg_scratch),
                                                                              table, table + n_tiles + 1, n_tiles, L, KC, dx);
  return check_launch("cnet_conv_bwd_dx_tc");
}
