// ImprovedRnn forward on the tensor cores (reference src/model.py:18-21, nn.GRU cell of model.py:19):
// input projection AND recurrence of the variable-length bidirectional GRU in ONE persistent tcgen05 kernel.
//
//   * one CTA per SM, bound to one direction; W_ih (with the bias column) and W_hh stay resident in shared memory for the
//     whole launch as bf16 hi/lo images (3xBF16: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM);
//   * a CTA runs TWO tiles of 128 length-sorted sequences at once ("slots"), each with its own 256 TMEM columns
//       [0,64) r   [64,128) z   (x-part and h-part accumulated together)   [128,192) W_in x + b_in   [192,256) W_hn h
//     so that the gate math of one slot (CUDA cores / MUFU) overlaps the MMAs of the other (tensor pipe);
//   * per time step and slot:  D  = x_t · W_ih^T          (N=192, K=64: 12 MMAs, independent of h)
//                              D += h_{t-1} · W_hh^T      (N=128 into r,z and N=64 into its own n columns: 24 MMAs)
//     then 8 gate warps read their TMEM lanes, apply the ATen GRU cell, mask each row by its own length, write the
//     ImprovedRnn output row in the reference's doubly-permuted order (zeros beyond the length) and h_t as the next step's bf16
//     hi/lo A-operand; in training the driver also streams every h image out (hq, one TMA bulk store per step): together with
//     the token images it is ALL the backward kernel needs - the gates are recomputed there, nothing else is saved;
//   * warp roles: 0-7 gates of slot 0, 8-15 gates of slot 1 (TMEM lane quarter = warp%4, hidden half = (warp/4)%2); warps 16, 17:
//     one driver thread per slot (TMA bulk copy of the pre-split bf16 hi/lo token image, then the step's tcgen05.mma);
//     hand-offs are mbarriers, no __syncthreads in the steady state, the two slots never wait for each other;
//   * several ImprovedRnn calls that share weights (user+item in R-Net, ui+user+item in C-Net) run as "segments" of one
//     launch; the host orders tiles longest-first into per-slot queues (plan.py) so slots finish together.
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc.cuh"
#include "gru_tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int RT_GATE_WARPS = 16;                // 8 per slot
constexpr int RT_THREADS = (RT_GATE_WARPS + 2) * 32;   // 576 (18 warps are allocated as 20: 96 registers per thread)
constexpr int RT_W_BYTES = 2 * G3 * 128;         // hi | lo, [192][64 bf16]
constexpr int RT_A_BYTES = 2 * RT_R * 128;       // hi | lo, [128][64 bf16]
constexpr int RT_SMEM = 2 * RT_W_BYTES + 4 * RT_A_BYTES + 1024;

#ifdef UMPR_TRACE
#define TRACE(role, X, n, ev) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (n) < 64) a.trace[(((role) * 2 + (X)) * 64 + (n)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TRACE(role, X, n, ev) do {} while (0)
#endif

struct RecSeg {
  const unsigned char* xq; const int* plan; float* out; float* hn; unsigned char* hq;
  int n_tiles, n_slabs, N, L, tile_base;
};
struct RecArgs {
  RecSeg seg[RT_MAX_SEG];
  int n_seg;
  const int* q_off; const int* q_tile;
  const float* w[8];
  int E, kx;
  long long* trace;
};

// per-slot state of one gate thread: its row of the tile (32 of the 64 hidden units)
struct GateRow {
  float h[32];
  int len, rowo;
};

constexpr float K_RZ = -1.4426950408889634f;    // -log2(e): sigmoid(a) = 1 / (1 + 2^(K_RZ a))
constexpr float K_N = 2.8853900817779268f;      // 2 log2(e): tanh(a) = 1 - 2 / (1 + 2^(K_N a))

// One time step of one slot for one gate thread.  The resident weights are pre-scaled (r,z rows by K_RZ, n rows and b_hn by
// K_N), so the accumulator columns already hold the exponents.  Straight-line code: every row is computed, rows beyond their
// length keep their state through one select (their accumulator rows are garbage but never observed).
template <bool TRAIN>
__device__ __forceinline__ void gate_step(const int X, const RecArgs& a, const Cur& c, GateRow& g, int n, int dir, int row, int hf,
                                          unsigned char* hs, const float* s_bhn, uint64_t* h_ready, uint64_t* acc_full, uint32_t tmem) {
  const RecSeg& sg = a.seg[c.si];
  const int u0 = hf * 32;
  const int lane = row & 31;
  unsigned char* h_hi = hs + X * RT_A_BYTES, *h_lo = h_hi + RT_R * 128;
  if (c.s == 0) {
    // tile start: this row's job, h_0 = 0 (registers and the A-operand image)
    const int Rp = sg.n_tiles * RT_R, k = c.tile * RT_R + row;
    g.rowo = sg.plan[Rp + k];
    g.len = sg.plan[2 * Rp + k];
#pragma unroll
    for (int i = 0; i < 32; ++i) g.h[i] = 0.f;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const uint32_t off = (uint32_t)(row * 128 + (((hf * 4 + cc) ^ (row & 7)) << 4));
      *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();          // orders this thread's TMEM reads of the slot's previous tile before the next MMAs
    mbar_arrive(&h_ready[X]);
  }
  const int t = dir ? (c.Lj - 1 - c.s) : c.s;
  const bool live = t < g.len;
  const uint32_t trow = tmem + ((uint32_t)((row >> 5) * 32) << 16) + X * 256 + u0;

  if ((threadIdx.x & 255) == 0) TRACE(0, X, n, 0);
  mbar_wait(&acc_full[X], n & 1);
  tc_fence_after();
  if ((threadIdx.x & 255) == 0) TRACE(0, X, n, 1);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t v[32];
    tmem_ld8_issue(trow + cc * 8, v);
    tmem_ld8_issue(trow + 64 + cc * 8, v + 8);
    tmem_ld8_issue(trow + 128 + cc * 8, v + 16);
    tmem_ld8_issue(trow + 192 + cc * 8, v + 24);
    const float4 b0 = *reinterpret_cast<const float4*>(s_bhn + u0 + cc * 8), b1 = *reinterpret_cast<const float4*>(s_bhn + u0 + cc * 8 + 4);
    const float bh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    tmem_ld_wait();
    float hv[8];
    const f2 one = f2_set(1.f, 1.f), mone = f2_set(-1.f, -1.f), two = f2_set(2.f, 2.f), mtwo = f2_set(-2.f, -2.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // two hidden units per instruction (packed fp32); with p = 1 / (1 + e^{2 a_n}):  n = tanh(a_n) = 1 - 2p
      const int i = 2 * j;
      const f2 er = f2_ex2(f2_bits(v[i], v[i + 1]));                                   // e^{-a_r}
      const f2 r = f2_rcp(f2_add(er, one));                                             // sigmoid
      const f2 hcand = f2_add(f2_bits(v[24 + i], v[25 + i]), f2_set(bh[i], bh[i + 1])); // K_N (W_hn h + b_hn)
      const f2 xn = f2_fma(r, hcand, f2_bits(v[16 + i], v[17 + i]));                    // K_N (n pre-activation)
      const f2 ez = f2_ex2(f2_min(f2_bits(v[8 + i], v[9 + i]), 60.f));                  // e^{-a_z}
      const f2 en = f2_ex2(f2_min(xn, 60.f));                                           // e^{2 a_n}
      const f2 dz = f2_add(ez, one), dn = f2_add(en, one);
      const f2 inv = f2_rcp(f2_mul(dz, dn));                                            // one reciprocal for z and n
      const f2 z = f2_mul(dn, inv);
      const f2 pp = f2_mul(dz, inv);
      const f2 nv = f2_fma(pp, mtwo, one);
      const f2 hold = f2_set(g.h[cc * 8 + i], g.h[cc * 8 + i + 1]);
      const f2 hmn = f2_fma(pp, two, f2_add(hold, mone));                               // h - n
      const f2 hnew = f2_fma(hmn, z, nv);                                               // ATen GRU cell: (h - n) * z + n
      hv[i] = live ? hnew.x : hold.x;
      hv[i + 1] = live ? hnew.y : hold.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g.h[cc * 8 + i] = hv[i];
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(hv[2 * i], hv[2 * i + 1], hi[i], lo[i]);
    const uint32_t off = (uint32_t)(row * 128 + (((hf * 4 + cc) ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if ((threadIdx.x & 255) == 0) TRACE(0, X, n, 2);
  if (c.s + 1 < c.Lj) {
    fence_async_smem();        // h_t image visible to the tensor core's shared-memory reads
    tc_fence_before();
    mbar_arrive(&h_ready[X]);
  }
  {
    // ImprovedRnn output row (row_of[job], t): h_t for live rows, zeros for t >= length (pad_packed_sequence, model.py:20).
    // Transposed inside groups of 8 lanes so that every store instruction writes 4 rows x 128 contiguous bytes.
    float4 o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o[k] = live ? make_float4(g.h[4 * k], g.h[4 * k + 1], g.h[4 * k + 2], g.h[4 * k + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    transpose8x8_f4(o, lane);
    const int my = g.rowo < 0 ? -1 : g.rowo * sg.L + t;
    float* obase = sg.out + dir * H + u0 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ri = __shfl_sync(0xffffffffu, my, (lane & 24) + i);
      if (ri >= 0) *reinterpret_cast<float4*>(obase + (size_t)ri * D) = o[i];
    }
  }
  if ((threadIdx.x & 255) == 0) TRACE(0, X, n, 3);
  if (c.s + 1 == c.Lj) {
    // tile end: zero padding up to total_length (model.py:17,20) and h_n in ORIGINAL sequence order
    float* obase = sg.out + dir * H + u0 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ro = __shfl_sync(0xffffffffu, g.rowo, (lane & 24) + i);
      if (ro < 0) continue;
      for (int tt = c.Lj; tt < sg.L; ++tt)
        *reinterpret_cast<float4*>(obase + ((size_t)ro * sg.L + tt) * D) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (g.rowo >= 0 && sg.hn) {
      const int seq = sg.plan[c.tile * RT_R + row];
      float* hp = sg.hn + ((size_t)dir * sg.N + seq) * H + u0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(hp + i * 4) = make_float4(g.h[i * 4], g.h[i * 4 + 1], g.h[i * 4 + 2], g.h[i * 4 + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// embedding gather + length-aware pack straight into tensor-core operand images (model.py:262-264 + the pack half of :18):
//   xq[slab][hi|lo][128 rows][64 bf16]   K-major, SWIZZLE_128B: row r at byte r*128, its 16-byte chunk c at chunk c ^ (r & 7)
// columns: E embedding values, a 1.0 (bias column), zeros.  fp32 = hi + lo (bf16 each, "3xBF16" operands).  The same image is
// the A operand of the recurrence's x-part MMA and the (token-major) B operand of the weight-gradient MMA.
// ------------------------------------------------------------------------------------------------
// up to three review sides in one launch (the native step: user, item, user->item): blocks [slab0[k], slab0[k+1]) pack side k
struct GatherSides {
  const int64_t* ids[3]; Plan plan[3]; unsigned char* xq[3]; int L[3]; int slab0[4]; int n;
};
__global__ void __launch_bounds__(256) gather_pack_tc_kernel(const float* __restrict__ table, const float* __restrict__ dense, GatherSides gs,
                                                             int E) {
  int side = 0;
  if (gs.n > 1 && (int)blockIdx.x >= gs.slab0[1]) side = (gs.n > 2 && (int)blockIdx.x >= gs.slab0[2]) ? 2 : 1;
  const Plan& p = gs.plan[side];
  const int64_t* __restrict__ ids = gs.ids[side];
  unsigned char* __restrict__ xq = gs.xq[side];
  const int L = gs.L[side];
  const int sl = blockIdx.x - gs.slab0[side];
  const int j = p.slab_tile[sl];
  const int t = sl - p.tile_off[j];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec2 = (E & 1) == 0;
  unsigned char* img = xq + (size_t)sl * RT_A_BYTES;
  // The dependent chain plan -> token id -> table row is pure latency: a warp owns 16 rows (warp + 8q) and walks the chain for all
  // of them at once - lane q resolves row q's table offset (one plan / id load per lane instead of a broadcast load per row), the
  // offsets are handed round by shuffles, and the 16 row loads are outstanding together.
  long long mine = -1;
  {
    const int k = j * RT_R + warp + 8 * (lane & 15);
    if (t < p.len_of[k]) {
      const size_t tok = (size_t)p.seq_of[k] * L + t;
      mine = (long long)(dense ? tok : (size_t)ids[tok]) * E;
    }
  }
  const float* base = dense ? dense : table;
  const int e0 = 2 * lane;
  float2 v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const long long o = __shfl_sync(0xffffffffu, mine, q);
    v[q] = make_float2(0.f, 0.f);
    if (o >= 0) {
      const float* src = base + o;
      if (vec2 && e0 + 1 < E) {
        v[q] = *reinterpret_cast<const float2*>(src + e0);
      } else {
        if (e0 < E) v[q].x = src[e0]; else if (e0 == E) v[q].x = 1.f;
        if (e0 + 1 < E) v[q].y = src[e0 + 1]; else if (e0 + 1 == E) v[q].y = 1.f;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int r = warp + 8 * q;
    uint32_t hi, lo;
    split2(v[q].x, v[q].y, hi, lo);
    const uint32_t off = sw128_off(r, 2 * lane);
    *reinterpret_cast<uint32_t*>(img + off) = hi;
    *reinterpret_cast<uint32_t*>(img + RT_R * 128 + off) = lo;
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(RT_THREADS, 1) gru_fwd_tc_kernel(const __grid_constant__ RecArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t x_full[2], x_empty[2], h_ready[2], acc_full[2], stagger;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bhn[H];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wih = base;                              // [hi|lo][192][128 B]
  unsigned char* whh = base + RT_W_BYTES;
  unsigned char* hs = base + 2 * RT_W_BYTES;              // [slot][hi|lo][128][128 B]
  unsigned char* xs = hs + 2 * RT_A_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&h_ready[s], RT_GATE_WARPS * 16);
      mbar_init(&acc_full[s], 1);
    }
    mbar_init(&stagger, RT_GATE_WARPS * 16);
    mbar_fence_init();
  }
  if (warp == RT_GATE_WARPS) tmem_alloc(&tmem_slot, 512);
  {
    // resident weights of this direction, pre-scaled so the accumulators are base-2 exponents (see gate_step):
    // W_ih with column E = b_ih (+ b_hh for r,z) (multiplied by the packed tokens' 1.0 column), W_hh, b_hn
    const float* w_ih = a.w[dir * 4 + 0], *w_hh = a.w[dir * 4 + 1], *b_ih = a.w[dir * 4 + 2], *b_hh = a.w[dir * 4 + 3];
    for (int idx = tid; idx < G3 * 16; idx += RT_THREADS) {
      const int n = idx >> 4, k = (idx & 15) * 4;
      const float sc = n < 2 * H ? K_RZ : K_N;
      float t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = k + q;
        t[q] = sc * (kk < a.E ? w_ih[(size_t)n * a.E + kk] : (kk == a.E ? b_ih[n] + (n < 2 * H ? b_hh[n] : 0.f) : 0.f));
      }
      store_split4(wih, wih + G3 * 128, n, k, make_float4(t[0], t[1], t[2], t[3]));
      const float4 wh = *reinterpret_cast<const float4*>(w_hh + n * H + k);
      store_split4(whh, whh + G3 * 128, n, k, make_float4(sc * wh.x, sc * wh.y, sc * wh.z, sc * wh.w));
    }
    if (tid < H) s_bhn[tid] = K_N * b_hh[2 * H + tid];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < RT_GATE_WARPS) {
    // ------------------------------------------------------------------ gate warps: warps 0-7 serve slot 0, warps 8-15 slot 1
    const int X = warp >> 3;
    const int row = (warp & 3) * 32 + lane, hf = (warp >> 2) & 1;
    GateRow g;
    g.len = 0; g.rowo = -1;
    Cur c;
    cur_init(a, c, 2 * blockIdx.x + X);
    for (int n = 0; c.active; ++n) {
      gate_step<TRAIN>(X, a, c, g, n, dir, row, hf, hs, s_bhn, h_ready, acc_full, tmem);
      if (X == 0 && n == 0) mbar_arrive(&stagger);
      cur_next(a, c);
    }
  } else {
    // ------------------------------------------------------------------ slot drivers (warp 16: slot 0, warp 17: slot 1): TMA producer of the
    // packed-token image AND tcgen05.mma issuer of its slot.  The whole warp runs this code converged so that descriptors and
    // addresses stay in uniform registers; only the elected lane's predicate lets the TMA / MMA instructions through (tc.cuh)
    const uint32_t el = elect_one_sync();
    const int X = warp - RT_GATE_WARPS;
    constexpr uint32_t id192 = idesc_bf16(128, 192), id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
    const uint64_t wih_h = smem_desc_sw128(smem_u32(wih)), wih_l = smem_desc_sw128(smem_u32(wih + G3 * 128));
    const uint64_t whh_h = smem_desc_sw128(smem_u32(whh)), whh_l = smem_desc_sw128(smem_u32(whh + G3 * 128));
    const uint64_t whn_h = smem_desc_sw128(smem_u32(whh + 128 * 128)), whn_l = smem_desc_sw128(smem_u32(whh + G3 * 128 + 128 * 128));
    unsigned char* xbuf = xs + X * RT_A_BYTES;
    const uint64_t x_h = smem_desc_sw128(smem_u32(xbuf)), x_l = smem_desc_sw128(smem_u32(xbuf + RT_R * 128));
    const uint64_t h_h = smem_desc_sw128(smem_u32(hs + X * RT_A_BYTES)), h_l = smem_desc_sw128(smem_u32(hs + X * RT_A_BYTES + RT_R * 128));
    const uint32_t d = tmem + X * 256;
    Cur c;
    cur_init(a, c, 2 * blockIdx.x + X);
    auto load_x = [&]() {       // the token image is already bf16 hi|lo, SWIZZLE_128B (gather_pack_tc_kernel): 32 KB land as they are
      const RecSeg& sg = a.seg[c.si];
      const int t = dir ? (c.Lj - 1 - c.s) : c.s;
      const int slab = c.slab0 + t;
      mbar_arrive_expect_tx_e(el, &x_full[X], RT_A_BYTES);
      bulk_copy_g2s_e(el, xbuf, sg.xq + (size_t)slab * RT_A_BYTES, RT_A_BYTES, &x_full[X]);
    };
    if (c.active) load_x();
    // start slot 1 half a period late: its MMAs then run under slot 0's gate math and vice versa (anti-phase is self-sustaining)
    if (X == 1 && c.active) mbar_wait(&stagger, 0);
    for (int n = 0; c.active; ++n) {
      if (el) TRACE(1, X, n, 0);
      mbar_wait(&x_full[X], n & 1);
      if (el) TRACE(1, X, n, 1);
      mbar_wait(&h_ready[X], n & 1);            // h_{t-1} image written AND the slot's accumulator columns drained
      tc_fence_after();
      if (el) TRACE(1, X, n, 2);
      if (TRAIN && c.s > 0) {
        // training: keep h_{t-1} as an operand image (bf16 hi|lo, 32 KB) for the backward kernel: it is the A operand of this
        // step's MMAs, the h_prev of the GRU cell's backward and the token-major operand of the weight-gradient MMA
        const RecSeg& sg = a.seg[c.si];
        const int tprev = dir ? (c.Lj - c.s) : (c.s - 1);
        const size_t slab = (size_t)c.slab0 + tprev;
        bulk_copy_s2g_e(el, sg.hq + (slab * 2 + dir) * RT_A_BYTES, hs + X * RT_A_BYTES, RT_A_BYTES);
      }
      for (int kk = 0; kk < a.kx; ++kk) {       // x_t · W_ih^T  -> r, z, n_x  (overwrites)
        const uint64_t o = (uint64_t)(kk * 2);
        umma_bf16_e(el, d, x_h + o, wih_h + o, id192, kk != 0);
        umma_bf16_e(el, d, x_h + o, wih_l + o, id192, 1);
        umma_bf16_e(el, d, x_l + o, wih_h + o, id192, 1);
      }
      umma_commit_e(el, &x_empty[X]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {          // h_{t-1} · W_hh^T -> += r, z ; n_h (own columns)
        const uint64_t o = (uint64_t)(kk * 2);
        umma_bf16_e(el, d, h_h + o, whh_h + o, id128, 1);
        umma_bf16_e(el, d, h_h + o, whh_l + o, id128, 1);
        umma_bf16_e(el, d, h_l + o, whh_h + o, id128, 1);
        umma_bf16_e(el, d + 192, h_h + o, whn_h + o, id64, kk != 0);
        umma_bf16_e(el, d + 192, h_h + o, whn_l + o, id64, 1);
        umma_bf16_e(el, d + 192, h_l + o, whn_h + o, id64, 1);
      }
      if (TRAIN) bulk_wait_read();              // the h image has been read out before the gate threads may overwrite it
      umma_commit_e(el, &acc_full[X]);
      if (el) TRACE(1, X, n, 3);
      cur_next(a, c);
      if (c.active) {
        mbar_wait(&x_empty[X], n & 1);          // the x-part MMAs of this step have consumed the buffer
        load_x();
      }
    }
    if (TRAIN) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RT_GATE_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_gather_pack_tc(const float* table, const int64_t* ids, const float* dense, const int32_t* plan, int n_tiles,
                                   int n_slabs, int L, int E, void* xq, void* stream) {
  if ((!table || !ids) && !dense) return fail_arg("gather_pack_tc: need (table, ids) or dense");
  if (E < 1 || E >= KP) return fail_arg("gather_pack_tc: embedding width E=%d must be in [1, %d)", E, KP);
  if (n_slabs == 0) return 0;
  GatherSides gs{};
  gs.n = 1; gs.ids[0] = ids; gs.plan[0] = make_plan(plan, n_tiles, n_slabs, RT_R); gs.xq[0] = reinterpret_cast<unsigned char*>(xq); gs.L[0] = L;
  gs.slab0[0] = 0; gs.slab0[1] = n_slabs;
  gather_pack_tc_kernel<<<n_slabs, 256, 0, (cudaStream_t)stream>>>(table, dense, gs, E);
  return check_launch("gather_pack_tc");
}

namespace umpr {
// umpr_gather_pack_tc for up to three sides in ONE launch (csrc/step.cu): same table, one plan / id tensor / image buffer per side
int gather_pack_tc_sides(const float* table, int n_sides, const int64_t* const* ids, const int32_t* const* plan, const int* n_tiles,
                         const int* n_slabs, const int* L, int E, void* const* xq, void* stream) {
  if (!table) return fail_arg("gather_pack_tc: table is NULL");
  if (n_sides < 1 || n_sides > 3) return fail_arg("gather_pack_tc: %d sides", n_sides);
  if (E < 1 || E >= KP) return fail_arg("gather_pack_tc: embedding width E=%d must be in [1, %d)", E, KP);
  GatherSides gs{};
  gs.n = n_sides;
  int total = 0;
  for (int k = 0; k < n_sides; ++k) {
    gs.ids[k] = ids[k]; gs.plan[k] = make_plan(plan[k], n_tiles[k], n_slabs[k], RT_R); gs.xq[k] = reinterpret_cast<unsigned char*>(xq[k]); gs.L[k] = L[k];
    gs.slab0[k] = total;
    total += n_slabs[k];
  }
  for (int k = n_sides; k < 4; ++k) gs.slab0[k] = total;
  if (total == 0) return 0;
  gather_pack_tc_kernel<<<total, 256, 0, (cudaStream_t)stream>>>(table, nullptr, gs, E);
  return check_launch("gather_pack_tc");
}
}  // namespace umpr

extern "C" int umpr_gru_fwd_tc(const umpr_gru_seg* segs, int n_seg, const float* const* w, int E, const int32_t* sched,
                               int n_queues, void* stream) {
  if (n_seg < 1 || n_seg > RT_MAX_SEG) return fail_arg("gru_fwd_tc: n_seg=%d not in [1,%d]", n_seg, RT_MAX_SEG);
  if (E < 1 || E >= KP) return fail_arg("gru_fwd_tc: E=%d must be in [1,%d)", E, KP);
  if (n_queues < 2 || (n_queues & 1)) return fail_arg("gru_fwd_tc: n_queues=%d must be even and >= 2", n_queues);
  RecArgs a{};
  int base = 0;
  for (int i = 0; i < n_seg; ++i) {
    const umpr_gru_seg& s = segs[i];
    if (s.n_tiles < 1 || s.L < 1 || !s.xq || !s.plan || !s.out) return fail_arg("gru_fwd_tc: segment %d is incomplete", i);
    a.seg[i] = RecSeg{reinterpret_cast<const unsigned char*>(s.xq), s.plan, s.out, s.hn, reinterpret_cast<unsigned char*>(s.hq), s.n_tiles, s.n_slabs, s.N, s.L, base};
    base += s.n_tiles;
  }
  a.n_seg = n_seg;
  a.q_off = sched;
  a.q_tile = sched + n_queues + 1;
  for (int i = 0; i < 8; ++i) a.w[i] = w[i];
  a.E = E;
  a.kx = (E + 1 + 15) / 16;
#ifdef UMPR_TRACE
  static long long* tr = nullptr;
  if (!tr) cudaMalloc(&tr, 3 * 2 * 64 * 4 * sizeof(long long));
  cudaMemsetAsync(tr, 0, 3 * 2 * 64 * 4 * sizeof(long long), (cudaStream_t)stream);
  a.trace = tr;
#else
  a.trace = nullptr;
#endif
  bool train = segs[0].hq != nullptr;
  for (int i = 0; i < n_seg; ++i)
    if ((segs[i].hq != nullptr) != train)
      return fail_arg("gru_fwd_tc: hq must be given for all segments (training) or for none (inference)");
  auto kern = train ? gru_fwd_tc_kernel<true> : gru_fwd_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM);
  if (e != cudaSuccess) { set_error("gru_fwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  kern<<<dim3(n_queues / 2, 2), RT_THREADS, RT_SMEM, (cudaStream_t)stream>>>(a);
#ifdef UMPR_TRACE
  {
    cudaStreamSynchronize((cudaStream_t)stream);
    static long long host[3 * 2 * 64 * 4];
    cudaMemcpy(host, tr, sizeof(host), cudaMemcpyDeviceToHost);
    FILE* f = fopen("gpurun_out/gru_trace.txt", "w");
    if (f) {
      long long t0 = host[(1 * 2 + 0) * 64 * 4];
      for (int role = 0; role < 3; ++role) for (int X = 0; X < 2; ++X) for (int n = 0; n < 24; ++n) {
        fprintf(f, "role %d slot %d step %2d:", role, X, n);
        for (int ev = 0; ev < 4; ++ev) fprintf(f, " %8lld", host[((role * 2 + X) * 64 + n) * 4 + ev] ? host[((role * 2 + X) * 64 + n) * 4 + ev] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#endif
  return check_launch("gru_fwd_tc");
}
