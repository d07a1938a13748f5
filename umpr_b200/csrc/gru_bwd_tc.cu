// ImprovedRnn backward on the tensor cores: reverse-time recurrence AND the GRU weight gradients in one persistent kernel
// (backward of reference src/model.py:19-21; main.py:36).  NOTHING but the operand images of the forward is read: the gates are
// RECOMPUTED from x_t and h_{t-1} with the forward's own MMAs, and the gate gradients never leave the SM.
//
// Per CTA (one direction, one 128-sequence tile at a time, same tile queues as the forward), per time step in reverse order:
//     [r, z, n_x, n_h] = [x_t | h_{t-1}] · [W_ih | W_hh]^T          (the forward's 36 MMAs, bit-identical accumulators)
//     dh_t   = dy_t + dh_{t+1} (.) z_{t+1} + [dr, dz, dn*r]_{t+1} · W_hh                         (carry)
//     dn_pre = dh (1-z)(1-n^2)   dz_pre = dh (h_{t-1} - n) z (1-z)   dr_pre = dn_pre (W_hn h + b_hn) r (1-r)
//     dW^T[feature][gate] += [x_t | h_{t-1}]^T · [dr, dz, dn*r, dn]                                 (weight gradients)
//   * shared memory (225 KB): W_ih and W_hh images (bf16 hi|lo, pre-scaled by -log2e / 2 log2e exactly as in the forward), the
//     step's token image xq[t] and hidden image hq[t-1] (ONE TMA bulk copy each - they serve K-major as the A operand of the
//     recomputation and MN-major as the A operand of the weight-gradient product), and a 64 KB gate-gradient tile;
//   * the gate-gradient tile (token-major bf16 hi|lo, two blocks of 64 gates: [dr | dz] and [dn*r | dn] of 32 hidden units) is
//     BOTH the K-major A operand of the carry product (B = the resident W_hh image read MN-major) and the MN-major B operand of
//     the weight-gradient product; its values carry the inverse of the weight pre-scaling, which the final flush undoes.  It is
//     written twice per step - hidden units 0..31, then 32..63 - so that the MMAs of the first half run under the gate math of
//     the second;
//   * tensor memory (512 columns): [0,64) r | [64,128) z | [128,192) n_x | [192,256) n_h, re-used for dh (the carry product lands
//     where the next step's W_hn h will: the gate threads fold dh into their registers first) | [256,512) dW^T accumulator, kept
//     over the CTA's whole queue and flushed once with atomics into the eight nn.GRU gradients;
//   * warps 0-15: gate threads (one sequence row x 16 hidden units each: 5 MUFU per unit for the recomputed cell, so 16 warps
//     keep the XU pipe fed); warp 16: driver thread (TMA copies, L2 prefetch of the next step's images, all MMAs);
//   * HBM traffic per token and direction: xq 256 B + hq 256 B + d_out 256 B (the saved-gate tensor of the first design, 1 KB
//     written by the forward and read back here, is gone).
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc.cuh"
#include "gru_tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

// clock64 trace of CTA (0,0) (UMPR_TRACE_BWD=<file> at run time): role 0 = gate thread 0, role 1 = driver; 8 events x 64 steps each
#define BTRACE(role, n, ev) do { if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (n) < 64) a.trace[((role) * 64 + (n)) * 8 + (ev)] = clock64(); } while (0)

constexpr int RB_GATE_WARPS = 16;
constexpr int RB_GATE_THREADS = RB_GATE_WARPS * 32;       // 512
constexpr int RB_THREADS = (RB_GATE_WARPS + 1) * 32;      // 544
constexpr int RB_W_BYTES = 2 * G3 * 128;                  // one weight image: hi | lo, [192][64 bf16] = 48 KB
constexpr int RB_IMG = 2 * RT_R * 128;                    // one operand image: hi | lo, [128][64 bf16] = 32 KB
constexpr int RB_BLK = RT_R * 128;                        // one block of the gate-gradient tile: [128 sequences][64 gates bf16] = 16 KB
constexpr int RB_GT = 4 * RB_BLK;                         // gate-gradient tile: [hi|lo][2 blocks] = 64 KB
constexpr int RB_SMEM = 2 * RB_W_BYTES + 2 * RB_IMG + RB_GT + 1024;

constexpr float RB_K_RZ = -1.4426950408889634f;           // the forward's weight pre-scaling (gru_rec_tc.cu): r, z rows by -log2(e),
constexpr float RB_K_N = 2.8853900817779268f;             // n rows and b_hn by 2 log2(e)

struct BwdSeg {
  const float* d_out; const float* d_hn; const unsigned char* xq; const unsigned char* hq; const int* plan;
  int n_tiles, n_slabs, N, L, tile_base;
};
struct BwdArgs {
  BwdSeg seg[RT_MAX_SEG];
  int n_seg;
  const int* q_off; const int* q_tile;
  const float* w[8];
  float* dw[8];
  const unsigned char* zero_img;      // 32 KB of zeros: h_{t-1} of a sequence's first step
  int E, kx;
  long long* trace;
};

struct BwdRow {
  float part[16];      // dh_{t+1} (.) z_{t+1} + the carry product (+ d_hn at the row's last step): everything of dh_t but dy_t
  int len, rowo;
};

struct BwdBars {
  uint64_t img_q[2];        // the step's operand images land separately: x (hi|lo), h (hi|lo)
  uint64_t p1_full, ta_ready, ta_done, tb_ready, dh_full, step_done, dh_read;
};

// bf16 hi + lo of 8 consecutive units (one 16-byte chunk of each image row) -> fp32
__device__ __forceinline__ void unpack8(const uint4 hi, const uint4 lo, float* v) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
    v[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
  }
}

// Gate gradients of 8 hidden units (u .. u+7) of one row: recompute the cell from the accumulators, back-propagate dh.
// -> packed bf16 hi / lo words of [dr | dz | dn*r | dn] (4 words each), carrying the inverse of the resident weights' pre-scaling.
// v: the unit's four accumulator groups (bwd_cell_load), already waited for.
__device__ __forceinline__ void bwd_cell_load(uint32_t trow, int u, uint32_t (&v)[32]) {
  tmem_ld8_issue(trow + u, v);
  tmem_ld8_issue(trow + 64 + u, v + 8);
  tmem_ld8_issue(trow + 128 + u, v + 16);
  tmem_ld8_issue(trow + 192 + u, v + 24);
}
__device__ __forceinline__ void bwd_cell8(const uint32_t (&v)[32], int u, const unsigned char* himg, uint32_t hoff, const float* s_bhn, const float4 dya,
                                          const float4 dyb, bool live, float* part, uint32_t (&hi)[4][4], uint32_t (&lo)[4][4]) {
  const uint4 hhi = *reinterpret_cast<const uint4*>(himg + hoff);
  const uint4 hlo = *reinterpret_cast<const uint4*>(himg + RT_R * 128 + hoff);
  const float4 b0 = *reinterpret_cast<const float4*>(s_bhn + u), b1 = *reinterpret_cast<const float4*>(s_bhn + u + 4);
  const float bh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float dy[8] = {dya.x, dya.y, dya.z, dya.w, dyb.x, dyb.y, dyb.z, dyb.w};
  float hp[8];
  unpack8(hhi, hlo, hp);
  const f2 one = f2_set(1.f, 1.f), mone = f2_set(-1.f, -1.f), two = f2_set(2.f, 2.f);
  const f2 c4n = f2_set(4.f / RB_K_N, 4.f / RB_K_N), crz = f2_set(1.f / RB_K_RZ, 1.f / RB_K_RZ);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // two hidden units per instruction (packed fp32).  The forward's cell (gru_rec_tc.cu gate_step): with p = 1 / (1 + e^{2 a_n}),
    // n = 1 - 2p and 1 - n^2 = 4 p (1 - p)
    const int i = 2 * j;
    const f2 er = f2_ex2(f2_bits(v[i], v[i + 1]));
    const f2 r = f2_rcp(f2_add(er, one));
    const f2 hcand = f2_add(f2_bits(v[24 + i], v[25 + i]), f2_set(bh[i], bh[i + 1]));   // 2 log2e (W_hn h + b_hn)
    const f2 xn = f2_fma(r, hcand, f2_bits(v[16 + i], v[17 + i]));
    const f2 ez = f2_ex2(f2_min(f2_bits(v[8 + i], v[9 + i]), 60.f));
    const f2 en = f2_ex2(f2_min(xn, 60.f));
    const f2 dzd = f2_add(ez, one), dnd = f2_add(en, one);
    const f2 inv = f2_rcp(f2_mul(dzd, dnd));
    const f2 z = f2_mul(dnd, inv);
    const f2 pp = f2_mul(dzd, inv);
    // backward of the cell.  Rows beyond their length: dh = 0 makes every gate gradient an exact zero (the gates are finite), and
    // the carry just passes through.  With the inverse pre-scaling folded in:
    //   dn' = dh (1-z) 4p(1-p) / (2 log2e)   dnr' = dn' r   dr' = dn' hcand r (1-r) / (-log2e)   dz' = dh (h_prev - n) z (1-z) / (-log2e)
    f2 dh = f2_add(f2_set(part[i], part[i + 1]), f2_set(dy[i], dy[i + 1]));
    if (!live) dh = f2_set(0.f, 0.f);
    const f2 omz = f2_fma(z, mone, one), omp = f2_fma(pp, mone, one), omr = f2_fma(r, mone, one);
    const f2 dnp = f2_mul(f2_mul(dh, c4n), f2_mul(f2_mul(omz, pp), omp));
    const f2 dnrp = f2_mul(dnp, r);
    const f2 drp = f2_mul(f2_mul(dnp, f2_mul(r, omr)), f2_mul(hcand, crz));
    const f2 hmn = f2_fma(pp, two, f2_add(f2_set(hp[i], hp[i + 1]), mone));            // h_prev - n
    const f2 dzp = f2_mul(f2_mul(f2_mul(dh, crz), f2_mul(z, omz)), hmn);
    const f2 pnew = f2_mul(dh, z);
    split2(drp.x, drp.y, hi[0][j], lo[0][j]);
    split2(dzp.x, dzp.y, hi[1][j], lo[1][j]);
    split2(dnrp.x, dnrp.y, hi[2][j], lo[2][j]);
    split2(dnp.x, dnp.y, hi[3][j], lo[3][j]);
    if (live) { part[i] = pnew.x; part[i + 1] = pnew.y; }
  }
}

// One reverse time step of one tile for one gate thread: row = sequence of the tile, ug = which 8 hidden units of each half
// (units ug*8 .. +7 of 0..31, then of 32..63).
__device__ __forceinline__ void bwd_gate_step(const BwdArgs& a, const Cur& c, BwdRow& g, int n, int dir, int row, int ug,
                                              const unsigned char* himg, unsigned char* gt, const float* s_bhn, BwdBars* bar, uint32_t tmem) {
  const BwdSeg& sg = a.seg[c.si];
  const int ua = ug * 8, ub = 32 + ug * 8;              // this thread's units in the two halves
  const int Rp = sg.n_tiles * RT_R;
  if (c.s == 0) {
    const int k = c.tile * RT_R + row;
    g.rowo = sg.plan[Rp + k];
    g.len = sg.plan[2 * Rp + k];
#pragma unroll
    for (int i = 0; i < 16; ++i) g.part[i] = 0.f;
    if (sg.d_hn && g.rowo >= 0) {
      const float* hp = sg.d_hn + ((size_t)dir * sg.N + sg.plan[k]) * H;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(hp + (i < 2 ? ua : ub - 8) + i * 4);
        g.part[4 * i] = v.x; g.part[4 * i + 1] = v.y; g.part[4 * i + 2] = v.z; g.part[4 * i + 3] = v.w;
      }
    }
  }
  const int t = dir ? c.s : (c.Lj - 1 - c.s);          // reverse of the forward kernel's order
  const bool live = t < g.len;
  const bool any_live = __any_sync(0xffffffffu, live);
  const uint32_t trow = tmem + ((uint32_t)((row >> 5) * 32) << 16);
  const uint32_t off0 = (uint32_t)(row * 128);

  // d_out row of this step: requested before the barrier waits, so that its latency hides behind them
  float4 dy4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) dy4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* dyrow = sg.d_out + ((size_t)g.rowo * sg.L + t) * D + dir * H;
    dy4[0] = *reinterpret_cast<const float4*>(dyrow + ua);
    dy4[1] = *reinterpret_cast<const float4*>(dyrow + ua + 4);
    dy4[2] = *reinterpret_cast<const float4*>(dyrow + ub);
    dy4[3] = *reinterpret_cast<const float4*>(dyrow + ub + 4);
  }
  const bool tr = threadIdx.x == 0;
  if (tr) BTRACE(0, n, 0);
  if (n > 0) {
    // the previous step's carry product [dr, dz, dn*r] · W_hh: fold it into the running part, then release its columns
    mbar_wait(&bar->dh_full, (n - 1) & 1);
    tc_fence_after();
    if (c.s > 0) {                                      // (a tile's first step: what is there belongs to the previous tile)
      uint32_t acc[16];
      tmem_ld8_issue(trow + 192 + ua, acc);
      tmem_ld8_issue(trow + 192 + ub, acc + 8);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) g.part[i] += __uint_as_float(acc[i]);
    }
    tc_fence_before();
    mbar_arrive(&bar->dh_read);
  }
  if (tr) BTRACE(0, n, 1);
  mbar_wait(&bar->img_q[1], n & 1);                     // h_{t-1} image readable (TMA write)
  if (tr) BTRACE(0, n, 2);
  mbar_wait(&bar->p1_full, n & 1);                      // recomputed accumulators complete
  tc_fence_after();
  if (tr) BTRACE(0, n, 3);

  // tile rows: block 0 = [dr (32 units of the half) | dz (32)], block 1 = [dn*r | dn]; this thread's 8 units are 16-byte chunk ug
  // of the first gate type and chunk 4 + ug of the second (chunks are XOR-swizzled with the row, SWIZZLE_128B)
  const uint32_t o1 = off0 + ((uint32_t)(ug ^ (row & 7)) << 4), o2 = off0 + ((uint32_t)((4 + ug) ^ (row & 7)) << 4);
  auto store_tile = [&](const uint32_t (&hi)[4][4], const uint32_t (&lo)[4][4]) {
    *reinterpret_cast<uint4*>(gt + o1) = make_uint4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);                     // dr
    *reinterpret_cast<uint4*>(gt + o2) = make_uint4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);                     // dz
    *reinterpret_cast<uint4*>(gt + RB_BLK + o1) = make_uint4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);            // dn*r
    *reinterpret_cast<uint4*>(gt + RB_BLK + o2) = make_uint4(hi[3][0], hi[3][1], hi[3][2], hi[3][3]);            // dn
    *reinterpret_cast<uint4*>(gt + 2 * RB_BLK + o1) = make_uint4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]);
    *reinterpret_cast<uint4*>(gt + 2 * RB_BLK + o2) = make_uint4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
    *reinterpret_cast<uint4*>(gt + 3 * RB_BLK + o1) = make_uint4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]);
    *reinterpret_cast<uint4*>(gt + 3 * RB_BLK + o2) = make_uint4(lo[3][0], lo[3][1], lo[3][2], lo[3][3]);
  };
  auto store_zero = [&]() {                             // no live row in this warp at this step: zero rows in both products
    const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      *reinterpret_cast<uint4*>(gt + b * RB_BLK + o1) = zz;
      *reinterpret_cast<uint4*>(gt + b * RB_BLK + o2) = zz;
    }
  };
  // ---- half A: hidden units 0..31
  uint32_t v[32];
  if (any_live) {
    uint32_t hi[4][4], lo[4][4];
    bwd_cell_load(trow, ua, v);
    tmem_ld_wait();
    bwd_cell8(v, ua, himg, o1, s_bhn, dy4[0], dy4[1], live, g.part, hi, lo);
    store_tile(hi, lo);
    // half B's accumulators are read BEFORE half A is handed over: the carry product of half A lands in the n_h columns
    bwd_cell_load(trow, ub, v);
    tmem_ld_wait();
  } else {
    store_zero();
  }
  fence_async_smem();
  tc_fence_before();                   // every accumulator column this thread reads has been read
  mbar_arrive(&bar->ta_ready);
  if (tr) BTRACE(0, n, 4);
  // ---- half B: hidden units 32..63, computed while the MMAs of half A run; written once they have read the tile
  if (any_live) {
    uint32_t hi[4][4], lo[4][4];
    bwd_cell8(v, ub, himg, o2, s_bhn, dy4[2], dy4[3], live, g.part + 8, hi, lo);
    mbar_wait(&bar->ta_done, n & 1);
    if (tr) BTRACE(0, n, 5);
    store_tile(hi, lo);
  } else {
    mbar_wait(&bar->ta_done, n & 1);
    store_zero();
  }
  fence_async_smem();
  mbar_arrive(&bar->tb_ready);
  if (tr) BTRACE(0, n, 6);
}

__global__ void __launch_bounds__(RB_THREADS, 1) gru_bwd_tc_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ BwdBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bhn[H];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wih = base;                                  // [hi|lo][192][128 B]
  unsigned char* whh = base + RB_W_BYTES;
  unsigned char* ximg = base + 2 * RB_W_BYTES;                // xq image (hi|lo), then hq image (hi|lo)
  unsigned char* himg = ximg + RB_IMG;
  unsigned char* gt = ximg + 2 * RB_IMG;                      // gate-gradient tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) mbar_init(&bars.img_q[q], 1);
    mbar_init(&bars.p1_full, 1);
    mbar_init(&bars.ta_ready, RB_GATE_THREADS);
    mbar_init(&bars.ta_done, 1);
    mbar_init(&bars.tb_ready, RB_GATE_THREADS);
    mbar_init(&bars.dh_full, 1);
    mbar_init(&bars.step_done, 1);
    mbar_init(&bars.dh_read, RB_GATE_THREADS);
    mbar_fence_init();
  }
  if (warp == RB_GATE_WARPS) tmem_alloc(&tmem_slot, 512);
  {
    // resident weights of this direction, scaled exactly as in the forward kernel (the recomputed accumulators are then
    // bit-identical to the forward's): W_ih with column E = b_ih (+ b_hh for r, z), W_hh, b_hn
    const float* w_ih = a.w[dir * 4 + 0], *w_hh = a.w[dir * 4 + 1], *b_ih = a.w[dir * 4 + 2], *b_hh = a.w[dir * 4 + 3];
    for (int idx = tid; idx < G3 * 16; idx += RB_THREADS) {
      const int n = idx >> 4, k = (idx & 15) * 4;
      const float sc = n < 2 * H ? RB_K_RZ : RB_K_N;
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
    if (tid < H) s_bhn[tid] = RB_K_N * b_hh[2 * H + tid];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) BTRACE(1, 63, 0);

  // this CTA walks slot queue 2c, then 2c+1, one tile at a time
  if (warp < RB_GATE_WARPS) {
    const int row = (warp & 3) * 32 + lane, ug = warp >> 2;
    BwdRow g;
    g.len = 0; g.rowo = -1;
    int n_total = 0;
    for (int qi = 0; qi < 2; ++qi) {
      Cur c;
      cur_init(a, c, 2 * blockIdx.x + qi);
      for (; c.active; ++n_total) {
        bwd_gate_step(a, c, g, n_total, dir, row, ug, himg, gt, s_bhn, &bars, tmem);
        cur_next(a, c);
      }
    }
  } else {
    // ------------------------------------------------------------------ driver warp: the whole warp runs this code converged, so
    // that descriptors and addresses stay in uniform registers; the elected lane alone issues the TMA copies and the MMAs (tc.cuh)
    const uint32_t el = elect_one_sync();
    constexpr uint32_t id192 = idesc_bf16(128, 192), id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
    constexpr uint32_t id_carry = idesc_bf16(128, 64) | (1u << 16);                  // A = gate tile (K-major), B = W_hh image MN-major
    constexpr uint32_t id_wg = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);       // A = [xq | hq] images, B = gate tile: both token-major
    const uint64_t wih_h = smem_desc_sw128(smem_u32(wih)), wih_l = smem_desc_sw128(smem_u32(wih + G3 * 128));
    const uint64_t whh_h = smem_desc_sw128(smem_u32(whh)), whh_l = smem_desc_sw128(smem_u32(whh + G3 * 128));
    const uint64_t whn_h = smem_desc_sw128(smem_u32(whh + 128 * 128)), whn_l = smem_desc_sw128(smem_u32(whh + G3 * 128 + 128 * 128));
    const uint64_t x_h = smem_desc_sw128(smem_u32(ximg)), x_l = smem_desc_sw128(smem_u32(ximg + RT_R * 128));
    const uint64_t h_h = smem_desc_sw128(smem_u32(himg)), h_l = smem_desc_sw128(smem_u32(himg + RT_R * 128));
    // descriptor of byte offset `o` (a multiple of 16) from a base descriptor: the start-address field counts 16-byte units
    const uint64_t cb_h = smem_desc_mn_sw128(smem_u32(whh)), cb_l = smem_desc_mn_sw128(smem_u32(whh + G3 * 128));       // carry B: W_hh MN-major
    const uint64_t ca_h = smem_desc_sw128(smem_u32(gt)), ca_l = smem_desc_sw128(smem_u32(gt + 2 * RB_BLK));             // carry A: gate tile K-major
    const uint64_t wa_h = desc_mn(smem_u32(ximg), RB_IMG), wa_l = desc_mn(smem_u32(ximg + RT_R * 128), RB_IMG);         // wgrad A: [xq | hq] MN-major
    const uint64_t wb_h = desc_mn(smem_u32(gt), RB_BLK), wb_l = desc_mn(smem_u32(gt + 2 * RB_BLK), RB_BLK);             // wgrad B: gate tile MN-major
    const uint32_t d_gates = tmem, d_dh = tmem + 192, d_w = tmem + 256;
    auto slab_of = [&](const Cur& cc, int& t, int& tp) -> size_t {
      const BwdSeg& sg = a.seg[cc.si];
      t = dir ? cc.s : (cc.Lj - 1 - cc.s);
      tp = dir ? t + 1 : t - 1;
      return (size_t)cc.slab0;
    };
    auto images_of = [&](const Cur& cc, const unsigned char*& xsrc, const unsigned char*& hsrc) {
      const BwdSeg& sg = a.seg[cc.si];
      int t, tp;
      const size_t slab0 = slab_of(cc, t, tp);
      xsrc = sg.xq + (slab0 + t) * RB_IMG;
      hsrc = (tp >= 0 && tp < cc.Lj) ? sg.hq + ((slab0 + tp) * 2 + dir) * RB_IMG : a.zero_img;
    };
    // the step's two operand images, one TMA bulk copy each, landing on their own barriers (the x-part of the recomputation
    // starts as soon as the token image is there)
    auto produce = [&](const unsigned char* xsrc, const unsigned char* hsrc) {
      mbar_arrive_expect_tx_e(el, &bars.img_q[0], RB_IMG);
      bulk_copy_g2s_e(el, ximg, xsrc, RB_IMG, &bars.img_q[0]);
      mbar_arrive_expect_tx_e(el, &bars.img_q[1], RB_IMG);
      bulk_copy_g2s_e(el, himg, hsrc, RB_IMG, &bars.img_q[1]);
    };
    Cur c;
    int qi = 0;
    cur_init(a, c, 2 * blockIdx.x);
    if (!c.active) { qi = 1; cur_init(a, c, 2 * blockIdx.x + 1); }
    const unsigned char* nx_x = nullptr, *nx_h = nullptr;
    if (c.active) { images_of(c, nx_x, nx_h); produce(nx_x, nx_h); }
    int n = 0;
    for (; c.active; ++n) {
      if (el) BTRACE(1, n, 0);
      // ---- recomputation: the forward's MMAs (gru_rec_tc.cu)
      mbar_wait(&bars.img_q[0], n & 1);         // token image
      tc_fence_after();
      if (el) BTRACE(1, n, 1);
      for (int kk = 0; kk < a.kx; ++kk) {       // x_t · W_ih^T  -> r, z, n_x  (overwrites)
        const uint64_t o = (uint64_t)(kk * 2);
        umma_bf16_e(el, d_gates, x_h + o, wih_h + o, id192, kk != 0);
        umma_bf16_e(el, d_gates, x_h + o, wih_l + o, id192, 1);
        umma_bf16_e(el, d_gates, x_l + o, wih_h + o, id192, 1);
      }
      mbar_wait(&bars.img_q[1], n & 1);         // hidden image
      if (n > 0) mbar_wait(&bars.dh_read, (n - 1) & 1);      // the previous carry product has been folded in: its columns (= n_h) are free
      tc_fence_after();
      if (el) BTRACE(1, n, 2);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {          // h_{t-1} · W_hh^T -> += r, z ; n_h (own columns)
        const uint64_t o = (uint64_t)(kk * 2);
        umma_bf16_e(el, d_gates, h_h + o, whh_h + o, id128, 1);
        umma_bf16_e(el, d_gates, h_h + o, whh_l + o, id128, 1);
        umma_bf16_e(el, d_gates, h_l + o, whh_h + o, id128, 1);
        umma_bf16_e(el, d_gates + 192, h_h + o, whn_h + o, id64, kk != 0);
        umma_bf16_e(el, d_gates + 192, h_h + o, whn_l + o, id64, 1);
        umma_bf16_e(el, d_gates + 192, h_l + o, whn_h + o, id64, 1);
      }
      umma_commit_e(el, &bars.p1_full);
      if (el) BTRACE(1, n, 3);
      // while the gate threads work: the next step's operand images - addresses resolved now (queue / plan look-ups are global
      // loads), pulled into L2, copied as soon as this step's MMAs have retired
      Cur nx = c;
      cur_next(a, nx);
      if (!nx.active && qi == 0) { qi = 1; cur_init(a, nx, 2 * blockIdx.x + 1); }
      if (nx.active) {
        images_of(nx, nx_x, nx_h);
        bulk_prefetch_l2_e(el, nx_x, RB_IMG);
        if (nx_h != a.zero_img) bulk_prefetch_l2_e(el, nx_h, RB_IMG);
      }
      // ---- the two halves of the hidden units: tile = [dr | dz][dn*r | dn] of units 0..31, then of units 32..63
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && el) BTRACE(1, n, 5);
        mbar_wait(half == 0 ? &bars.ta_ready : &bars.tb_ready, n & 1);
        tc_fence_after();
        if (el) BTRACE(1, n, half == 0 ? 4 : 6);
        // carry: dh[128 x 64] (+)= tile[128 x 96: dr, dz, dn*r of this half] · W_hh[those 96 rows][64]; six k-steps of 16 gates
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
          // k-step ks: gate type ks / 2 (dr, dz in block 0; dn*r in block 1), units 16 (ks % 2) .. +15 of the half
          const uint64_t ao = (uint64_t)((((ks >> 2) * RB_BLK) + (ks & 3) * 32) >> 4);
          const uint64_t bo = (uint64_t)((((ks >> 1) * 4 + half * 2 + (ks & 1)) * 2048) >> 4);      // W_hh rows 64 type + 32 half + 16 (ks % 2)
          umma_bf16_e(el, d_dh, ca_h + ao, cb_h + bo, id_carry, (half | ks) != 0);
          umma_bf16_e(el, d_dh, ca_h + ao, cb_l + bo, id_carry, 1);
          umma_bf16_e(el, d_dh, ca_l + ao, cb_h + bo, id_carry, 1);
        }
        if (half == 1) umma_commit_e(el, &bars.dh_full);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {        // dW^T[128 features][half*128 + 128 gates] += [xq | hq]^T · tile, K = 128 sequences
          const uint64_t o = (uint64_t)((ks * 2048) >> 4);
          const uint32_t accf = (n | ks) != 0;
          umma_bf16_e(el, d_w + half * 128, wa_h + o, wb_h + o, id_wg, accf);
          umma_bf16_e(el, d_w + half * 128, wa_h + o, wb_l + o, id_wg, 1);
          umma_bf16_e(el, d_w + half * 128, wa_l + o, wb_h + o, id_wg, 1);
        }
        umma_commit_e(el, half == 0 ? &bars.ta_done : &bars.step_done);
      }
      if (nx.active) {
        mbar_wait(&bars.step_done, n & 1);      // every MMA that reads this step's images and tile has retired
        if (el) BTRACE(1, n, 7);
        produce(nx_x, nx_h);
      }
      c = nx;
    }
    if (n > 0) mbar_wait(&bars.step_done, (n - 1) & 1);      // the accumulator is complete before the flush below
  }
  __syncthreads();
  tc_fence_after();
  if (tid == 0) BTRACE(1, 63, 1);
  // ---- flush dW^T: TMEM lane = feature (x features 0..63, hidden features 64..127); column c: half = c / 128 (hidden units 0..31 |
  //      32..63), gate type = (c % 128) / 32 (dr, dz, dn*r, dn), unit = 32 half + c % 32; scaled by 1 / (the weights' pre-scaling)
  const bool has_work = a.q_off[2 * blockIdx.x + 2] > a.q_off[2 * blockIdx.x];       // otherwise the accumulator was never written
  if (warp < RB_GATE_WARPS && has_work) {
    // warp w reads the lanes of quarter w % 4 (all a warp may touch) and the 64 columns of column group w / 4
    const int f = (warp & 3) * 32 + lane;
    float* dw_ih = a.dw[dir * 4 + 0];
    float* dw_hh = a.dw[dir * 4 + 1];
    float* db_ih = a.dw[dir * 4 + 2];
    float* db_hh = a.dw[dir * 4 + 3];
    const int E = a.E;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 64; c0 < (warp >> 2) * 64 + 64; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + c0, v);
      const int type = (c0 & 127) >> 5;          // one gate type per 32 columns
      const float sc = type < 2 ? RB_K_RZ : RB_K_N;
      const int ubase = (c0 >> 7) * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int u = ubase + i;
        const float val = v[i] * sc;
        if (f < KP) {
          // token features: W_ih columns, the 1.0 column feeds the biases.  dr -> row u, dz -> 64 + u, dn -> 128 + u
          if (type != 2) {
            const int grow = (type == 3 ? 2 : type) * H + u;
            if (f < E) atomicAdd(&dw_ih[grow * E + f], val);
            else if (f == E) { atomicAdd(&db_ih[grow], val); if (type < 2) atomicAdd(&db_hh[grow], val); }
          } else if (f == E) {
            atomicAdd(&db_hh[2 * H + u], val);   // dn*r -> b_hn
          }
        } else if (type != 3) {
          atomicAdd(&dw_hh[(type * H + u) * H + (f - KP)], val);       // hidden features: W_hh columns; dr, dz, dn*r -> rows u, 64+u, 128+u
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) BTRACE(1, 63, 2);
  if (warp == RB_GATE_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_gru_bwd_tc(const umpr_gru_bwd_seg* segs, int n_seg, const float* const* w, float* const* dw, int E,
                               const void* zero_img, const int32_t* sched, int n_queues, void* stream) {
  if (n_seg < 1 || n_seg > RT_MAX_SEG) return fail_arg("gru_bwd_tc: n_seg=%d not in [1,%d]", n_seg, RT_MAX_SEG);
  if (n_queues < 2 || (n_queues & 1)) return fail_arg("gru_bwd_tc: n_queues=%d must be even and >= 2", n_queues);
  if (E < 1 || E >= KP) return fail_arg("gru_bwd_tc: E=%d must be in [1,%d)", E, KP);
  if (!zero_img) return fail_arg("gru_bwd_tc: zero_img (32 KB of zeros) is required");
  BwdArgs a{};
  int base = 0;
  for (int i = 0; i < n_seg; ++i) {
    const umpr_gru_bwd_seg& s = segs[i];
    if (s.n_tiles < 1 || s.L < 1 || !s.d_out || !s.xq || !s.hq || !s.plan) return fail_arg("gru_bwd_tc: segment %d is incomplete", i);
    if (reinterpret_cast<uintptr_t>(s.d_out) & 15) return fail_arg("gru_bwd_tc: d_out must be 16-byte aligned");
    a.seg[i] = BwdSeg{s.d_out, s.d_hn, reinterpret_cast<const unsigned char*>(s.xq), reinterpret_cast<const unsigned char*>(s.hq), s.plan,
                      s.n_tiles, s.n_slabs, s.N, s.L, base};
    base += s.n_tiles;
  }
  a.n_seg = n_seg;
  a.q_off = sched;
  a.q_tile = sched + n_queues + 1;
  for (int i = 0; i < 8; ++i) { a.w[i] = w[i]; a.dw[i] = dw[i]; }
  a.zero_img = reinterpret_cast<const unsigned char*>(zero_img);
  a.E = E;
  a.kx = (E + 1 + 15) / 16;
  a.trace = nullptr;
  const char* trace_path = getenv("UMPR_TRACE_BWD");
  static long long* tr = nullptr;
  if (trace_path) {
    if (!tr) cudaMalloc(&tr, 2 * 64 * 8 * sizeof(long long));
    cudaMemsetAsync(tr, 0, 2 * 64 * 8 * sizeof(long long), (cudaStream_t)stream);
    a.trace = tr;
  }
  cudaError_t e = cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM);
  if (e != cudaSuccess) { set_error("gru_bwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  gru_bwd_tc_kernel<<<dim3(n_queues / 2, 2), RB_THREADS, RB_SMEM, (cudaStream_t)stream>>>(a);
  if (trace_path) {
    cudaStreamSynchronize((cudaStream_t)stream);
    static long long host[2 * 64 * 8];
    cudaMemcpy(host, tr, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      const long long t0 = host[(1 * 64 + 0) * 8 + 0];
      for (int role = 0; role < 2; ++role)
        for (int n = 0; n < 40; ++n) {
          fprintf(f, "%s step %2d:", role ? "driver" : "gate  ", n);
          for (int ev = 0; ev < 8; ++ev) fprintf(f, " %8lld", host[(role * 64 + n) * 8 + ev] ? host[(role * 64 + n) * 8 + ev] - t0 : -1);
          fprintf(f, "\n");
        }
      fprintf(f, "phases (cycles since the driver's first event): main loop starts %lld, ends %lld, flush done %lld\n",
              host[(1 * 64 + 63) * 8 + 0] - t0, host[(1 * 64 + 63) * 8 + 1] - t0, host[(1 * 64 + 63) * 8 + 2] - t0);
      fclose(f);
    }
  }
  return check_launch("gru_bwd_tc");
}
