// ImprovedRnn backward on the tensor cores: reverse-time recurrence AND the GRU weight gradients in one persistent kernel
// (backward of reference src/model.py:19-21; main.py:36).  The gate gradients never leave the SM.
//
// Per CTA (one direction, one 128-sequence tile at a time, same tile queues as the forward), per time step in reverse order:
//     dh_t   = dy_t + dh_{t+1} (.) z_{t+1} + [dr, dz, dn*r]_{t+1} · W_hh                         (carry)
//     dn_pre = dh (1-z)(1-n^2)   dz_pre = dh (h_{t-1} - n) z (1-z)   dr_pre = dn_pre (W_hn h + b_hn) r (1-r)
//     dW^T[feature][gate] += [x_t | h_{t-1}]^T · [dr, dz, dn, dn*r]                                 (weight gradients)
//   * carry product (K=192, N=64, 36 MMAs): A operand = the gate threads' own rows written into TENSOR MEMORY (tcgen05.st),
//     B operand = the resident W_hh image read MN-major (the bytes the forward reads K-major);
//   * weight-gradient product (M=128 features, N=256 gates in two passes of 128, K=128 sequences, 2 x 24 MMAs): A operand =
//     the forward's token image xq and hidden image hq (bf16 hi|lo operand images, token-major = MN-major, loaded by ONE TMA
//     bulk copy each), B operand = the gate gradients written by the gate threads as a token-major bf16 hi|lo tile;
//     accumulated in TMEM over the CTA's whole queue and flushed once with atomics into the eight nn.GRU gradients;
//   * TMEM (512 columns): [0,64) dh | [64,160) carry A hi | [160,256) carry A lo | [256,512) dW^T accumulator;
//   * warps 0-7: gate threads (one sequence row x 32 hidden units each): saved gates, d_out and h_{t-1} (hq image, hi + lo) are
//     read straight from global memory (prefetched into L2 one step ahead); warp 8: driver thread (TMA copies of the two
//     operand images, MMAs);
//   * `out` is not touched; saved gates svT are column-major inside a (slab, direction) tile (lanes = rows read 128 contiguous bytes).
#include "common.cuh"
#include "tc.cuh"
#include "gru_tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int RB_GATE_WARPS = 8;
constexpr int RB_THREADS = (RB_GATE_WARPS + 1) * 32;      // 288
constexpr int RB_W_BYTES = 2 * G3 * 128;                  // W_hh hi | lo, [192][64 bf16]
constexpr int RB_IMG = 2 * RT_R * 128;                    // one operand image: hi | lo, [128][64 bf16] = 32 KB
constexpr int RB_GT = 4 * RT_R * 128;                     // gate-gradient tile: [hi|lo][2 blocks of 64 gates][128 sequences][128 B] = 64 KB
constexpr int RB_SMEM = RB_W_BYTES + 2 * RB_IMG + RB_GT + 1024;

struct BwdSeg {
  const float* d_out; const float* d_hn; const float* sv; const unsigned char* xq; const unsigned char* hq; const int* plan;
  int n_tiles, n_slabs, N, L, tile_base;
};
struct BwdArgs {
  BwdSeg seg[RT_MAX_SEG];
  int n_seg;
  const int* q_off; const int* q_tile;
  const float* w[8];
  float* dw[8];
  const unsigned char* zero_img;      // 32 KB of zeros: h_{t-1} of a sequence's first step
  int E;
};

struct BwdRow {
  float part[32];      // dh_{t+1} (.) z_{t+1} (+ d_hn at the row's last step): the element-wise half of the carry
  int len, rowo;
};

struct BwdBars {
  uint64_t stage_full, stage_free, p1_ready, p2_ready, acc_full, w1_done, w2_done;
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

// Everything a gate thread needs for 8 hidden units of its row that does NOT depend on the carry: saved gates (column-major
// tile: lanes = rows are contiguous), d_out (the row's own 32 bytes) and h_{t-1} (one 16-byte chunk of the hq image, hi and lo).
// Loaded straight from global memory (the driver pulls the step's tile into L2 one step ahead).
struct ChunkIn {
  float r[8], z[8], n[8], hh[8];
  float4 y0, y1;
  uint4 hhi, hlo;
};
__device__ __forceinline__ void load_chunk(ChunkIn& in, bool live, const float* svcol, const float* dyrow, const unsigned char* hrow, int cc,
                                           uint32_t off) {
  if (live) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* s1 = svcol + (size_t)(cc * 8 + i) * RT_R;
      in.r[i] = s1[0]; in.z[i] = s1[(size_t)H * RT_R]; in.n[i] = s1[(size_t)2 * H * RT_R]; in.hh[i] = s1[(size_t)3 * H * RT_R];
    }
    in.y0 = *reinterpret_cast<const float4*>(dyrow + cc * 8);
    in.y1 = *reinterpret_cast<const float4*>(dyrow + cc * 8 + 4);
    in.hhi = make_uint4(0u, 0u, 0u, 0u); in.hlo = in.hhi;
    if (hrow) {
      in.hhi = *reinterpret_cast<const uint4*>(hrow + off);
      in.hlo = *reinterpret_cast<const uint4*>(hrow + RT_R * 128 + off);
    }
  }
}

__device__ __forceinline__ void bwd_gate_step(const BwdArgs& a, const Cur& c, BwdRow& g, int n, int dir, int row, int hf,
                                              unsigned char* base, BwdBars* bar, uint32_t tmem) {
  const BwdSeg& sg = a.seg[c.si];
  const int u0 = hf * 32;
  const int Rp = sg.n_tiles * RT_R;
  unsigned char* gt = base + RB_W_BYTES + 2 * RB_IMG;       // gate-gradient tile
  if (c.s == 0) {
    const int k = c.tile * RT_R + row;
    g.rowo = sg.plan[Rp + k];
    g.len = sg.plan[2 * Rp + k];
#pragma unroll
    for (int i = 0; i < 32; ++i) g.part[i] = 0.f;
    if (sg.d_hn && g.rowo >= 0) {
      const float* hp = sg.d_hn + ((size_t)dir * sg.N + sg.plan[k]) * H + u0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(hp + i * 4);
        g.part[4 * i] = v.x; g.part[4 * i + 1] = v.y; g.part[4 * i + 2] = v.z; g.part[4 * i + 3] = v.w;
      }
    }
  }
  const int t = dir ? c.s : (c.Lj - 1 - c.s);          // reverse of the forward kernel's order
  const int tp = dir ? t + 1 : t - 1;
  const bool live = t < g.len;
  const size_t slab0 = (size_t)sg.plan[3 * Rp + c.tile];
  const float* svcol = sg.sv + (((slab0 + t) * 2 + dir) * SV + u0) * RT_R + row;
  const float* dyrow = sg.d_out + ((size_t)(g.rowo < 0 ? 0 : g.rowo) * sg.L + t) * D + dir * H + u0;
  const unsigned char* hrow = (tp >= 0 && tp < c.Lj) ? sg.hq + ((slab0 + tp) * 2 + dir) * RB_IMG : nullptr;     // h_{t-1} image (zeros at the first step)
  const uint32_t trow = tmem + ((uint32_t)((row >> 5) * 32) << 16);
  const uint32_t off0 = (uint32_t)(row * 128);

  ChunkIn ci;                                   // chunk 0 is requested before the barrier waits: its latency hides behind them
  load_chunk(ci, live, svcol, dyrow, hrow, 0, off0 + (((hf * 4) ^ (row & 7)) << 4));
  if (n > 0) {
    mbar_wait(&bar->acc_full, (n - 1) & 1);     // previous carry product retired: dh readable, its TMEM A operand reusable
    mbar_wait(&bar->w2_done, (n - 1) & 1);      // previous weight-gradient MMAs retired: the gate-gradient tile is reusable
    tc_fence_after();
  }
  uint32_t dn_hi[16], dn_lo[16], dnr_hi[16], dnr_lo[16];          // pass-2 gate gradients, kept packed until pass 1 retires
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int chunk = hf * 4 + cc;
    const uint32_t off = off0 + ((chunk ^ (row & 7)) << 4);
    if (cc > 0) load_chunk(ci, live, svcol, dyrow, hrow, cc, off);
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0u;
    if (c.s > 0) { tmem_ld8_issue(trow + u0 + cc * 8, acc); tmem_ld_wait(); }     // warp-uniform: outside the per-row branch
    float dr[8], dz[8], dn8[8], dnr[8];
    if (live) {
      const float dy[8] = {ci.y0.x, ci.y0.y, ci.y0.z, ci.y0.w, ci.y1.x, ci.y1.y, ci.y1.z, ci.y1.w};
      float hp[8];
      unpack8(ci.hhi, ci.hlo, hp);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dh = g.part[cc * 8 + i] + __uint_as_float(acc[i]) + dy[i];
        const float dnv = dh * (1.f - ci.z[i]);
        const float dn_pre = dnv * (1.f - ci.n[i] * ci.n[i]);
        dz[i] = dh * (hp[i] - ci.n[i]) * ci.z[i] * (1.f - ci.z[i]);
        dr[i] = dn_pre * ci.hh[i] * ci.r[i] * (1.f - ci.r[i]);
        dnr[i] = dn_pre * ci.r[i];
        dn8[i] = dn_pre;
        g.part[cc * 8 + i] = dh * ci.z[i];
      }
    } else {
      // beyond this row's length: no gradient (zero rows in both products), the carry just passes through
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g.part[cc * 8 + i] += __uint_as_float(acc[i]);
        dr[i] = 0.f; dz[i] = 0.f; dn8[i] = 0.f; dnr[i] = 0.f;
      }
    }
    // carry A operand in tensor memory: k = gate block * 64 + unit, two bf16 per 32-bit column, hi at +64, lo at +160;
    // weight-gradient B operand in shared memory: token-major tile, block 0 = dr (pass 2: dn), block 1 = dz (pass 2: dn*r)
    const uint32_t acol = trow + 64 + (u0 + cc * 8) / 2;
    const uint32_t goff = off;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dr[2 * i], dr[2 * i + 1], hi[i], lo[i]);
    tmem_st4(acol, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(acol + 96, lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(gt + goff) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(gt + 2 * RT_R * 128 + goff) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dz[2 * i], dz[2 * i + 1], hi[i], lo[i]);
    tmem_st4(acol + 32, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(acol + 96 + 32, lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(gt + RT_R * 128 + goff) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(gt + 3 * RT_R * 128 + goff) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dnr[2 * i], dnr[2 * i + 1], dnr_hi[cc * 4 + i], dnr_lo[cc * 4 + i]);
    tmem_st4(acol + 64, dnr_hi[cc * 4], dnr_hi[cc * 4 + 1], dnr_hi[cc * 4 + 2], dnr_hi[cc * 4 + 3]);
    tmem_st4(acol + 96 + 64, dnr_lo[cc * 4], dnr_lo[cc * 4 + 1], dnr_lo[cc * 4 + 2], dnr_lo[cc * 4 + 3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dn8[2 * i], dn8[2 * i + 1], dn_hi[cc * 4 + i], dn_lo[cc * 4 + i]);
  }
  tmem_st_wait();
  fence_async_smem();
  tc_fence_before();
  mbar_arrive(&bar->p1_ready);         // carry A operand + pass-1 tile (dr | dz) complete
  // pass 2: the same tile buffer, once the pass-1 MMAs have read it
  mbar_wait(&bar->w1_done, n & 1);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int chunk = hf * 4 + cc;
    const uint32_t off = (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(gt + off) = make_uint4(dn_hi[cc * 4], dn_hi[cc * 4 + 1], dn_hi[cc * 4 + 2], dn_hi[cc * 4 + 3]);
    *reinterpret_cast<uint4*>(gt + 2 * RT_R * 128 + off) = make_uint4(dn_lo[cc * 4], dn_lo[cc * 4 + 1], dn_lo[cc * 4 + 2], dn_lo[cc * 4 + 3]);
    *reinterpret_cast<uint4*>(gt + RT_R * 128 + off) = make_uint4(dnr_hi[cc * 4], dnr_hi[cc * 4 + 1], dnr_hi[cc * 4 + 2], dnr_hi[cc * 4 + 3]);
    *reinterpret_cast<uint4*>(gt + 3 * RT_R * 128 + off) = make_uint4(dnr_lo[cc * 4], dnr_lo[cc * 4 + 1], dnr_lo[cc * 4 + 2], dnr_lo[cc * 4 + 3]);
  }
  fence_async_smem();
  mbar_arrive(&bar->p2_ready);
}

__global__ void __launch_bounds__(RB_THREADS, 1) gru_bwd_tc_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ BwdBars bars;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* whh = base;                                  // [hi|lo][192][128 B]
  unsigned char* ximg = base + RB_W_BYTES;                    // xq image (hi|lo), then hq image (hi|lo): operands of the weight-gradient MMAs
  unsigned char* gt = ximg + 2 * RB_IMG;                      // gate-gradient tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;

  if (tid == 0) {
    mbar_init(&bars.stage_full, 1);
    mbar_init(&bars.stage_free, 1);
    mbar_init(&bars.p1_ready, RB_GATE_WARPS * 32);
    mbar_init(&bars.p2_ready, RB_GATE_WARPS * 32);
    mbar_init(&bars.acc_full, 1);
    mbar_init(&bars.w1_done, 1);
    mbar_init(&bars.w2_done, 1);
    mbar_fence_init();
  }
  if (warp == RB_GATE_WARPS) tmem_alloc(&tmem_slot, 512);
  {
    const float* w_hh = a.w[dir * 4 + 1];
    for (int idx = tid; idx < G3 * 16; idx += RB_THREADS) {
      const int n = idx >> 4, k = (idx & 15) * 4;
      store_split4(whh, whh + G3 * 128, n, k, *reinterpret_cast<const float4*>(w_hh + n * H + k));
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // this CTA walks slot queue 2c, then 2c+1 (one tile at a time: the kernel is bound by the saved-gate traffic, not by overlap)
  int n_total = 0;
  if (warp < RB_GATE_WARPS) {
    const int row = (warp & 3) * 32 + lane, hf = warp >> 2;
    BwdRow g;
    g.len = 0; g.rowo = -1;
    for (int qi = 0; qi < 2; ++qi) {
      Cur c;
      cur_init(a, c, 2 * blockIdx.x + qi);
      for (; c.active; ++n_total) {
        bwd_gate_step(a, c, g, n_total, dir, row, hf, base, &bars, tmem);
        cur_next(a, c);
      }
    }
    if (n_total > 0) {
      mbar_wait(&bars.acc_full, (n_total - 1) & 1);
      mbar_wait(&bars.w2_done, (n_total - 1) & 1);
      tc_fence_after();
    }
  } else {
    // ------------------------------------------------------------------ driver warp
    constexpr uint32_t id_carry = idesc_bf16(128, 64) | (1u << 16);                  // A in TMEM (K-major), B = W_hh image MN-major
    constexpr uint32_t id_wg = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);       // A = [xq | hq] images, B = gate-gradient tile: both token-major
    const uint32_t b_hi = smem_u32(whh), b_lo = smem_u32(whh + G3 * 128);
    const uint32_t d_dh = tmem, a_hi = tmem + 64, a_lo = tmem + 160, d_w = tmem + 256;
    // weight-gradient A operand: M = 128 features = [64 of xq | 64 of hq]: the second 64-feature block lies one image (32 KB) further
    const uint32_t xa_hi = smem_u32(ximg), xa_lo = smem_u32(ximg + RT_R * 128);
    const uint32_t g_hi = smem_u32(gt), g_lo = smem_u32(gt + 2 * RT_R * 128);
    auto produce = [&](const Cur& cc) {      // the step's two operand images, one TMA bulk copy each (lane 0)
      if (lane != 0) return;
      const BwdSeg& sg = a.seg[cc.si];
      const int t = dir ? cc.s : (cc.Lj - 1 - cc.s);
      const int tp = dir ? t + 1 : t - 1;
      const size_t slab0 = (size_t)sg.plan[3 * sg.n_tiles * RT_R + cc.tile];
      mbar_arrive_expect_tx(&bars.stage_full, 2 * RB_IMG);
      bulk_copy_g2s(ximg, sg.xq + (slab0 + t) * RB_IMG, RB_IMG, &bars.stage_full);
      const unsigned char* hsrc = (tp >= 0 && tp < cc.Lj) ? sg.hq + ((slab0 + tp) * 2 + dir) * RB_IMG : a.zero_img;
      bulk_copy_g2s(ximg + RB_IMG, hsrc, RB_IMG, &bars.stage_full);
    };
    Cur c;
    int qi = 0;
    cur_init(a, c, 2 * blockIdx.x);
    if (!c.active) { qi = 1; cur_init(a, c, 2 * blockIdx.x + 1); }
    if (c.active) produce(c);
    for (int n = 0; c.active; ++n) {
      Cur nx = c;
      cur_next(a, nx);
      if (!nx.active && qi == 0) { qi = 1; cur_init(a, nx, 2 * blockIdx.x + 1); }
      if (nx.active && lane == 0) {
        // the next step's saved gates (128 KB) and operand images are pulled into L2 while this step computes
        const BwdSeg& sg = a.seg[nx.si];
        const int t = dir ? nx.s : (nx.Lj - 1 - nx.s);
        const int tp = dir ? t + 1 : t - 1;
        const size_t slab0 = (size_t)sg.plan[3 * sg.n_tiles * RT_R + nx.tile];
        bulk_prefetch_l2(sg.sv + ((slab0 + t) * 2 + dir) * SV * RT_R, SV * RT_R * 4);
        bulk_prefetch_l2(sg.xq + (slab0 + t) * RB_IMG, RB_IMG);
        if (tp >= 0 && tp < nx.Lj) bulk_prefetch_l2(sg.hq + ((slab0 + tp) * 2 + dir) * RB_IMG, RB_IMG);
      }
      if (lane == 0) {
        mbar_wait(&bars.p1_ready, n & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 12; ++ks) {       // carry: dh[128 x 64] = A[128 x 192] (TMEM) · W_hh[192 x 64]
          const uint64_t bh = smem_desc_mn_sw128(b_hi + ks * 2048), bl = smem_desc_mn_sw128(b_lo + ks * 2048);
          umma_bf16_ts(d_dh, a_hi + ks * 8, bh, id_carry, ks != 0);
          umma_bf16_ts(d_dh, a_hi + ks * 8, bl, id_carry, 1);
          umma_bf16_ts(d_dh, a_lo + ks * 8, bh, id_carry, 1);
        }
        umma_commit(&bars.acc_full);
        mbar_wait(&bars.stage_full, n & 1);      // the step's operand images have landed
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {  // dW^T[128 features][pass*128 + 128 gates] += [xq | hq]^T · tile, K = 128 sequences
          if (pass == 1) { mbar_wait(&bars.p2_ready, n & 1); tc_fence_after(); }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ah = desc_mn(xa_hi + ks * 2048, RB_IMG), al = desc_mn(xa_lo + ks * 2048, RB_IMG);
            const uint64_t gh = desc_mn(g_hi + ks * 2048, RT_R * 128), gl = desc_mn(g_lo + ks * 2048, RT_R * 128);
            const uint32_t accf = (n | ks) != 0;
            umma_bf16(d_w + pass * 128, ah, gh, id_wg, accf);
            umma_bf16(d_w + pass * 128, ah, gl, id_wg, 1);
            umma_bf16(d_w + pass * 128, al, gh, id_wg, 1);
          }
          umma_commit(pass == 0 ? &bars.w1_done : &bars.w2_done);
        }
        umma_commit(&bars.stage_free);           // (same completion as w2_done: operand images and staging are free again)
      }
      __syncwarp();
      if (nx.active) {
        mbar_wait(&bars.stage_free, n & 1);     // all MMAs that read this step's images have retired (the gate threads read theirs before p1_ready)
        produce(nx);
      }
      c = nx;
    }
  }
  __syncthreads();
  tc_fence_after();
  // ---- flush dW^T: TMEM lane = feature (x features 0..63, hidden features 64..127), column = gate (dr, dz, dn, dn*r blocks of 64)
  const bool has_work = a.q_off[2 * blockIdx.x + 2] > a.q_off[2 * blockIdx.x];       // otherwise the accumulator was never written
  if (warp < 4 && has_work) {
    const int f = warp * 32 + lane;
    float* dw_ih = a.dw[dir * 4 + 0];
    float* dw_hh = a.dw[dir * 4 + 1];
    float* db_ih = a.dw[dir * 4 + 2];
    float* db_hh = a.dw[dir * 4 + 3];
    const int E = a.E;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int gc = c0 + i;                 // 0..255: dr, dz, dn, dn*r
        if (f < KP) {
          if (gc < G3) {
            if (f < E) atomicAdd(&dw_ih[gc * E + f], v[i]);
            else if (f == E) { atomicAdd(&db_ih[gc], v[i]); if (gc < 2 * H) atomicAdd(&db_hh[gc], v[i]); }
          } else if (f == E) {
            atomicAdd(&db_hh[gc - H], v[i]);   // dn*r column u -> b_hn at index 128 + u
          }
        } else {
          const int j = f - KP;
          if (gc < 2 * H) atomicAdd(&dw_hh[gc * H + j], v[i]);
          else if (gc >= G3) atomicAdd(&dw_hh[(gc - H) * H + j], v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
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
    if (s.n_tiles < 1 || s.L < 1 || !s.d_out || !s.sv || !s.xq || !s.hq || !s.plan) return fail_arg("gru_bwd_tc: segment %d is incomplete", i);
    if (reinterpret_cast<uintptr_t>(s.d_out) & 15) return fail_arg("gru_bwd_tc: d_out must be 16-byte aligned");
    a.seg[i] = BwdSeg{s.d_out, s.d_hn, s.sv, reinterpret_cast<const unsigned char*>(s.xq), reinterpret_cast<const unsigned char*>(s.hq), s.plan,
                      s.n_tiles, s.n_slabs, s.N, s.L, base};
    base += s.n_tiles;
  }
  a.n_seg = n_seg;
  a.q_off = sched;
  a.q_tile = sched + n_queues + 1;
  for (int i = 0; i < 8; ++i) { a.w[i] = w[i]; a.dw[i] = dw[i]; }
  a.zero_img = reinterpret_cast<const unsigned char*>(zero_img);
  a.E = E;
  cudaError_t e = cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM);
  if (e != cudaSuccess) { set_error("gru_bwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  gru_bwd_tc_kernel<<<dim3(n_queues / 2, 2), RB_THREADS, RB_SMEM, (cudaStream_t)stream>>>(a);
  return check_launch("gru_bwd_tc");
}
