// ImprovedRnn backward recurrence on the tensor cores (backward of reference src/model.py:19-21; main.py:36).
//
// Reverse-time walk over the same 128-sequence tiles / slot queues as the fused forward (gru_rec_tc.cu).  Per step and slot
//     dh_t   = dy_t + dh_{t+1} (.) z_{t+1} + [dr, dz, dn*r]_{t+1} · W_hh          (carry)
//     dn_pre = dh (1-z)(1-n^2)   dz_pre = dh (h_{t-1} - n) z (1-z)   dr_pre = dn_pre (W_hn h + b_hn) r (1-r)
// the [128 x 192] · [192 x 64] carry product is ONE tcgen05 chain per step whose A operand (the three gate-gradient blocks,
// bf16 hi and lo) is written by the gate threads straight into TENSOR MEMORY (tcgen05.st, lane = sequence) - no shared-memory
// round trip - and whose B operand is the resident W_hh image read MN-major (the same bytes the forward reads K-major).
//   TMEM per slot (256 columns): [0,64) dh accumulator | [64,160) A hi (192 k as bf16 pairs) | [160,256) A lo
//   * warps 0-7 / 8-15: gate threads of slot 0 / 1 (one sequence row x 32 hidden units each);
//   * warps 16 / 17: slot drivers: 32 lanes issue the per-row TMA bulk copies of dy (d_out row) and h_{t-1} (out row) into a padded
//     staging buffer, lane 0 issues the step's 36 MMAs;
//   * saved gates svT and the produced dGT = [dr, dz, dn, dn*r] are column-major inside a (slab, direction) tile, so lanes
//     (= rows) read and write contiguous 128-byte lines; umpr_gru_wgrad_tc2 consumes dGT as a K-major operand.
#include "common.cuh"
#include "tc.cuh"
#include "gru_tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int RB_GATE_WARPS = 16;
constexpr int RB_THREADS = (RB_GATE_WARPS + 2) * 32;
constexpr int RB_W_BYTES = 2 * G3 * 128;            // W_hh hi | lo, [192][64 bf16]
constexpr int RB_ROW = 272;                         // staged row: 64 floats + 16 B pad (conflict-free 128-bit row reads)
constexpr int RB_STAGE = RT_R * RB_ROW;             // one staged [128 x 64] fp32 tile
constexpr int RB_SMEM = RB_W_BYTES + 4 * RB_STAGE + 1024;

struct BwdSeg {
  const float* d_out; const float* d_hn; const float* out; const float* sv; float* dG; const int* plan;
  int n_tiles, n_slabs, N, L, tile_base;
};
struct BwdArgs {
  BwdSeg seg[RT_MAX_SEG];
  int n_seg;
  const int* q_off; const int* q_tile;
  const float* w[8];
};

struct BwdRow {
  float part[32];      // dh_{t+1} (.) z_{t+1} (+ d_hn at the row's last step): the element-wise half of the carry
  int len, rowo;
};

__device__ __forceinline__ void bwd_gate_step(const int X, const BwdArgs& a, const Cur& c, BwdRow& g, int n, int dir, int row, int hf,
                                              const unsigned char* stage, uint64_t* stage_full, uint64_t* stage_empty, uint64_t* a_ready,
                                              uint64_t* acc_full, uint32_t tmem) {
  const BwdSeg& sg = a.seg[c.si];
  const int u0 = hf * 32;
  const int Rp = sg.n_tiles * RT_R;
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
  const bool live = t < g.len;
  const int tp = dir ? t + 1 : t - 1;
  const bool has_prev = live && tp >= 0 && tp < g.len;
  const size_t tile_col = ((size_t)(sg.plan[3 * Rp + c.tile] + t) * 2 + dir) * SV + u0;     // column (gate block 0, unit u0) of this slab
  const float* svcol = sg.sv + tile_col * RT_R + row;
  float* dgcol = sg.dG + tile_col * RT_R + row;
  const float* dy_s = reinterpret_cast<const float*>(stage + (X * 2 + 0) * RB_STAGE + row * RB_ROW) + u0;
  const float* hp_s = reinterpret_cast<const float*>(stage + (X * 2 + 1) * RB_STAGE + row * RB_ROW) + u0;
  const uint32_t trow = tmem + ((uint32_t)((row >> 5) * 32) << 16) + X * 256;

  if (n > 0) {                      // previous step's carry product done: its accumulator is readable, its A operand reusable
    mbar_wait(&acc_full[X], (n - 1) & 1);
    tc_fence_after();
  }
  mbar_wait(&stage_full[X], n & 1);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0u;
    if (c.s > 0) { tmem_ld8_issue(trow + u0 + cc * 8, acc); tmem_ld_wait(); }     // warp-uniform: outside the per-row branch
    float dr[8], dz[8], dnr[8];
    if (live) {
      const float4 y0 = *reinterpret_cast<const float4*>(dy_s + cc * 8), y1 = *reinterpret_cast<const float4*>(dy_s + cc * 8 + 4);
      float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
      if (has_prev) { p0 = *reinterpret_cast<const float4*>(hp_s + cc * 8); p1 = *reinterpret_cast<const float4*>(hp_s + cc * 8 + 4); }
      const float dy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      const float hp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float rr[8], zz[8], nn[8], hh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* s1 = svcol + (size_t)(cc * 8 + i) * RT_R;
        rr[i] = s1[0]; zz[i] = s1[(size_t)H * RT_R]; nn[i] = s1[(size_t)2 * H * RT_R]; hh[i] = s1[(size_t)3 * H * RT_R];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dh = g.part[cc * 8 + i] + __uint_as_float(acc[i]) + dy[i];
        const float dn = dh * (1.f - zz[i]);
        const float dn_pre = dn * (1.f - nn[i] * nn[i]);
        const float dz_pre = dh * (hp[i] - nn[i]) * zz[i] * (1.f - zz[i]);
        const float dr_pre = dn_pre * hh[i] * rr[i] * (1.f - rr[i]);
        const float dnr_pre = dn_pre * rr[i];
        g.part[cc * 8 + i] = dh * zz[i];
        float* d1 = dgcol + (size_t)(cc * 8 + i) * RT_R;
        d1[0] = dr_pre; d1[(size_t)H * RT_R] = dz_pre; d1[(size_t)2 * H * RT_R] = dn_pre; d1[(size_t)3 * H * RT_R] = dnr_pre;
        dr[i] = dr_pre; dz[i] = dz_pre; dnr[i] = dnr_pre;
      }
    } else {
      // beyond this row's length: no gradient (zeros for the weight-gradient sum), the carry just passes through
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g.part[cc * 8 + i] += __uint_as_float(acc[i]);
        float* d1 = dgcol + (size_t)(cc * 8 + i) * RT_R;
        d1[0] = 0.f; d1[(size_t)H * RT_R] = 0.f; d1[(size_t)2 * H * RT_R] = 0.f; d1[(size_t)3 * H * RT_R] = 0.f;
        dr[i] = 0.f; dz[i] = 0.f; dnr[i] = 0.f;
      }
    }
    // A operand of the carry product: k = gate block * 64 + unit, two bf16 per 32-bit TMEM column, hi at +64, lo at +160
    const uint32_t acol = trow + 64 + (u0 + cc * 8) / 2;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dr[2 * i], dr[2 * i + 1], hi[i], lo[i]);
    tmem_st4(acol, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(acol + 96, lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dz[2 * i], dz[2 * i + 1], hi[i], lo[i]);
    tmem_st4(acol + 32, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(acol + 96 + 32, lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(dnr[2 * i], dnr[2 * i + 1], hi[i], lo[i]);
    tmem_st4(acol + 64, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(acol + 96 + 64, lo[0], lo[1], lo[2], lo[3]);
  }
  mbar_arrive(&stage_empty[X]);       // staged dy / h_{t-1} rows consumed
  tmem_st_wait();
  tc_fence_before();
  mbar_arrive(&a_ready[X]);
}

__global__ void __launch_bounds__(RB_THREADS, 1) gru_bwd_tc_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t stage_full[2], stage_empty[2], a_ready[2], acc_full[2];
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* whh = base;                          // [hi|lo][192][128 B]
  unsigned char* stage = base + RB_W_BYTES;           // [slot][dy|hp][128][272 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&stage_full[s], 32);
      mbar_init(&stage_empty[s], RB_GATE_WARPS * 16);
      mbar_init(&a_ready[s], RB_GATE_WARPS * 16);
      mbar_init(&acc_full[s], 1);
    }
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

  if (warp < RB_GATE_WARPS) {
    const int X = warp >> 3;
    const int row = (warp & 3) * 32 + lane, hf = (warp >> 2) & 1;
    BwdRow g;
    g.len = 0; g.rowo = -1;
    Cur c;
    cur_init(a, c, 2 * blockIdx.x + X);
    int n = 0;
    for (; c.active; ++n) {
      bwd_gate_step(X, a, c, g, n, dir, row, hf, stage, stage_full, stage_empty, a_ready, acc_full, tmem);
      cur_next(a, c);
    }
    if (n > 0) mbar_wait(&acc_full[X], (n - 1) & 1);      // the last carry product must retire before tensor memory is freed
  } else {
    // ------------------------------------------------------------------ slot driver warp
    const int X = warp - RB_GATE_WARPS;
    constexpr uint32_t idesc = idesc_bf16(128, 64) | (1u << 16);            // B (W_hh image) is read MN-major: B[n = unit][k = gate row]
    const uint32_t b_hi = smem_u32(whh), b_lo = smem_u32(whh + G3 * 128);
    const uint32_t d = tmem + X * 256, a_hi = d + 64, a_lo = d + 160;
    unsigned char* st_dy = stage + (X * 2 + 0) * RB_STAGE;
    unsigned char* st_hp = stage + (X * 2 + 1) * RB_STAGE;
    Cur c;
    cur_init(a, c, 2 * blockIdx.x + X);
    int cur_tile_id = -1, cur_seg = -1;
    int rowo[4], len[4];
    auto produce = [&](const Cur& cc) {
      // per-row TMA bulk copies of dy (d_out row) and h_{t-1} (out row) for this (tile, step); lane handles rows lane + 32 q
      const BwdSeg& sg = a.seg[cc.si];
      if (cc.tile != cur_tile_id || cc.si != cur_seg) {
        const int Rp = sg.n_tiles * RT_R;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = cc.tile * RT_R + lane + 32 * q;
          rowo[q] = sg.plan[Rp + k];
          len[q] = sg.plan[2 * Rp + k];
        }
        cur_tile_id = cc.tile; cur_seg = cc.si;
      }
      const int t = dir ? cc.s : (cc.Lj - 1 - cc.s);
      const int tp = dir ? t + 1 : t - 1;
      uint32_t bytes = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (rowo[q] >= 0 && t < len[q]) bytes += 256 + ((tp >= 0 && tp < len[q]) ? 256 : 0);
      }
      mbar_arrive_expect_tx(&stage_full[X], bytes);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (rowo[q] >= 0 && t < len[q]) {
          const int r = lane + 32 * q;
          bulk_copy_g2s(st_dy + r * RB_ROW, sg.d_out + ((size_t)rowo[q] * sg.L + t) * D + dir * H, 256, &stage_full[X]);
          if (tp >= 0 && tp < len[q])
            bulk_copy_g2s(st_hp + r * RB_ROW, sg.out + ((size_t)rowo[q] * sg.L + tp) * D + dir * H, 256, &stage_full[X]);
        }
      }
    };
    if (c.active) produce(c);
    for (int n = 0; c.active; ++n) {
      Cur nx = c;
      cur_next(a, nx);
      if (nx.active) {
        mbar_wait(&stage_empty[X], n & 1);      // the gate threads have read step n's staged rows
        produce(nx);
      }
      if (lane == 0) {
        mbar_wait(&a_ready[X], n & 1);          // A operand written (tcgen05.st) and the dh accumulator drained
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 12; ++ks) {       // K = 192 gate rows, 16 per MMA: dh[128 x 64] = A[128 x 192] · W_hh[192 x 64]
          const uint64_t bh = smem_desc_mn_sw128(b_hi + ks * 2048), bl = smem_desc_mn_sw128(b_lo + ks * 2048);
          umma_bf16_ts(d, a_hi + ks * 8, bh, idesc, ks != 0);
          umma_bf16_ts(d, a_hi + ks * 8, bl, idesc, 1);
          umma_bf16_ts(d, a_lo + ks * 8, bh, idesc, 1);
        }
        umma_commit(&acc_full[X]);
      }
      __syncwarp();
      c = nx;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RB_GATE_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_gru_bwd_tc(const umpr_gru_bwd_seg* segs, int n_seg, const float* const* w, const int32_t* sched, int n_queues,
                               void* stream) {
  if (n_seg < 1 || n_seg > RT_MAX_SEG) return fail_arg("gru_bwd_tc: n_seg=%d not in [1,%d]", n_seg, RT_MAX_SEG);
  if (n_queues < 2 || (n_queues & 1)) return fail_arg("gru_bwd_tc: n_queues=%d must be even and >= 2", n_queues);
  BwdArgs a{};
  int base = 0;
  for (int i = 0; i < n_seg; ++i) {
    const umpr_gru_bwd_seg& s = segs[i];
    if (s.n_tiles < 1 || s.L < 1 || !s.d_out || !s.out || !s.sv || !s.dG || !s.plan) return fail_arg("gru_bwd_tc: segment %d is incomplete", i);
    if ((reinterpret_cast<uintptr_t>(s.d_out) | reinterpret_cast<uintptr_t>(s.out)) & 15) return fail_arg("gru_bwd_tc: d_out / out must be 16-byte aligned");
    a.seg[i] = BwdSeg{s.d_out, s.d_hn, s.out, s.sv, s.dG, s.plan, s.n_tiles, s.n_slabs, s.N, s.L, base};
    base += s.n_tiles;
  }
  a.n_seg = n_seg;
  a.q_off = sched;
  a.q_tile = sched + n_queues + 1;
  for (int i = 0; i < 8; ++i) a.w[i] = w[i];
  cudaError_t e = cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM);
  if (e != cudaSuccess) { set_error("gru_bwd_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  gru_bwd_tc_kernel<<<dim3(n_queues / 2, 2), RB_THREADS, RB_SMEM, (cudaStream_t)stream>>>(a);
  return check_launch("gru_bwd_tc");
}
