// Shared pieces of the fused tensor-core GRU kernels (gru_rec_tc.cu forward, gru_bwd_tc.cu backward): the per-slot tile-queue
// cursor every warp role walks, and small device helpers.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace umpr {
using namespace tc;

constexpr int RT_R = 128;          // sequences per tile = MMA M
constexpr int RT_MAX_SEG = 3;      // ImprovedRnn calls (review sides) per launch

// one slot's position in its tile queue; every warp role walks the same deterministic sequence
struct Cur {
  int q, qend, s, Lj, tile, si;
  int slab0;          // first slab of the tile (tile_off[tile]): looked up once per tile - a global load per step would sit on the
  bool active;        // critical path of the drivers
};
template <class Args> __device__ __forceinline__ void cur_tile(const Args& a, Cur& c) {
  const int g = a.q_tile[c.q];
  int si = 0;
  if (a.n_seg > 1 && g >= a.seg[1].tile_base) si = 1;
  if (a.n_seg > 2 && g >= a.seg[2].tile_base) si = 2;
  c.si = si;
  c.tile = g - a.seg[si].tile_base;
  c.Lj = a.seg[si].plan[2 * a.seg[si].n_tiles * RT_R + c.tile * RT_R];    // len_of[tile*R]: the tile's longest job
  c.slab0 = a.seg[si].plan[3 * a.seg[si].n_tiles * RT_R + c.tile];         // tile_off[tile]
  c.s = 0;
}
template <class Args> __device__ __forceinline__ void cur_init(const Args& a, Cur& c, int qi) {
  c.q = a.q_off[qi]; c.qend = a.q_off[qi + 1];
  c.active = c.q < c.qend;
  c.s = 0; c.Lj = 0; c.tile = 0; c.si = 0; c.slab0 = 0;
  if (c.active) cur_tile(a, c);
}
template <class Args> __device__ __forceinline__ void cur_next(const Args& a, Cur& c) {
  if (++c.s == c.Lj) {
    if (++c.q < c.qend) cur_tile(a, c); else c.active = false;
  }
}

__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// wait for the outstanding TMEM loads; the registers are listed as in/out operands so no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait8(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 of sm_100): the gate math is bound by the FP32 pipe's issue rate, one instruction per two
// hidden units halves it.  MUFU (ex2, rcp) and min stay scalar.
typedef float2 f2;
__device__ __forceinline__ f2 f2_set(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 f2_bits(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 f2_ex2(f2 a) { return make_float2(ex2f(a.x), ex2f(a.y)); }
__device__ __forceinline__ f2 f2_rcp(f2 a) { return make_float2(rcpf(a.x), rcpf(a.y)); }
__device__ __forceinline__ f2 f2_min(f2 a, float m) { return make_float2(fminf(a.x, m), fminf(a.y, m)); }
__device__ __forceinline__ void st_zero8(float* p) {
  *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(p + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}


// In-warp transpose of an 8x8 matrix of float4 inside each group of 8 lanes: before, lane j of a group holds a[k] = float4 #k of
// ITS row; after, lane j holds a[i] = float4 #j of the group's row i.  A warp-wide 128-bit store of a[i] then writes 4 rows x
// 128 contiguous bytes (full lines) instead of 32 rows x 16 bytes.  3 butterfly stages, 16 shuffles each.
__device__ __forceinline__ void transpose8x8_f4(float4 (&a)[8], int lane) {
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k & s) continue;
      float4 snd = up ? a[k] : a[k | s];
      float4 rcv;
      rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, s);
      rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, s);
      rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, s);
      rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, s);
      if (up) a[k] = rcv; else a[k | s] = rcv;
    }
  }
}

}  // namespace umpr
