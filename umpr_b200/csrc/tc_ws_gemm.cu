// Weight-stationary persistent tcgen05 GEMM: C[M][BN] = A[M][K] · B[BN][K]^T, K = KB*64, B resident in shared memory.
// One CTA per SM loops over 128-row tiles of A:
//   warps 0-7  loaders   : global fp32 rows -> bf16 hi/lo -> SWIZZLE_128B shared memory (NSTAGE-deep ring); the whole tile's loads
//                          are in flight together, the valid-row tables of the next tiles are fetched one / two tiles ahead
//   warp  8    MMA issuer: one elected thread, 3xBF16 (hi*hi + hi*lo + lo*hi) into one of two TMEM accumulators
//   warps 9-12 epilogue  : tcgen05.ld -> bias/accumulate/activation -> shared-memory transpose -> coalesced 128-B row stores
// so the loads of tile i+1, the MMAs of tile i and the stores of tile i-1 overlap.  The same skeleton (resident weights,
// TMEM double buffer, mbarrier hand-offs) is what the GRU recurrence kernel uses.
//   MODE 0: plain (gi·M, dgiM·M^T of model.py:50 and its backward)      MODE 1: GRU input projection (model.py:19)
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int WS_THREADS = 416;   // 13 warps: 0-7 loaders, 8 MMA issuer, 9-12 epilogue
constexpr int WS_STG_LD = 36;     // floats per staged row (32 + 4: conflict-free 128-bit accesses)

struct WsArgs {
  const float* A; long lda;
  const float* B; long ldb; int b_kn;
  float* C; long ldc;
  int M, N, K;
  int accumulate, act;
  const float* bias;
  const float* w_ih[2]; const float* b_ih[2]; const float* b_hh[2];
  int E, R;
  // valid-row mode (MODE 0): tiles of whole consecutive sentences, <= 128 valid rows each (plan.py:snet_table); row r of sentence n is
  // global row n*L + r.  tso == NULL: plain 128-row tiles of all M rows
  const int* tso; const int* cst; int L, n_row_tiles;
  int dbg;
};

constexpr int WS_NMETA = 8;      // deeper than stages + accumulators: the slot of tile it-8 is free for every NSTAGE
struct WsMeta { int rows; int rowmap[128]; };

template <int BN, int KB, int NSTAGE> struct WsSmem {
  static constexpr int B_BYTES = KB * 2 * BN * 128;            // [kb][hi|lo][BN][128 B]
  static constexpr int A_STAGE = KB * 2 * 128 * 128;           // [kb][hi|lo][128][128 B]
  static constexpr int STG_BYTES = 4 * 32 * WS_STG_LD * 4;
  static constexpr int TOTAL = B_BYTES + NSTAGE * A_STAGE + STG_BYTES + 1024;
};

template <int BN, int KB, int NSTAGE, int MODE>
__global__ void __launch_bounds__(WS_THREADS, 1) tc_ws_gemm_kernel(const WsArgs a) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t a_full[NSTAGE], a_empty[NSTAGE], acc_full[2], acc_empty[2], m_full[WS_NMETA];
  __shared__ uint32_t tmem_slot;
  __shared__ WsMeta meta[MODE == 0 ? WS_NMETA : 1];
  using SM = WsSmem<BN, KB, NSTAGE>;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* bsm = base;
  unsigned char* asm_ = base + SM::B_BYTES;
  float* stg = reinterpret_cast<float*>(base + SM::B_BYTES + NSTAGE * SM::A_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  constexpr uint32_t TCOLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  const bool by_rows = MODE == 0 && a.tso != nullptr;
  const int n_tiles = by_rows ? a.n_row_tiles : (a.M + 127) / 128;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&a_full[s], 256); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    for (int s = 0; s < WS_NMETA; ++s) mbar_init(&m_full[s], 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, TCOLS);
  // ---- resident B (weights), loaded once by all threads
  for (int idx = tid; idx < KB * BN * 16; idx += WS_THREADS) {
    const int kb = idx / (BN * 16), rem = idx - kb * BN * 16;
    int n, k;
    if (MODE == 0 && a.b_kn) { n = rem % BN; k = (rem / BN) * 4; } else { n = rem >> 4; k = (rem & 15) * 4; }
    const int kg = kb * 64 + k;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < a.N) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = kg + q;
        if (MODE == 1) {
          t[q] = kk < a.E ? a.w_ih[dir][(long)n * a.E + kk] : (kk == a.E ? a.b_ih[dir][n] + (n < 2 * H ? a.b_hh[dir][n] : 0.f) : 0.f);
        } else if (kk < a.K) {
          t[q] = a.b_kn ? a.B[(long)kk * a.ldb + n] : a.B[(long)n * a.ldb + kk];
        }
      }
    }
    unsigned char* bt = bsm + kb * 2 * BN * 128;
    store_split4(bt, bt + BN * 128, n, k, make_float4(t[0], t[1], t[2], t[3]));
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ loaders
    // valid-row tables, software-pipelined: tile boundaries (tso) two tiles ahead, sentence offsets (cst) one tile ahead, so that no
    // table load sits in front of a tile's row loads
    int s0c = 0, s1c = 0, s0n = 0, s1n = 0, cbc = 0, cec = 0, c0c = 0, cendc = 0;
    auto ld_tso = [&](int tile, int& s0, int& s1) { if (tile < n_tiles) { s0 = a.tso[tile]; s1 = a.tso[tile + 1]; } else { s0 = s1 = 0; } };
    auto ld_cst = [&](int s0, int s1, int& cb, int& ce, int& c0, int& cend) {
      c0 = a.cst[s0]; cend = a.cst[s1];
      if (tid < s1 - s0) { cb = a.cst[s0 + tid]; ce = a.cst[s0 + tid + 1]; }
    };
    if (by_rows) {
      ld_tso(blockIdx.x, s0c, s1c);
      ld_tso(blockIdx.x + gridDim.x, s0n, s1n);
      ld_cst(s0c, s1c, cbc, cec, c0c, cendc);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % NSTAGE;
      int s0nn = 0, s1nn = 0, cbn = 0, cen = 0, c0n = 0, cendn = 0;
      if (by_rows) {
        ld_cst(s0n, s1n, cbn, cen, c0n, cendn);                    // next tile (s0n == s1n == 0 past the end: harmless reads of cst[0])
        ld_tso(tile + 2 * gridDim.x, s0nn, s1nn);
      }
      if (it >= NSTAGE) mbar_wait(&a_empty[s], ((it / NSTAGE) - 1) & 1);
      unsigned char* st = asm_ + s * SM::A_STAGE;
      const int m0 = tile * 128;
      const WsMeta& mt = meta[MODE == 0 ? it % WS_NMETA : 0];
      int rows = 128;
      if (by_rows) {
        // slot it % 8 is free: the ring wait above implies the epilogue of tile it-5 has finished (it arrives acc_empty last)
        WsMeta& mw = meta[MODE == 0 ? it % WS_NMETA : 0];
        if (tid < s1c - s0c) {
          const int b = cbc - c0c, e = cec - c0c, g0 = (s0c + tid) * a.L - b;
          for (int r = b; r < e; ++r) mw.rowmap[r] = g0 + r;
        }
        rows = cendc - c0c;
        if (tid == 0) mw.rows = rows;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      float4 va[KB][8];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 256 + tid, r = idx >> 4, k = kb * 64 + (idx & 15) * 4;
          const int m = by_rows ? (r < rows ? mt.rowmap[r] : a.M) : m0 + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < a.M && !(a.dbg & 2)) {
            const float* p = a.A + (long)m * a.lda + k;
            if (k + 3 < a.K) v = *reinterpret_cast<const float4*>(p);
            else {
              if (k < a.K) v.x = p[0];
              if (k + 1 < a.K) v.y = p[1];
              if (k + 2 < a.K) v.z = p[2];
            }
          }
          va[kb][i] = v;
        }
      }
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        if (a.dbg & 4) break;
        unsigned char* a_hi = st + kb * 2 * 128 * 128, *a_lo = a_hi + 128 * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 256 + tid;
          store_split4(a_hi, a_lo, idx >> 4, (idx & 15) * 4, va[kb][i]);
        }
      }
      fence_async_smem();
      mbar_arrive(&a_full[s]);
      if (by_rows) mbar_arrive(&m_full[it % WS_NMETA]);
      s0c = s0n; s1c = s1n; s0n = s0nn; s1n = s1nn;
      cbc = cbn; cec = cen; c0c = c0n; cendc = cendn;
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, the elected lane issues)
    {
      const uint32_t el = elect_one_sync();
      constexpr uint32_t idesc = idesc_bf16(128, BN);
      const uint32_t b0 = smem_u32(bsm);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % NSTAGE, acc = it & 1;
        if (it >= 2) mbar_wait(&acc_empty[acc], ((it >> 1) - 1) & 1);
        mbar_wait(&a_full[s], (it / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(asm_ + s * SM::A_STAGE);
        const uint32_t d = tmem + acc * BN;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t ah = smem_desc_sw128(a0 + kb * 2 * 128 * 128), al = smem_desc_sw128(a0 + kb * 2 * 128 * 128 + 128 * 128);
          const uint64_t bh = smem_desc_sw128(b0 + kb * 2 * BN * 128), bl = smem_desc_sw128(b0 + kb * 2 * BN * 128 + BN * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (a.dbg & 8) break;
            const uint64_t o = (uint64_t)(kk * 2);
            umma_bf16_e(el, d, ah + o, bh + o, idesc, (kb | kk) != 0);
            umma_bf16_e(el, d, ah + o, bl + o, idesc, 1);
            umma_bf16_e(el, d, al + o, bh + o, idesc, 1);
          }
        }
        umma_commit_e(el, &a_empty[s]);
        umma_commit_e(el, &acc_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 9..12 -> TMEM lane quarter warp%4)
    const int q = warp & 3;
    float* sw = stg + q * 32 * WS_STG_LD;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int m0 = tile * 128 + q * 32;
      const int c4 = (lane & 7) * 4;
      const WsMeta& mt = meta[MODE == 0 ? it % WS_NMETA : 0];
      if (by_rows) mbar_wait(&m_full[it % WS_NMETA], (it / WS_NMETA) & 1);
      // global row of tile row rr (a.M = none)
      auto grow = [&](int rr) { return (by_rows && !(a.dbg & 64)) ? (q * 32 + rr < mt.rows ? mt.rowmap[q * 32 + rr] : a.M) : m0 + rr; };
      // the 8 output rows this thread stores (row j*4 + lane/8 of the warp's 32, 4 columns at c4): resolved once per tile
      float* crow[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int m = grow(j * 4 + (lane >> 3));
        crow[j] = nullptr;
        if (m < a.M) {
          if (MODE == 1) {
            const int slab = m / a.R, rr = m - slab * a.R;
            crow[j] = a.C + (((long)slab * 2 + dir) * a.R + rr) * G3 + c4;
          } else {
            crow[j] = a.C + (long)m * a.ldc + c4;
          }
        }
      }
      const bool acc_c = MODE == 0 && a.accumulate;
      const bool plain = MODE == 1 || (!a.accumulate && !a.bias && a.act == 0);
      const bool acc_only = MODE == 0 && a.accumulate && !a.bias && a.act == 0;
      // accumulate mode: the C rows of a 32-column chunk are requested one chunk ahead (the first one before the accumulator
      // is even ready), so their latency hides behind the TMEM load / staging of the previous chunk
      float4 cnext[8];
      auto fetch_c = [&](int c0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          cnext[j] = (crow[j] && c0 + c4 < a.N) ? *reinterpret_cast<const float4*>(crow[j] + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (acc_c) fetch_c(0);
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (a.dbg & 1) break;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * BN + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(&sw[lane * WS_STG_LD + j * 4]) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
        __syncwarp();
        const bool col_ok = c0 + c4 < a.N;
        if (plain) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 o = *reinterpret_cast<const float4*>(&sw[(j * 4 + (lane >> 3)) * WS_STG_LD + c4]);
            if (crow[j] && col_ok) *reinterpret_cast<float4*>(crow[j] + c0) = o;
          }
        } else if (acc_only) {                                       // C += A·B^T, nothing else (the co-attention backward's dgi)
          float4 ccur[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ccur[j] = cnext[j];
          if (c0 + 32 < BN) fetch_c(c0 + 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = *reinterpret_cast<const float4*>(&sw[(j * 4 + (lane >> 3)) * WS_STG_LD + c4]);
            o.x += ccur[j].x; o.y += ccur[j].y; o.z += ccur[j].z; o.w += ccur[j].w;
            if (crow[j] && col_ok) *reinterpret_cast<float4*>(crow[j] + c0) = o;
          }
        } else {
          float4 ccur[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ccur[j] = cnext[j];
          if (acc_c && c0 + 32 < BN) fetch_c(c0 + 32);
          const int n = c0 + c4;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (crow[j] && col_ok) {
              float4 o = *reinterpret_cast<const float4*>(&sw[(j * 4 + (lane >> 3)) * WS_STG_LD + c4]);
              if (a.accumulate) { o.x += ccur[j].x; o.y += ccur[j].y; o.z += ccur[j].z; o.w += ccur[j].w; }
              if (a.bias) { o.x += a.bias[n]; o.y += a.bias[n + 1]; o.z += a.bias[n + 2]; o.w += a.bias[n + 3]; }
              if (a.act == 1) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
              else if (a.act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(crow[j] + c0) = o;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, TCOLS);
}

template <int BN, int KB, int NSTAGE, int MODE> static int launch_ws(const WsArgs& a, int n_ctas, int gy, cudaStream_t st) {
  constexpr int smem = WsSmem<BN, KB, NSTAGE>::TOTAL;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(tc_ws_gemm_kernel<BN, KB, NSTAGE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("tc_ws_gemm smem: %s", cudaGetErrorString(e)); return (int)e; }
  const int n_tiles = (MODE == 0 && a.tso) ? a.n_row_tiles : (a.M + 127) / 128;
  const int grid = n_tiles < n_ctas ? n_tiles : n_ctas;
  tc_ws_gemm_kernel<BN, KB, NSTAGE, MODE><<<dim3(grid, gy), WS_THREADS, smem, st>>>(a);
  return check_launch("tc_ws_gemm");
}

}  // namespace umpr

using namespace umpr;

// C[M][N<=128] = act(acc*C + A[M][K<=128] · B^T + bias): persistent weight-stationary tensor-core GEMM (N multiple of 4).
// table (optional) = [tile_sent_off (n_row_tiles+1) | cstart (M/L+1)] (plan.py:snet_table) for an A whose rows are sentences of L
// positions with exactly-zero rows at and beyond each sentence's length: only the valid rows are read, multiplied and written
// (the other rows of C are left untouched).
extern "C" int umpr_tc_gemm_ws(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K,
                               int accumulate, const float* bias, int act, int b_kn, const int32_t* table, int n_row_tiles, int L,
                               int n_ctas, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (N > 128 || K > 128 || (N & 3)) return fail_arg("tc_gemm_ws: N=%d K=%d (needs N<=128, N%%4==0, K<=128)", N, K);
  if ((lda & 3) || (ldc & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(C) & 15))
    return fail_arg("tc_gemm_ws: A and C must be 16-byte aligned with leading dimensions that are multiples of 4");
  if (act < 0 || act > 2) return fail_arg("tc_gemm_ws: act=%d", act);
  WsArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.b_kn = b_kn; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.accumulate = accumulate; a.act = act; a.bias = bias; a.dbg = dbg_flags();
  if (table) {
    if (n_row_tiles < 1 || L < 1 || L > 128 || M % L) return fail_arg("tc_gemm_ws: row table inconsistent (n_row_tiles=%d, L=%d, M=%d)", n_row_tiles, L, M);
    a.tso = table; a.cst = table + n_row_tiles + 1; a.L = L; a.n_row_tiles = n_row_tiles;
  }
  if (n_ctas < 1) n_ctas = 148;
  if (K <= 64) return launch_ws<128, 1, 3, 0>(a, n_ctas, 1, (cudaStream_t)stream);
  return launch_ws<128, 2, 2, 0>(a, n_ctas, 1, (cudaStream_t)stream);
}

// tensor-core variant of umpr_gru_inproj (same arguments, same output layout)
extern "C" int umpr_gru_inproj_tc(const float* xp, const float* const* w, int n_slabs, int R, int E, float* G, int n_ctas, void* stream) {
  if (E < 1 || E >= KP) return fail_arg("gru_inproj: E=%d", E);
  if (R != 32 && R != 64 && R != 128) return fail_arg("gru_inproj: R=%d", R);
  if (n_slabs == 0) return 0;
  WsArgs a{};
  a.A = xp; a.lda = KP; a.C = G; a.M = n_slabs * R; a.N = G3; a.K = KP; a.E = E; a.R = R;
  a.w_ih[0] = w[0]; a.b_ih[0] = w[2]; a.b_hh[0] = w[3];
  a.w_ih[1] = w[4]; a.b_ih[1] = w[6]; a.b_hh[1] = w[7];
  if (n_ctas < 2) n_ctas = 148;
  return launch_ws<192, 1, 3, 1>(a, n_ctas / 2, 2, (cudaStream_t)stream);
}
