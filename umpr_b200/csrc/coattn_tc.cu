// R-Net co-attention affinity on the tensor cores (reference src/model.py:50-53), flash-style.
//
// Per CTA: one sample b and one 128-row tile of giM (resident in shared memory as bf16 hi/lo, SWIZZLE_128B); the gu
// tiles stream through a 2-stage ring.  For every (i-tile, j-tile) pair TWO products are issued from the same operand
// tiles: S = giM_i · gu_j^T (rows = i) and S^T = gu_j · giM_i^T (rows = j), so both the row maxima (over j) and the
// column maxima (over i) are plain per-thread scans of TMEM rows - no cross-thread reduction.  TMEM: 2 x (S, S^T) = 512 columns.
//
// Arg-max routing decides where gradients flow, and 3xBF16 products carry ~1e-5 relative error, so near-ties are NOT
// resolved here: each row / column keeps up to 4 candidates within a conservative error bound of its running maximum;
// coattn_resolve_kernel re-scores the surviving candidates with exact fp32 dot products and picks the winner.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
using namespace tc;

constexpr int CA_THREADS = 288;
constexpr float CA_EPS = 6.2e-5f;        // 2^-14: bound on the relative error of a 3xBF16 dot product (vs |a||b|)
constexpr float CA_GNORM = 11.32f;       // sqrt(128): |gu_j| <= sqrt(D) because GRU outputs lie in (-1, 1)

struct Cand4 {
  float v[4]; int id[4];
  float thr;            // v[0] - tau: anything above it is a candidate
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = -INFINITY; id[q] = -1; }
    thr = -INFINITY;
  }
  // scan 8 consecutive accumulator columns; the common case is one compare per value
  __device__ __forceinline__ void scan8(const float* x, int idx0, int nvalid, float tau) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nvalid && x[c] > thr) { push(x[c], idx0 + c); thr = fmaxf(thr, v[0] - tau); }
    }
  }
  // keep the 4 largest distinct values, sorted descending (rare path: kept out of line so the scan loops stay small)
  __device__ __noinline__ void push(float x, int idx) {
    if (x == v[0] || x == v[1] || x == v[2] || x == v[3]) return;     // exact duplicates (zero padding) carry no information
    if (x > v[3]) {
      v[3] = x; id[3] = idx;
#pragma unroll
      for (int q = 3; q > 0; --q) {
        if (v[q] > v[q - 1]) {
          const float tv = v[q]; v[q] = v[q - 1]; v[q - 1] = tv;
          const int ti = id[q]; id[q] = id[q - 1]; id[q - 1] = ti;
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------------
// version 2: one CTA per sample, two passes over the sample's tile pairs.
//   pass 0: maxima only (row maxima per thread of the S tile, column maxima per thread of the S^T tile);
//   pass 1: the same products again (MMAs are cheap), now every value within the error bound of its FINAL row / column maximum
//           is kept as a candidate (up to 4 distinct values) for coattn_resolve_kernel's exact fp32 re-scoring.
// Operands are pre-split bf16 hi|lo images (coattn_images_kernel), loaded by TMA bulk copies: no conversion in the hot loop.
//   warps 0-3: scan S rows (i), warps 4-7: scan S^T rows (j), warp 8: TMA producer, warp 9: MMA issuer.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int CI_TILE = 128 * 128;          // bytes of one [128][64 bf16] tile
constexpr int CI_IMG = 4 * CI_TILE;         // [kb 2][hi|lo][128][128 B]: a 128-row, K=128 operand image
constexpr int C2_THREADS = 320;
constexpr int C2_MAXP = 2560;               // valid positions per sample and side: 20 tiles of 128 (S = 20 sentences of L = 128, BASELINE configs[4])
constexpr int C2_MAXS = 512;                // sentences per sample

// Valid-position bookkeeping of one side of one sample.  Inputs that come out of ImprovedRnn have exactly-zero rows at and beyond each
// sentence's length (model.py:20); their affinity entries are exactly 0, so only the VALID rows are compacted into the operand
// images and multiplied, and "0" enters every maximum analytically (coattn_resolve_kernel).
//   meta[(b*2 + which)*2 + {0,1}] = {number of valid rows, first padded position (or -1)};  pos[b*P + c] = position of compact row c
// which = 0: giM rows (item side), 1: gu rows (user side).  cst = exclusive prefix sum of the sentence lengths (NULL: all P rows valid).
__global__ void __launch_bounds__(256) coattn_images_kernel(const float* __restrict__ giM, const float* __restrict__ gu, int P, int T,
                                                            const int* __restrict__ cst_a, int S_a, int L_a, const int* __restrict__ cst_b,
                                                            int S_b, int L_b, unsigned char* __restrict__ imgA,
                                                            unsigned char* __restrict__ imgB, float* __restrict__ n2a, float* __restrict__ n2b,
                                                            int* __restrict__ posA, int* __restrict__ posB, int* __restrict__ meta) {
  __shared__ int sb[C2_MAXS + 1];
  const int t = blockIdx.x, b = blockIdx.y, which = blockIdx.z, tid = threadIdx.x;
  const int* cst = which ? cst_b : cst_a;
  const int S_ = which ? S_b : S_a, L_ = which ? L_b : L_a;
  const float* src = (which ? gu : giM) + (size_t)b * P * D;
  unsigned char* img = (which ? imgB : imgA) + ((size_t)b * T + t) * CI_IMG;
  float* n2 = (which ? n2b : n2a) + (size_t)b * P;
  int* pos = (which ? posB : posA) + (size_t)b * P;
  int Pv = P;
  if (cst) {
    const int base0 = cst[(size_t)b * S_];
    for (int s_ = tid; s_ <= S_; s_ += 256) sb[s_] = cst[(size_t)b * S_ + s_] - base0;
    __syncthreads();
    Pv = sb[S_];
  }
  if (t == 0 && tid == 0) {
    int fp = -1;
    if (cst) for (int s_ = 0; s_ < S_; ++s_) { const int len = sb[s_ + 1] - sb[s_]; if (len < L_) { fp = s_ * L_ + len; break; } }
    meta[(b * 2 + which) * 2] = Pv;
    meta[(b * 2 + which) * 2 + 1] = fp;
  }
  if (t * 128 >= Pv) return;                           // nothing valid in this tile: the affinity kernel never loads it
  const int k = (tid & 15) * 4;
  for (int r = tid >> 4; r < 128; r += 16) {          // half-warp per row, both k-blocks
    const int c = t * 128 + r;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    int p = c;
    if (c < Pv) {
      if (cst) {                                        // sentence of compact row c: largest s with sb[s] <= c
        int lo = 0, hi = S_;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sb[mid] <= c) lo = mid; else hi = mid; }
        p = lo * L_ + (c - sb[lo]);
      }
      v0 = *reinterpret_cast<const float4*>(src + (size_t)p * D + k);
      v1 = *reinterpret_cast<const float4*>(src + (size_t)p * D + 64 + k);
    }
    store_split4(img, img + CI_TILE, r, k, v0);
    store_split4(img + 2 * CI_TILE, img + 3 * CI_TILE, r, k, v1);
    float q = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
    q += __shfl_xor_sync(0xffffffffu, q, 8); q += __shfl_xor_sync(0xffffffffu, q, 4);
    q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 1);
    if ((tid & 15) == 0 && c < Pv) { n2[c] = q; pos[c] = p; }
  }
}

struct C2Bars { uint64_t a_full, a_empty, b_full[2], b_empty[2], acc_full[2], acc_empty[2]; };

__global__ void __launch_bounds__(C2_THREADS, 1) coattn_affinity_tc2_kernel(const unsigned char* __restrict__ imgA, const unsigned char* __restrict__ imgB,
                                                                            const float* __restrict__ n2a, const float* __restrict__ n2b, int P, int T,
                                                                            const int* __restrict__ meta, float4* __restrict__ rc_v, int4* __restrict__ rc_i,
                                                                            float4* __restrict__ cc_v, int4* __restrict__ cc_i, int dbg) {
  extern __shared__ unsigned char raw[];
  __shared__ C2Bars bars;
  __shared__ uint32_t tmem_slot;
  // per-position state: the running maxima stay in shared memory (20 KB at the limit of 2560 positions); row norms are read from
  // global memory once per tile pair, and the column candidates of pass 1 are parked in the output arrays cc_v / cc_i themselves
  // (each column is owned by ONE thread for the whole kernel, so there is no sharing to order)
  __shared__ float s_rowmax[C2_MAXP], s_colmax[C2_MAXP];
  __shared__ float red[32];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_img = base;                  // giM tile of the current i-tile
  unsigned char* b_img = base + CI_IMG;         // 2 stages of gu tiles
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int Pa = meta[b * 4], Pb = meta[b * 4 + 2];               // valid rows of giM (i) / gu (j) of this sample
  const int Ta = (Pa + 127) >> 7, Tb = (Pb + 127) >> 7;

  if (tid == 0) {
    mbar_init(&bars.a_full, 1); mbar_init(&bars.a_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1); mbar_init(&bars.acc_full[s], 1); mbar_init(&bars.acc_empty[s], 256); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(&tmem_slot, 512);
  float amax = 0.f;
  for (int p = tid; p < Pa; p += C2_THREADS) amax = fmaxf(amax, n2a[(size_t)b * P + p]);
  for (int p = tid; p < Tb * 128; p += C2_THREADS) {
    s_colmax[p] = -INFINITY;
    if (p < Pb) {
      cc_v[(size_t)b * P + p] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      cc_i[(size_t)b * P + p] = make_int4(-1, -1, -1, -1);
    }
  }
  amax = sqrtf(amax);
  amax = block_max(amax, red);              // max_i |giM_i| of the sample: bound for the column-side tolerance
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue: S rows (warps 0-3) / S^T rows (warps 4-7)
    const bool is_t = warp >= 4;
    const int q = warp & 3, r = q * 32 + lane;
    float m = -INFINITY;
    Cand4 cd;                 // candidates of this thread's row (S side: lives across the j-tiles) / column (S^T side: parked in shared memory)
    cd.init();
    float thr = -INFINITY;    // candidate threshold (final maximum - tau), kept in a register: Cand4 itself lives on the stack for push()
    int n = 0;
    const int Pown = is_t ? Pb : Pa;
    for (int pass = 0; pass < 2; ++pass)
      for (int it = 0; it < Ta; ++it)
#pragma unroll 1
        for (int jt = 0; jt < Tb; ++jt, ++n) {
          const int s = n & 1;
          const int own = (is_t ? jt : it) * 128 + r;                // the (compact) row i / column j this thread scans
          const int o0 = (is_t ? it : jt) * 128;                     // first index of the scanned direction
          const int nv = min(128, (is_t ? Pa : Pb) - o0);
          const float own_norm = own < Pown ? sqrtf((is_t ? n2b : n2a)[(size_t)b * P + own]) : 0.f;
          const float tau = is_t ? CA_EPS * own_norm * amax : CA_EPS * own_norm * CA_GNORM;
          if (pass == 1) {
            if (is_t) {
              float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); int4 id = make_int4(-1, -1, -1, -1);
              if (own < Pown) { v = cc_v[(size_t)b * P + own]; id = cc_i[(size_t)b * P + own]; }
              cd.v[0] = v.x; cd.v[1] = v.y; cd.v[2] = v.z; cd.v[3] = v.w;
              cd.id[0] = id.x; cd.id[1] = id.y; cd.id[2] = id.z; cd.id[3] = id.w;
              thr = s_colmax[own] - tau;
            } else if (jt == 0) {
              cd.init();
              thr = s_rowmax[own] - tau;
            }
          } else if (!is_t && jt == 0) {
            m = -INFINITY;
          }
          mbar_wait(&bars.acc_full[s], (n >> 1) & 1);
          tc_fence_after();
          const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + s * 256 + (is_t ? 128 : 0);
          float tm = -INFINITY;
          const bool scan = pass == 1 && own < Pown;
          const bool skip = (dbg & 1) != 0;
          uint32_t ra[2][16];
          const int nch = (nv + 15) >> 4;                           // chunks of 16 columns that hold valid entries
          if (!skip) tmem_ld16_issue(t0, ra[0]);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            if (skip || ch >= nch) break;
            tmem_ld_wait();
            if (ch + 1 < nch) tmem_ld16_issue(t0 + (ch + 1) * 16, ra[(ch + 1) & 1]);
            const int left = nv - ch * 16;
            float cm = -INFINITY;                                   // chunk maximum over the valid columns
#pragma unroll
            for (int c = 0; c < 16; ++c) cm = fmaxf(cm, c < left ? __uint_as_float(ra[ch & 1][c]) : -INFINITY);
            tm = fmaxf(tm, cm);
            if (scan && cm > thr) {                              // rare: something in this chunk is within tau of the maximum
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float x = __uint_as_float(ra[ch & 1][c]);
                if (c < left && x > thr) cd.push(x, o0 + ch * 16 + c);
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&bars.acc_empty[s]);
          if (pass == 0) {
            if (is_t) s_colmax[own] = fmaxf(s_colmax[own], tm);
            else { m = fmaxf(m, tm); if (jt == Tb - 1) s_rowmax[own] = m; }
          } else if (own < Pown) {
            const bool done = is_t ? (it == Ta - 1) : (jt == Tb - 1);
            if (done) {
              const size_t o = (size_t)b * P + own;
#pragma unroll
              for (int k = 1; k < 4; ++k) if (cd.v[k] < cd.v[0] - tau) { cd.v[k] = -INFINITY; cd.id[k] = -1; }
              (is_t ? cc_v : rc_v)[o] = make_float4(cd.v[0], cd.v[1], cd.v[2], cd.v[3]);
              (is_t ? cc_i : rc_i)[o] = make_int4(cd.id[0], cd.id[1], cd.id[2], cd.id[3]);
            } else if (is_t) {
              cc_v[(size_t)b * P + own] = make_float4(cd.v[0], cd.v[1], cd.v[2], cd.v[3]);
              cc_i[(size_t)b * P + own] = make_int4(cd.id[0], cd.id[1], cd.id[2], cd.id[3]);
            }
          }
        }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int n = 0, na = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int it = 0; it < Ta; ++it, ++na) {
          if (na > 0) mbar_wait(&bars.a_empty, (na - 1) & 1);
          mbar_arrive_expect_tx(&bars.a_full, CI_IMG);
          bulk_copy_g2s(a_img, imgA + ((size_t)b * T + it) * CI_IMG, CI_IMG, &bars.a_full);
          for (int jt = 0; jt < Tb; ++jt, ++n) {
            const int s = n & 1;
            if (n >= 2) mbar_wait(&bars.b_empty[s], ((n >> 1) - 1) & 1);
            mbar_arrive_expect_tx(&bars.b_full[s], CI_IMG);
            bulk_copy_g2s(b_img + s * CI_IMG, imgB + ((size_t)b * T + jt) * CI_IMG, CI_IMG, &bars.b_full[s]);
          }
        }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, the elected lane issues)
    const uint32_t el = elect_one_sync();
    const uint32_t a0 = smem_u32(a_img);
    int n = 0, na = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int it = 0; it < Ta; ++it, ++na) {
        mbar_wait(&bars.a_full, na & 1);
        // only the valid columns of a partly filled tile are multiplied (N rounded up to 16; the scans never read beyond them)
        const uint32_t idesc_t = idesc_bf16(128, (min(128, Pa - it * 128) + 15) & ~15);        // S^T: columns = rows i of this tile
        for (int jt = 0; jt < Tb; ++jt, ++n) {
          const uint32_t idesc_s = idesc_bf16(128, (min(128, Pb - jt * 128) + 15) & ~15);      // S: columns = rows j of this tile
          const int s = n & 1;
          mbar_wait(&bars.b_full[s], (n >> 1) & 1);
          if (n >= 2) mbar_wait(&bars.acc_empty[s], ((n >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t b0 = smem_u32(b_img + s * CI_IMG);
          const uint32_t d1 = tmem + s * 256, d2 = d1 + 128;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ah = smem_desc_sw128(a0 + kb * 2 * CI_TILE), al = smem_desc_sw128(a0 + kb * 2 * CI_TILE + CI_TILE);
            const uint64_t bh = smem_desc_sw128(b0 + kb * 2 * CI_TILE), bl = smem_desc_sw128(b0 + kb * 2 * CI_TILE + CI_TILE);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (dbg & 4) break;
              const uint64_t o = (uint64_t)(kk * 2);
              const uint32_t accf = (kb | kk) != 0;
              umma_bf16_e(el, d1, ah + o, bh + o, idesc_s, accf);      // S   (rows i, cols j)
              umma_bf16_e(el, d1, ah + o, bl + o, idesc_s, 1);
              umma_bf16_e(el, d1, al + o, bh + o, idesc_s, 1);
              umma_bf16_e(el, d2, bh + o, ah + o, idesc_t, accf);      // S^T (rows j, cols i)
              umma_bf16_e(el, d2, bh + o, al + o, idesc_t, 1);
              umma_bf16_e(el, d2, bl + o, ah + o, idesc_t, 1);
            }
          }
          umma_commit_e(el, &bars.b_empty[s]);
          umma_commit_e(el, &bars.acc_full[s]);
          if (jt == Tb - 1) umma_commit_e(el, &bars.a_empty);
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// exact fp32 <giM[i], gu[j]> by one warp (lane holds a float4 of each row)
__device__ __forceinline__ float exact_dot(const float* __restrict__ giM_b, const float* __restrict__ gu_b, int i, int j, int lane) {
  const float4 x = *reinterpret_cast<const float4*>(giM_b + (size_t)i * D + lane * 4);
  const float4 y = *reinterpret_cast<const float4*>(gu_b + (size_t)j * D + lane * 4);
  return warp_sum(x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w);
}

// resolve the candidates with exact fp32 re-scoring, then tanh / softmax over all P / pooling (model.py:52-55).
// Works on the compact (valid-row) index space of the affinity kernel and writes results at the original positions.  When the
// other side has padded rows, every maximum also competes with their exact 0: a maximum below 0 becomes (0, first padded
// position) - its gradient lands on a zero row and vanishes, exactly as in the dense computation.  Padded positions of the own
// side get value tanh(0) = 0 and take part in the softmax.
__global__ void __launch_bounds__(256) coattn_resolve_kernel(const float* __restrict__ giM, const float* __restrict__ gu, const float* __restrict__ gi,
                                                             const float* __restrict__ n2a, const float* __restrict__ n2b, int P,
                                                             const int* __restrict__ meta, const int* __restrict__ posA, const int* __restrict__ posB,
                                                             const float4* __restrict__ rc_v, const int4* __restrict__ rc_i,
                                                             const float4* __restrict__ cc_v, const int4* __restrict__ cc_i,
                                                             float* __restrict__ soft_u, float* __restrict__ soft_i, float* __restrict__ t_u,
                                                             float* __restrict__ t_i, int* __restrict__ arg_u, int* __restrict__ arg_i,
                                                             float* __restrict__ atte_u, float* __restrict__ atte_i) {
  extern __shared__ __align__(16) float smem[];
  const int P4 = (P + 3) & ~3;
  float* sp = smem;
  float* red = smem + P4;
  float4* part = reinterpret_cast<float4*>(smem + P4 + 32);
  const int b = blockIdx.x, side = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* giM_b = giM + (size_t)b * P * D;
  const float* gu_b = gu + (size_t)b * P * D;
  const float* g = (side ? gi : gu) + (size_t)b * P * D;
  float* soft = (side ? soft_i : soft_u) + (size_t)b * P;
  float* tv = (side ? t_i : t_u) + (size_t)b * P;
  int* arg = (side ? arg_i : arg_u) + (size_t)b * P;
  const int Pa = meta[b * 4], Pb = meta[b * 4 + 2];
  const int Po = side ? Pa : Pb;                                  // own valid count (side 1 = rows i, side 0 = columns j)
  const int* pos_own = (side ? posA : posB) + (size_t)b * P;
  const int* pos_oth = (side ? posB : posA) + (size_t)b * P;
  const int zero_at = meta[b * 4 + (side ? 3 : 1)];               // first padded position of the OTHER side, -1 if it has none
  const float* n2_own = (side ? n2a : n2b) + (size_t)b * P;
  const float4* cv = (side ? rc_v : cc_v) + (size_t)b * P;
  const int4* ci = (side ? rc_i : cc_i) + (size_t)b * P;
  // max_i |giM_i| of this sample (bound for the column-side tolerance); squared row norms come from coattn_images_kernel
  float gmax = 0.f;
  if (side == 0) {
    for (int p = tid; p < Pa; p += 256) gmax = fmaxf(gmax, n2a[(size_t)b * P + p]);
    gmax = sqrtf(block_max(gmax, red));
  }
  __shared__ int s_work[C2_MAXP], s_nwork;
  if (tid == 0) s_nwork = 0;
  for (int p = tid; p < P; p += 256) { tv[p] = 0.f; arg[p] = 0; }          // padded positions of the own side (valid ones overwritten below)
  __syncthreads();
  // (1) one thread per row / column: count the candidates that survive the tolerance; a unique survivor keeps its tensor-core value
  //     (the common case), the others go to a work list;  (2) one warp per work-list entry: exact fp32 re-scoring
  for (int p = tid; p < Po; p += 256) {
    const float tau = CA_EPS * sqrtf(n2_own[p]) * (side ? CA_GNORM : gmax);
    const float4 v = cv[p];
    const int4 id = ci[p];
    const float head = v.x;
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const int ii[4] = {id.x, id.y, id.z, id.w};
    float lone_v = 0.f;
    int lone_i = 0, n_surv = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ii[k] >= 0 && vv[k] >= head - tau) { ++n_surv; lone_v = vv[k]; lone_i = ii[k]; }
    const bool near_zero = zero_at >= 0 && fabsf(lone_v) <= tau;             // undecided against the padded rows' exact 0
    if (n_surv <= 1 && !near_zero) {
      const bool zero_wins = zero_at >= 0 && lone_v < 0.f;
      const float t = zero_wins ? 0.f : tanhf(lone_v);
      sp[p] = t; tv[pos_own[p]] = t; arg[pos_own[p]] = zero_wins ? zero_at : pos_oth[lone_i];
    } else {
      s_work[atomicAdd(&s_nwork, 1)] = p;
    }
  }
  __syncthreads();
  for (int wi = warp; wi < s_nwork; wi += 8) {
    const int p = s_work[wi];
    const float tau = CA_EPS * sqrtf(n2_own[p]) * (side ? CA_GNORM : gmax);
    const float4 v = cv[p];
    const int4 id = ci[p];
    const float head = v.x;
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const int ii[4] = {id.x, id.y, id.z, id.w};
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ii[k] < 0 || vv[k] < head - tau) continue;
      const int po = pos_oth[ii[k]];
      const float ex = side ? exact_dot(giM_b, gu_b, pos_own[p], po, lane) : exact_dot(giM_b, gu_b, po, pos_own[p], lane);
      if (ex > best || (ex == best && po < besti)) { best = ex; besti = po; }
    }
    if (zero_at >= 0 && (best < 0.f || (best == 0.f && zero_at < besti))) { best = 0.f; besti = zero_at; }    // first maximum wins
    if (lane == 0) {
      const float t = tanhf(best);
      sp[p] = t; tv[pos_own[p]] = t; arg[pos_own[p]] = besti;
    }
  }
  __syncthreads();
  const int n_pad = P - Po;                                       // own padded positions: value tanh(0) = 0 each
  float mx = n_pad > 0 ? 0.f : -INFINITY;
  for (int p = tid; p < Po; p += 256) mx = fmaxf(mx, sp[p]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int p = tid; p < Po; p += 256) { const float e = expf(sp[p] - mx); sp[p] = e; sum += e; }
  sum = block_sum(sum, red);
  const float e_pad = expf(-mx);
  sum += (float)n_pad * e_pad;
  const float inv = 1.f / sum;
  if (n_pad > 0) {
    const float s_pad = e_pad * inv;
    for (int p = tid; p < P; p += 256) soft[p] = s_pad;
    __syncthreads();
  }
  for (int p = tid; p < Po; p += 256) { const float s = sp[p] * inv; sp[p] = s; soft[pos_own[p]] = s; }
  __syncthreads();
  const int c4 = tid & 31, pg = tid >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = pg; p < Po; p += 8) {
    const float s = sp[p];
    const float4 v = *reinterpret_cast<const float4*>(g + (size_t)pos_own[p] * D + c4 * 4);
    acc.x += s * v.x; acc.y += s * v.y; acc.z += s * v.z; acc.w += s * v.w;
  }
  part[pg * 32 + c4] = acc;
  __syncthreads();
  if (tid < 32) {
    float4 r = part[tid];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 v = part[k * 32 + tid]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    *reinterpret_cast<float4*>((side ? atte_i : atte_u) + (size_t)b * D + tid * 4) = r;
  }
}

}  // namespace umpr

using namespace umpr;

// scratch (16-byte aligned, umpr_workspace_bytes("coattn_fwd_tc", B, P): sized for T = ceil(min(P, 2560) / 128)): [imgA B*T*64 KB][imgB B*T*64 KB][n2a B*P f32][n2b B*P f32][posA B*P i32][posB B*P i32][meta 4*B i32]
// [rc_v][rc_i][cc_v][cc_i] (float4/int4 per (b,p)),  T = ceil(pv_max/128): 2*B*T*65536 + 4*B*P*4 + 16*B + 4*B*P*16 + 256 bytes.
// cst_u / cst_i (optional, both or neither): exclusive prefix sums of the sentence lengths of the user / item side (S_x sentences of
// L_x positions per sample, S_x*L_x == P) for inputs whose rows beyond each sentence's length are exactly zero - only the valid rows
// are multiplied.  NULL: every row is treated as valid.
// pv_max (with length tables): the largest number of valid rows any sample has on either side - the limit of 512 then applies to it,
// not to P (long padded sentences that are mostly padding), and the operand images are sized by it; 0 = P.
extern "C" int umpr_coattn_fwd_tc(const float* gu, const float* gi, const float* giM, int B, int P, const int32_t* cst_u, int S_u, int L_u,
                                  const int32_t* cst_i, int S_i, int L_i, int pv_max, void* scratch, float* soft_u, float* soft_i, float* t_u,
                                  float* t_i, int32_t* arg_u, int32_t* arg_i, float* atte_u, float* atte_i, void* stream) {
  if (B <= 0 || P <= 0) return 0;
  if (B > 65535) return fail_arg("coattn_fwd_tc: batch %d > 65535", B);
  if (!cst_u || pv_max <= 0 || pv_max > P) pv_max = P;
  if (pv_max > C2_MAXP) return fail_arg("coattn_fwd_tc: %d valid positions per sample > %d (use umpr_coattn_fwd)", pv_max, C2_MAXP);
  if (P > 16384 || S_u > C2_MAXS || S_i > C2_MAXS) return fail_arg("coattn_fwd_tc: P=%d / S=%d,%d too large", P, S_u, S_i);
  if ((cst_u == nullptr) != (cst_i == nullptr)) return fail_arg("coattn_fwd_tc: length tables must be given for both sides or neither");
  if (cst_u && (S_u * L_u != P || S_i * L_i != P || S_u < 1 || S_i < 1))
    return fail_arg("coattn_fwd_tc: S*L must equal P=%d on both sides (got %d*%d, %d*%d)", P, S_u, L_u, S_i, L_i);
  const int T = (pv_max + 127) / 128;
  unsigned char* imgA = reinterpret_cast<unsigned char*>(scratch);
  unsigned char* imgB = imgA + (size_t)B * T * CI_IMG;
  float* n2a = reinterpret_cast<float*>(imgB + (size_t)B * T * CI_IMG);
  float* n2b = n2a + (size_t)B * P;
  int* posA = reinterpret_cast<int*>(n2b + (size_t)B * P);
  int* posB = posA + (size_t)B * P;
  int* meta = posB + (size_t)B * P;
  uintptr_t q = (reinterpret_cast<uintptr_t>(meta + (size_t)4 * B) + 15) & ~uintptr_t(15);
  float4* rc_v = reinterpret_cast<float4*>(q);
  int4* rc_i = reinterpret_cast<int4*>(rc_v + (size_t)B * P);
  float4* cc_v = reinterpret_cast<float4*>(rc_i + (size_t)B * P);
  int4* cc_i = reinterpret_cast<int4*>(cc_v + (size_t)B * P);
  cudaStream_t st = (cudaStream_t)stream;
  coattn_images_kernel<<<dim3(T, B, 2), 256, 0, st>>>(giM, gu, P, T, cst_i, S_i, L_i, cst_u, S_u, L_u, imgA, imgB, n2a, n2b, posA, posB, meta);
  if (int rc = check_launch("coattn_images")) return rc;
  const int sm1 = 3 * CI_IMG + 1024;
  cudaError_t e = cudaFuncSetAttribute(coattn_affinity_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1);
  if (e != cudaSuccess) { set_error("coattn_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  coattn_affinity_tc2_kernel<<<B, C2_THREADS, sm1, st>>>(imgA, imgB, n2a, n2b, P, T, meta, rc_v, rc_i, cc_v, cc_i, dbg_flags());
  if (int rc = check_launch("coattn_affinity_tc2")) return rc;
  const size_t sm2 = sizeof(float) * (((P + 3) & ~3) + 32) + sizeof(float4) * 8 * 32;
  if (sm2 > 48 * 1024) cudaFuncSetAttribute(coattn_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  coattn_resolve_kernel<<<dim3(B, 2), 256, sm2, st>>>(giM, gu, gi, n2a, n2b, P, meta, posA, posB, rc_v, rc_i, cc_v, cc_i, soft_u, soft_i, t_u,
                                                     t_i, arg_u, arg_i, atte_u, atte_i);
  return check_launch("coattn_resolve");
}
