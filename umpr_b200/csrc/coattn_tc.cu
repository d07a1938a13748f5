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
  // keep the 4 largest distinct values, sorted descending
  __device__ __forceinline__ void push(float x, int idx) {
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

__global__ void __launch_bounds__(CA_THREADS, 1) coattn_affinity_tc_kernel(const float* __restrict__ giM, const float* __restrict__ gu,
                                                                           int P, int n_it, float4* __restrict__ rc_v, int4* __restrict__ rc_i,
                                                                           float4* __restrict__ cc_v, int4* __restrict__ cc_i) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t b_full[2], b_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float a_norm2[128];          // |giM_i|^2 of the resident tile
  __shared__ float b_norm2[4][128];       // |gu_j|^2 of the staged tiles
  __shared__ float norm_red[4];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int TILE = 128 * 128;            // bytes of one [128][64 bf16] tile
  unsigned char* a_res = base;               // [kb 2][hi|lo][128][128 B]  giM tile
  unsigned char* b_stg = base + 4 * TILE;    // 2 stages of the same shape   gu tiles
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, it_ = blockIdx.x, i0 = it_ * 128;
  const int ni = min(128, P - i0);
  const int n_jt = (P + 127) / 128;
  const float* giM_b = giM + (size_t)b * P * D;
  const float* gu_b = gu + (size_t)b * P * D;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], 128); mbar_init(&b_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, 512);
  if (tid < 256) {
    for (int r = tid >> 4; r < 128; r += 16) {          // half-warp per row, both k-blocks
      const int k = (tid & 15) * 4;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (r < ni) {
        v0 = *reinterpret_cast<const float4*>(giM_b + (size_t)(i0 + r) * D + k);
        v1 = *reinterpret_cast<const float4*>(giM_b + (size_t)(i0 + r) * D + 64 + k);
      }
      store_split4(a_res, a_res + TILE, r, k, v0);
      store_split4(a_res + 2 * TILE, a_res + 3 * TILE, r, k, v1);
      float t = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
      t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
      if ((tid & 15) == 0) a_norm2[r] = t;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ---------------------------------------------------------------- loaders: gu tiles
    for (int jt = 0; jt < n_jt; ++jt) {
      const int s = jt & 1;
      if (jt >= 2) mbar_wait(&b_empty[s], ((jt >> 1) - 1) & 1);
      unsigned char* st = b_stg + s * 4 * TILE;
      const int j0 = jt * 128, nj = min(128, P - j0);
      float n2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) n2[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        float4 va[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid, r = idx >> 4, k = (idx & 15) * 4;
          va[i] = r < nj ? *reinterpret_cast<const float4*>(gu_b + (size_t)(j0 + r) * D + kb * 64 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = i * 128 + tid;
          store_split4(st + kb * 2 * TILE, st + kb * 2 * TILE + TILE, idx >> 4, (idx & 15) * 4, va[i]);
          n2[i] += va[i].x * va[i].x + va[i].y * va[i].y + va[i].z * va[i].z + va[i].w * va[i].w;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {       // the 16 lanes of a half-warp share row (i*128+tid)>>4
        float t = n2[i];
        t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 4);
        t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 1);
        if ((lane & 15) == 0) b_norm2[jt & 3][(i * 128 + tid) >> 4] = t;
      }
      fence_async_smem();
      mbar_arrive(&b_full[s]);
    }
  } else if (warp == 4) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, 128);
      const uint32_t a0 = smem_u32(a_res);
      for (int jt = 0; jt < n_jt; ++jt) {
        const int s = jt & 1;
        if (jt >= 2) mbar_wait(&acc_empty[s], ((jt >> 1) - 1) & 1);
        mbar_wait(&b_full[s], (jt >> 1) & 1);
        tc_fence_after();
        const uint32_t b0 = smem_u32(b_stg + s * 4 * TILE);
        const uint32_t d1 = tmem + s * 256, d2 = d1 + 128;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ah = smem_desc_sw128(a0 + kb * 2 * TILE), al = smem_desc_sw128(a0 + kb * 2 * TILE + TILE);
          const uint64_t bh = smem_desc_sw128(b0 + kb * 2 * TILE), bl = smem_desc_sw128(b0 + kb * 2 * TILE + TILE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            const uint32_t accf = (kb | kk) != 0;
            umma_bf16(d1, ah + o, bh + o, idesc, accf);      // S   (rows i, cols j)
            umma_bf16(d1, ah + o, bl + o, idesc, 1);
            umma_bf16(d1, al + o, bh + o, idesc, 1);
            umma_bf16(d2, bh + o, ah + o, idesc, accf);      // S^T (rows j, cols i)
            umma_bf16(d2, bh + o, al + o, idesc, 1);
            umma_bf16(d2, bl + o, ah + o, idesc, 1);
          }
        }
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: candidate scans
    const int q = warp & 3, row = q * 32 + lane;
    // error bounds: tau_i = eps * |giM_i| * max|gu_j| ; tau_j = eps * |gu_j| * max_{i in tile} |giM_i|
    const float nrm = sqrtf(a_norm2[row]);
    const float wmax = warp_max(nrm);
    if (lane == 0) norm_red[q] = wmax;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float tile_norm = fmaxf(fmaxf(norm_red[0], norm_red[1]), fmaxf(norm_red[2], norm_red[3]));
    const float tau_i = CA_EPS * nrm * CA_GNORM;
    Cand4 rc;
    rc.init();
    for (int jt = 0; jt < n_jt; ++jt) {
      const int s = jt & 1, j0 = jt * 128, nj = min(128, P - j0);
      mbar_wait(&acc_full[s], (jt >> 1) & 1);
      tc_fence_after();
      const uint32_t t1 = tmem + ((uint32_t)(q * 32) << 16) + s * 256, t2 = t1 + 128;
      const float tau_j = CA_EPS * sqrtf(b_norm2[jt & 3][row]) * tile_norm;
      Cand4 cc;
      cc.init();
      // phase A: branch-free tile maxima, so that candidates are judged against an up-to-date head (few slow-path events).
      // TMEM loads are software-pipelined: the next 16-column chunk of S and S^T is in flight while the current one is scanned.
      float m1 = -INFINITY, m2 = -INFINITY;
      uint32_t ra[2][16], rb[2][16];
      tmem_ld16_issue(t1, ra[0]);
      tmem_ld16_issue(t2, rb[0]);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        tmem_ld_wait();
        if (ch + 1 < 8) { tmem_ld16_issue(t1 + (ch + 1) * 16, ra[(ch + 1) & 1]); tmem_ld16_issue(t2 + (ch + 1) * 16, rb[(ch + 1) & 1]); }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          m1 = fmaxf(m1, ch * 16 + c < nj ? __uint_as_float(ra[ch & 1][c]) : -INFINITY);
          m2 = fmaxf(m2, ch * 16 + c < ni ? __uint_as_float(rb[ch & 1][c]) : -INFINITY);
        }
      }
      rc.thr = fmaxf(rc.thr, m1 - tau_i);
      cc.thr = m2 - tau_j;
      // phase B: collect the candidates within tau of the head
      tmem_ld16_issue(t1, ra[0]);
      tmem_ld16_issue(t2, rb[0]);
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        tmem_ld_wait();
        float v1[16], v2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { v1[c] = __uint_as_float(ra[0][c]); v2[c] = __uint_as_float(rb[0][c]); }
        if (ch + 1 < 8) { tmem_ld16_issue(t1 + (ch + 1) * 16, ra[0]); tmem_ld16_issue(t2 + (ch + 1) * 16, rb[0]); }
        if (row < ni) { rc.scan8(v1, j0 + ch * 16, nj - ch * 16, tau_i); rc.scan8(v1 + 8, j0 + ch * 16 + 8, nj - ch * 16 - 8, tau_i); }
        if (row < nj) { cc.scan8(v2, i0 + ch * 16, ni - ch * 16, tau_j); cc.scan8(v2 + 8, i0 + ch * 16 + 8, ni - ch * 16 - 8, tau_j); }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[s]);
      if (row < nj) {
        const size_t o = ((size_t)b * n_it + it_) * P + j0 + row;
#pragma unroll
        for (int k = 1; k < 4; ++k) if (cc.v[k] < cc.v[0] - tau_j) { cc.v[k] = -INFINITY; cc.id[k] = -1; }
        cc_v[o] = make_float4(cc.v[0], cc.v[1], cc.v[2], cc.v[3]);
        cc_i[o] = make_int4(cc.id[0], cc.id[1], cc.id[2], cc.id[3]);
      }
    }
    if (row < ni) {
      const size_t o = (size_t)b * P + i0 + row;
#pragma unroll
      for (int k = 1; k < 4; ++k) if (rc.v[k] < rc.v[0] - tau_i) { rc.v[k] = -INFINITY; rc.id[k] = -1; }
      rc_v[o] = make_float4(rc.v[0], rc.v[1], rc.v[2], rc.v[3]);
      rc_i[o] = make_int4(rc.id[0], rc.id[1], rc.id[2], rc.id[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// exact fp32 <giM[i], gu[j]> by one warp (lane holds a float4 of each row)
__device__ __forceinline__ float exact_dot(const float* __restrict__ giM_b, const float* __restrict__ gu_b, int i, int j, int lane) {
  const float4 x = *reinterpret_cast<const float4*>(giM_b + (size_t)i * D + lane * 4);
  const float4 y = *reinterpret_cast<const float4*>(gu_b + (size_t)j * D + lane * 4);
  return warp_sum(x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w);
}

// resolve the candidates with exact fp32 re-scoring, then tanh / softmax over all P / pooling (model.py:52-55)
__global__ void __launch_bounds__(256) coattn_resolve_kernel(const float* __restrict__ giM, const float* __restrict__ gu, const float* __restrict__ gi,
                                                             int P, int n_it, const float4* __restrict__ rc_v, const int4* __restrict__ rc_i,
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
  // max_i |giM_i| of this sample (bound for the column-side tolerance)
  float gmax = 0.f;
  if (side == 0) {
    for (int p = warp; p < P; p += 8) {
      const float4 x = *reinterpret_cast<const float4*>(giM_b + (size_t)p * D + lane * 4);
      gmax = fmaxf(gmax, warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w));
    }
    gmax = sqrtf(block_max(gmax, red));
  }
  for (int p = warp; p < P; p += 8) {
    // tolerance of this row (side 1: i = p) / column (side 0: j = p): the same bound the producer used, or looser
    const float4 own = *reinterpret_cast<const float4*>((side ? giM_b : gu_b) + (size_t)p * D + lane * 4);
    const float nrm = sqrtf(warp_sum(own.x * own.x + own.y * own.y + own.z * own.z + own.w * own.w));
    const float tau = CA_EPS * nrm * (side ? CA_GNORM : gmax);
    const int nl = side ? 1 : n_it;
    float head = -INFINITY;
    for (int l = 0; l < nl; ++l) head = fmaxf(head, side ? rc_v[(size_t)b * P + p].x : cc_v[((size_t)b * n_it + l) * P + p].x);
    float best = -INFINITY, lone_v = 0.f;
    int besti = 0x7fffffff, lone_i = 0, n_surv = 0;
    for (int pass = 0; pass < 2; ++pass) {           // pass 0: count survivors; pass 1: exact re-scoring if more than one
      if (pass == 1 && n_surv <= 1) break;
      for (int l = 0; l < nl; ++l) {
        const size_t o = side ? (size_t)b * P + p : ((size_t)b * n_it + l) * P + p;
        const float4 v = side ? rc_v[o] : cc_v[o];
        const int4 id = side ? rc_i[o] : cc_i[o];
        const float vv[4] = {v.x, v.y, v.z, v.w};
        const int ii[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ii[k] < 0 || vv[k] < head - tau) continue;
          if (pass == 0) { ++n_surv; lone_v = vv[k]; lone_i = ii[k]; }
          else {
            const float ex = side ? exact_dot(giM_b, gu_b, p, ii[k], lane) : exact_dot(giM_b, gu_b, ii[k], p, lane);
            if (ex > best || (ex == best && ii[k] < besti)) { best = ex; besti = ii[k]; }
          }
        }
      }
    }
    if (n_surv <= 1) { best = lone_v; besti = lone_i; }     // unique candidate: keep the tensor-core value
    if (lane == 0) {
      const float t = tanhf(best);
      sp[p] = t; tv[p] = t; arg[p] = besti;
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int p = tid; p < P; p += 256) mx = fmaxf(mx, sp[p]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int p = tid; p < P; p += 256) { const float e = expf(sp[p] - mx); sp[p] = e; sum += e; }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int p = tid; p < P; p += 256) { const float s = sp[p] * inv; sp[p] = s; soft[p] = s; }
  __syncthreads();
  const int c4 = tid & 31, pg = tid >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = pg; p < P; p += 8) {
    const float s = sp[p];
    const float4 v = *reinterpret_cast<const float4*>(g + (size_t)p * D + c4 * 4);
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

// scratch: float4 rc_v[B*P], int4 rc_i[B*P], float4 cc_v[B*n_it*P], int4 cc_i[B*n_it*P], n_it = ceil(P/128)  -> 32*B*P*(1+n_it) bytes
extern "C" int umpr_coattn_fwd_tc(const float* gu, const float* gi, const float* giM, int B, int P, void* scratch, float* soft_u,
                                  float* soft_i, float* t_u, float* t_i, int32_t* arg_u, int32_t* arg_i, float* atte_u, float* atte_i,
                                  void* stream) {
  if (B <= 0 || P <= 0) return 0;
  if (B > 65535) return fail_arg("coattn_fwd_tc: batch %d > 65535", B);
  const int n_it = (P + 127) / 128;
  float4* rc_v = reinterpret_cast<float4*>(scratch);
  int4* rc_i = reinterpret_cast<int4*>(rc_v + (size_t)B * P);
  float4* cc_v = reinterpret_cast<float4*>(rc_i + (size_t)B * P);
  int4* cc_i = reinterpret_cast<int4*>(cc_v + (size_t)B * n_it * P);
  const int sm1 = 12 * 128 * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(coattn_affinity_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1);
  if (e != cudaSuccess) { set_error("coattn_tc smem: %s", cudaGetErrorString(e)); return (int)e; }
  coattn_affinity_tc_kernel<<<dim3(n_it, B), CA_THREADS, sm1, (cudaStream_t)stream>>>(giM, gu, P, n_it, rc_v, rc_i, cc_v, cc_i);
  if (int rc = check_launch("coattn_affinity_tc")) return rc;
  const size_t sm2 = sizeof(float) * (((P + 3) & ~3) + 32) + sizeof(float4) * 8 * 32;
  if (sm2 > 200 * 1024) return fail_arg("coattn_fwd_tc: P=%d too large", P);
  if (sm2 > 48 * 1024) cudaFuncSetAttribute(coattn_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  coattn_resolve_kernel<<<dim3(B, 2), 256, sm2, (cudaStream_t)stream>>>(giM, gu, gi, P, n_it, rc_v, rc_i, cc_v, cc_i, soft_u, soft_i, t_u,
                                                                       t_i, arg_u, arg_i, atte_u, atte_i);
  return check_launch("coattn_resolve");
}
