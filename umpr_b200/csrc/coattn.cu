// R-Net co-attention (reference src/model.py:50-55), flash-style: the (P x P) affinity matrix
// A = tanh(gi · M · gu^T) is never written to HBM.  tanh is monotone, so the row/column maxima are taken on
// the pre-activation and tanh is applied to 2P numbers per sample instead of P^2.
//   kernel 1: per (sample, 128-row tile of giM): all column tiles of gu -> packed (max, argmax) per row / column
//   kernel 2: per (sample, side): tanh, softmax over all P (unmasked, model.py:52-53), pooling atte = g^T soft
//   kernel 3: backward; dA is non-zero at <= 2P arg-max entries per sample, so it is a gather + scatter-add
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

constexpr int CT = 128;          // tile edge
constexpr int CLD = CT + 4;      // 132

// load a [rows<=128][128] row-major tile (k contiguous) into k-major shared memory T[k][row]; rows >= nvalid are zero
__device__ __forceinline__ void load_tile_kmajor(float* T, const float* __restrict__ src, int nvalid, int tid) {
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int e = tid + i * 256;
    const int k4lo = e & 3, m = (e >> 2) & 127, k4 = (e >> 9) * 4 + k4lo;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < nvalid) v = *reinterpret_cast<const float4*>(src + (size_t)m * D + k4 * 4);
    T[(k4 * 4 + 0) * CLD + m] = v.x; T[(k4 * 4 + 1) * CLD + m] = v.y;
    T[(k4 * 4 + 2) * CLD + m] = v.z; T[(k4 * 4 + 3) * CLD + m] = v.w;
  }
}

__global__ void __launch_bounds__(256, 1) coattn_affinity_kernel(const float* __restrict__ giM, const float* __restrict__ gu, int P,
                                                                 unsigned long long* __restrict__ rowkey,
                                                                 unsigned long long* __restrict__ colkey) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // [128 k][132]  giM tile
  float* Bs = smem + D * CLD;       // [128 k][132]  gu tile
  unsigned long long* cred = reinterpret_cast<unsigned long long*>(Bs + D * CLD);   // [8 warps][128]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, i0 = blockIdx.x * CT;
  const int ni = min(CT, P - i0);
  load_tile_kmajor(As, giM + ((size_t)b * P + i0) * D, ni, tid);
  float rmax[8];
  int rarg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { rmax[i] = -INFINITY; rarg[i] = 0; }

  for (int j0 = 0; j0 < P; j0 += CT) {
    const int nj = min(CT, P - j0);
    __syncthreads();                       // previous tile's Bs/cred fully consumed
    load_tile_kmajor(Bs, gu + ((size_t)b * P + j0) * D, nj, tid);
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k * CLD + ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k * CLD + 64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k * CLD + tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k * CLD + 64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * bb[j];
    }
    // row maxima (over gu positions j) accumulate in registers across column tiles
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
        if (jj < nj && acc[i][j] > rmax[i]) { rmax[i] = acc[i][j]; rarg[i] = j0 + jj; }
      }
    // column maxima (over giM rows i) for this tile: thread -> half-warp pair -> warps -> global atomicMax
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unsigned long long key = 0ull;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ii = (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (ii < ni) {
          const unsigned long long k2 = ((unsigned long long)f2ord(acc[i][j]) << 32) | (unsigned)(i0 + ii);
          key = k2 > key ? k2 : key;
        }
      }
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, 16);
      key = o > key ? o : key;
      if (lane < 16) cred[warp * CT + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4)] = key;
    }
    __syncthreads();
    if (tid < nj) {
      unsigned long long key = cred[tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) { const unsigned long long o = cred[w * CT + tid]; key = o > key ? o : key; }
      atomicMax(&colkey[(size_t)b * P + j0 + tid], key);
    }
  }
  // finish the row maxima across the 16 threads that share a row group
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long key = ((unsigned long long)f2ord(rmax[i]) << 32) | (unsigned)rarg[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    const int ii = (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (tx == 0 && ii < ni) rowkey[(size_t)b * P + i0 + ii] = key;
  }
}

// soft = softmax_P(tanh(max)), atte = sum_p soft[p] g[p]  (model.py:52-55)
__global__ void __launch_bounds__(256) coattn_pool_kernel(const unsigned long long* __restrict__ colkey,
                                                          const unsigned long long* __restrict__ rowkey,
                                                          const float* __restrict__ gu, const float* __restrict__ gi, int P,
                                                          float* __restrict__ soft_u, float* __restrict__ soft_i,
                                                          float* __restrict__ t_u, float* __restrict__ t_i,
                                                          int* __restrict__ arg_u, int* __restrict__ arg_i,
                                                          float* __restrict__ atte_u, float* __restrict__ atte_i) {
  extern __shared__ __align__(16) float smem[];
  const int P4 = (P + 3) & ~3;
  float* sp = smem;              // [P4]
  float* red = smem + P4;        // [32]
  float4* part = reinterpret_cast<float4*>(smem + P4 + 32);  // [8][32]
  const int b = blockIdx.x, side = blockIdx.y, tid = threadIdx.x;
  const unsigned long long* key = (side ? rowkey : colkey) + (size_t)b * P;
  const float* g = (side ? gi : gu) + (size_t)b * P * D;
  float* soft = (side ? soft_i : soft_u) + (size_t)b * P;
  float* tv = (side ? t_i : t_u) + (size_t)b * P;
  int* arg = (side ? arg_i : arg_u) + (size_t)b * P;
  float mx = -INFINITY;
  for (int p = tid; p < P; p += 256) {
    const unsigned long long k = key[p];
    const float t = tanhf(ord2f((unsigned)(k >> 32)));
    sp[p] = t; tv[p] = t; arg[p] = (int)(unsigned)(k & 0xffffffffull);
    mx = fmaxf(mx, t);
  }
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int p = tid; p < P; p += 256) { const float e = expf(sp[p] - mx); sp[p] = e; sum += e; }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int p = tid; p < P; p += 256) { const float s = sp[p] * inv; sp[p] = s; soft[p] = s; }
  __syncthreads();
  const int c4 = tid & 31, pg = tid >> 5;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = pg; p < P; p += 8) {
    const float s = sp[p];
    const float4 v = *reinterpret_cast<const float4*>(g + (size_t)p * D + c4 * 4);
    a.x += s * v.x; a.y += s * v.y; a.z += s * v.z; a.w += s * v.w;
  }
  part[pg * 32 + c4] = a;
  __syncthreads();
  if (tid < 32) {
    float4 r = part[tid];
#pragma unroll
    for (int q = 1; q < 8; ++q) { const float4 v = part[q * 32 + tid]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    *reinterpret_cast<float4*>((side ? atte_i : atte_u) + (size_t)b * D + tid * 4) = r;
  }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  return v;
}
// in-place exclusive prefix sum of a[0..n) by one 256-thread block (a[n] receives the total); `tmp` = 32 ints of shared memory
__device__ __forceinline__ void block_excl_scan(int* a, int n, int* tmp) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int per = (n + 255) / 256, lo = tid * per, hi = min(n, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  const int inc = warp_incl_scan(sum, lane);
  if (lane == 31) tmp[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int t = lane < 8 ? tmp[lane] : 0;
    const int ti = warp_incl_scan(t, lane);
    if (lane < 8) tmp[8 + lane] = ti - t;
    if (lane == 7) tmp[16] = ti;
  }
  __syncthreads();
  int run = tmp[8 + w] + inc - sum;
  for (int i = lo; i < hi; ++i) { const int c = a[i]; a[i] = run; run += c; }
  if (tid == 255) a[n] = tmp[16];
  __syncthreads();
}

// backward of model.py:50-55 with respect to gu, gi (through giM = gi·M; the M products are GEMMs outside).
// One CTA per sample.  dA is sparse - one entry per row maximum and one per column maximum - so
//     dgu[j]  = add_u[j] + soft_u[j] d_atte_u + w_j giM[arg_u[j]] + sum over {i : arg_i[i] = j} of v_i giM[i]
//     dgiM[i] = v_i gu[arg_i[i]]                                  + sum over {j : arg_u[j] = i} of w_j gu[j]
//     dgi[i]  = add_i[i] + soft_i[i] d_atte_i
// The sums over the inverse arg-max relations are GATHERS here: the relations are inverted per sample in shared memory (counting
// sort), every output row is written exactly once and no global atomics are issued (the first version scattered them with 128
// atomic adds per entry).  The whole batch is ONE wave of CTAs, so the kernel lasts as long as one CTA's chain of dependent loads:
// the valid rows of each side are compacted into a list first, the arg-max indices are staged in shared memory, and every pass
// over the rows keeps 8 (dot products) or 16 (gathers: 4 rows x [arg-max row, hand-over row, first two inverse entries]) 512-byte
// row loads in flight per warp.
__global__ void __launch_bounds__(256) coattn_bwd_kernel(const float* __restrict__ gu, const float* __restrict__ gi,
                                                         const float* __restrict__ giM, const float* __restrict__ soft_u,
                                                         const float* __restrict__ soft_i, const float* __restrict__ t_u,
                                                         const float* __restrict__ t_i, const int* __restrict__ arg_u,
                                                         const int* __restrict__ arg_i, const float* __restrict__ d_soft_u,
                                                         const float* __restrict__ d_soft_i, const float* __restrict__ d_atte_u,
                                                         const float* __restrict__ d_atte_i, int P, const int* __restrict__ cst_u,
                                                         int S_u, int L_u, const int* __restrict__ cst_i, int S_i, int L_i,
                                                         const float* __restrict__ add_u, const float* __restrict__ add_i,
                                                         float* __restrict__ dgu, float* __restrict__ dgi, float* __restrict__ dgiM) {
  extern __shared__ __align__(16) float smem[];
  const int P4 = (P + 3) & ~3;
  float* wu = smem;            // [P4] d(pre-activation col max)  -> entries (arg_u[j], j)
  float* vi = smem + P4;       // [P4] d(pre-activation row max)  -> entries (i, arg_i[i])
  float* dau = smem + 2 * P4;  // [128]
  float* dai = dau + D;        // [128]
  float* red = dai + D;        // [32]
  int* tmp = reinterpret_cast<int*>(red + 32);          // [36] scan scratch, [32..33] = number of valid rows per side
  int* off_i = tmp + 36;       // [P4 + 4] list of item row i = user positions j with arg_u[j] == i: ent_i[off_i[i] .. off_i[i+1])
  int* off_u = off_i + P4 + 4; // [P4 + 4] list of user row j = item positions i with arg_i[i] == j
  int* cur_i = off_u + P4 + 4; // [P4] fill cursors
  int* cur_u = cur_i + P4;
  int* ent_i = cur_u + P4;     // [P4]
  int* ent_u = ent_i + P4;     // [P4]
  int* au = ent_u + P4;        // [P4] arg_u / arg_i of this sample
  int* ai = au + P4;
  int* lst_u = ai + P4;        // [P4] the valid positions of each side, ascending
  int* lst_i = lst_u + P4;
  // with length tables: rows at or beyond a sentence's length are exactly zero (model.py:20) and are neither read nor written
  // (their gradients are never used); dgiM alone gets explicit zeros there because the dM reduction runs over every row
  unsigned char* mu = reinterpret_cast<unsigned char*>(lst_i + P4);   // [P4] user-side row is valid
  unsigned char* mi = mu + P4;                                         // [P4] item-side row is valid
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t bp = (size_t)b * P;
  if (tid < D) {
    dau[tid] = d_atte_u ? d_atte_u[(size_t)b * D + tid] : 0.f;
    dai[tid] = d_atte_i ? d_atte_i[(size_t)b * D + tid] : 0.f;
  }
  for (int p = tid; p < P; p += 256) {
    unsigned char a = 1, c = 1;
    if (cst_u) {
      const int su = p / L_u, si = p / L_i;
      a = (p - su * L_u) < cst_u[(size_t)b * S_u + su + 1] - cst_u[(size_t)b * S_u + su];
      c = (p - si * L_i) < cst_i[(size_t)b * S_i + si + 1] - cst_i[(size_t)b * S_i + si];
    }
    mu[p] = a; mi[p] = c;
    off_i[p] = 0; off_u[p] = 0;
    au[p] = arg_u[bp + p]; ai[p] = arg_i[bp + p];
    wu[p] = d_soft_u ? d_soft_u[bp + p] : 0.f;        // ds[p] = d_soft[p] (+ <g[p], d_atte> for the valid rows, below)
    vi[p] = d_soft_i ? d_soft_i[bp + p] : 0.f;
  }
  __syncthreads();
  if (warp < 2) {              // compaction of the valid positions: warp 0 the user side, warp 1 the item side
    const unsigned char* ok = warp ? mi : mu;
    int* lst = warp ? lst_i : lst_u;
    int n = 0;
    for (int p0 = 0; p0 < P; p0 += 32) {
      const int p = p0 + lane;
      const bool v = p < P && ok[p];
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) lst[n + __popc(m & ((1u << lane) - 1u))] = p;
      n += __popc(m);
    }
    if (lane == 0) tmp[32 + warp] = n;
  }
  __syncthreads();
  const int nu = tmp[32], ni = tmp[33];
  // ds[p] += <g[p], d_atte> over the valid rows: eight rows per warp and iteration, their loads issued together
  for (int side = 0; side < 2; ++side) {
    const float* g = (side ? gi : gu) + bp * D + lane * 4;
    const float4 d4 = *reinterpret_cast<const float4*>((side ? dai : dau) + lane * 4);
    float* dst = side ? vi : wu;
    const int* lst = side ? lst_i : lst_u;
    const int n = side ? ni : nu;
    for (int i0 = warp; i0 < n; i0 += 64) {
      float4 v[8];
      int pp[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = i0 + 8 * q;
        pp[q] = idx < n ? lst[idx] : -1;
        v[q] = pp[q] >= 0 ? *reinterpret_cast<const float4*>(g + (size_t)pp[q] * D) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float s = v[q].x * d4.x + v[q].y * d4.y + v[q].z * d4.z + v[q].w * d4.w;
        s = warp_sum(s);
        if (lane == 0 && pp[q] >= 0) dst[pp[q]] += s;
      }
    }
  }
  __syncthreads();
  // softmax backward, then through tanh
  for (int side = 0; side < 2; ++side) {
    const float* so = (side ? soft_i : soft_u) + bp;
    const float* tv = (side ? t_i : t_u) + bp;
    float* dst = side ? vi : wu;
    float dot = 0.f;
    for (int p = tid; p < P; p += 256) dot += so[p] * dst[p];
    dot = block_sum(dot, red);
    for (int p = tid; p < P; p += 256) {
      const float t = tv[p];
      dst[p] = so[p] * (dst[p] - dot) * (1.f - t * t);
    }
    __syncthreads();
  }
  // invert the two arg-max relations (entries with a zero weight or an invalid end carry nothing and are left out)
  for (int p = tid; p < P; p += 256) {
    if (wu[p] != 0.f && mu[p]) { const int a = au[p]; if (mi[a]) atomicAdd(&off_i[a], 1); }
    if (vi[p] != 0.f && mi[p]) { const int a = ai[p]; if (mu[a]) atomicAdd(&off_u[a], 1); }
  }
  __syncthreads();
  block_excl_scan(off_i, P, tmp);
  block_excl_scan(off_u, P, tmp);
  for (int p = tid; p < P; p += 256) { cur_i[p] = off_i[p]; cur_u[p] = off_u[p]; }
  __syncthreads();
  for (int p = tid; p < P; p += 256) {
    if (wu[p] != 0.f && mu[p]) { const int a = au[p]; if (mi[a]) ent_i[atomicAdd(&cur_i[a], 1)] = p; }
    if (vi[p] != 0.f && mi[p]) { const int a = ai[p]; if (mu[a]) ent_u[atomicAdd(&cur_u[a], 1)] = p; }
  }
  __syncthreads();
  const float4 dau4 = *reinterpret_cast<const float4*>(dau + lane * 4);
  const float4 dai4 = *reinterpret_cast<const float4*>(dai + lane * 4);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto row = [&](const float* base, int p) { return *reinterpret_cast<const float4*>(base + (bp + p) * D + lane * 4); };
  auto fma4 = [](float4& acc, float s, const float4 v) { acc.x += s * v.x; acc.y += s * v.y; acc.z += s * v.z; acc.w += s * v.w; };
  // user rows: four per warp and iteration
  for (int i0 = warp; i0 < nu; i0 += 32) {
    float4 m[4], eu[4], r0[4], r1[4];
    float sf[4];
    int pp[4], e0[4], e1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = i0 + 8 * q;
      const int p = idx < nu ? lst_u[idx] : -1;
      pp[q] = p;
      m[q] = eu[q] = r0[q] = r1[q] = zero4;
      sf[q] = 0.f; e0[q] = e1[q] = 0;
      if (p >= 0) {
        const int a = au[p];
        if (wu[p] != 0.f && mi[a]) m[q] = row(giM, a);
        if (add_u) eu[q] = row(add_u, p);
        sf[q] = soft_u[bp + p];
        e0[q] = off_u[p]; e1[q] = off_u[p + 1];
        if (e0[q] < e1[q]) r0[q] = row(giM, ent_u[e0[q]]);
        if (e0[q] + 1 < e1[q]) r1[q] = row(giM, ent_u[e0[q] + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = pp[q];
      if (p < 0) continue;
      float4 acc = eu[q];
      fma4(acc, sf[q], dau4);
      fma4(acc, wu[p], m[q]);
      if (e0[q] < e1[q]) fma4(acc, vi[ent_u[e0[q]]], r0[q]);          // item rows whose maximum sits in this user position
      if (e0[q] + 1 < e1[q]) fma4(acc, vi[ent_u[e0[q] + 1]], r1[q]);
      for (int e = e0[q] + 2; e < e1[q]; e += 2) {
        const int x0 = ent_u[e], x1 = e + 1 < e1[q] ? ent_u[e + 1] : -1;
        const float4 a0 = row(giM, x0), a1 = x1 >= 0 ? row(giM, x1) : zero4;
        fma4(acc, vi[x0], a0);
        if (x1 >= 0) fma4(acc, vi[x1], a1);
      }
      *reinterpret_cast<float4*>(dgu + (bp + p) * D + lane * 4) = acc;
    }
  }
  // item rows
  for (int i0 = warp; i0 < ni; i0 += 32) {
    float4 u[4], ei[4], r0[4], r1[4];
    float sf[4];
    int pp[4], e0[4], e1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = i0 + 8 * q;
      const int p = idx < ni ? lst_i[idx] : -1;
      pp[q] = p;
      u[q] = ei[q] = r0[q] = r1[q] = zero4;
      sf[q] = 0.f; e0[q] = e1[q] = 0;
      if (p >= 0) {
        const int a = ai[p];
        if (vi[p] != 0.f && mu[a]) u[q] = row(gu, a);
        if (add_i) ei[q] = row(add_i, p);
        sf[q] = soft_i[bp + p];
        e0[q] = off_i[p]; e1[q] = off_i[p + 1];
        if (e0[q] < e1[q]) r0[q] = row(gu, ent_i[e0[q]]);
        if (e0[q] + 1 < e1[q]) r1[q] = row(gu, ent_i[e0[q] + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = pp[q];
      if (p < 0) continue;
      float4 acc = zero4;
      fma4(acc, vi[p], u[q]);
      if (e0[q] < e1[q]) fma4(acc, wu[ent_i[e0[q]]], r0[q]);          // user positions whose maximum sits in this item row
      if (e0[q] + 1 < e1[q]) fma4(acc, wu[ent_i[e0[q] + 1]], r1[q]);
      for (int e = e0[q] + 2; e < e1[q]; e += 2) {
        const int x0 = ent_i[e], x1 = e + 1 < e1[q] ? ent_i[e + 1] : -1;
        const float4 a0 = row(gu, x0), a1 = x1 >= 0 ? row(gu, x1) : zero4;
        fma4(acc, wu[x0], a0);
        if (x1 >= 0) fma4(acc, wu[x1], a1);
      }
      *reinterpret_cast<float4*>(dgiM + (bp + p) * D + lane * 4) = acc;
      float4 g2 = ei[q];
      fma4(g2, sf[q], dai4);
      *reinterpret_cast<float4*>(dgi + (bp + p) * D + lane * 4) = g2;
    }
  }
  if (ni < P)                  // rows beyond the sentences' lengths: dgiM = 0 (the dM reduction reads every row)
    for (int p = warp; p < P; p += 8)
      if (!mi[p]) *reinterpret_cast<float4*>(dgiM + (bp + p) * D + lane * 4) = zero4;
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_coattn_fwd(const float* gu, const float* gi, const float* giM, int B, int P, unsigned long long* rowkey,
                               unsigned long long* colkey /* zero-initialised */, float* soft_u, float* soft_i, float* t_u,
                               float* t_i, int32_t* arg_u, int32_t* arg_i, float* atte_u, float* atte_i, void* stream) {
  if (B <= 0 || P <= 0) return 0;
  if (B > 65535) return fail_arg("coattn_fwd: batch %d > 65535", B);
  const size_t sm1 = sizeof(float) * 2 * D * CLD + sizeof(unsigned long long) * 8 * CT;
  {
    cudaError_t e = cudaFuncSetAttribute(coattn_affinity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    if (e != cudaSuccess) { set_error("coattn smem: %s", cudaGetErrorString(e)); return (int)e; }
  }
  coattn_affinity_kernel<<<dim3((P + CT - 1) / CT, B), 256, sm1, (cudaStream_t)stream>>>(giM, gu, P, rowkey, colkey);
  if (int e = check_launch("coattn_affinity")) return e;
  const size_t sm2 = sizeof(float) * (((P + 3) & ~3) + 32) + sizeof(float4) * 8 * 32;
  if (sm2 > 200 * 1024) return fail_arg("coattn_fwd: P=%d too large", P);
  if (sm2 > 48 * 1024) cudaFuncSetAttribute(coattn_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  coattn_pool_kernel<<<dim3(B, 2), 256, sm2, (cudaStream_t)stream>>>(colkey, rowkey, gu, gi, P, soft_u, soft_i, t_u, t_i, arg_u,
                                                                    arg_i, atte_u, atte_i);
  return check_launch("coattn_pool");
}

extern "C" int umpr_coattn_bwd(const float* gu, const float* gi, const float* giM, const float* soft_u, const float* soft_i,
                               const float* t_u, const float* t_i, const int32_t* arg_u, const int32_t* arg_i,
                               const float* d_soft_u, const float* d_soft_i, const float* d_atte_u, const float* d_atte_i, int B,
                               int P, const int32_t* cst_u, int S_u, int L_u, const int32_t* cst_i, int S_i, int L_i,
                               const float* add_u, const float* add_i, float* dgu, float* dgi, float* dgiM, void* stream) {
  if (B <= 0 || P <= 0) return 0;
  if ((cst_u == nullptr) != (cst_i == nullptr)) return fail_arg("coattn_bwd: length tables must be given for both sides or neither");
  if (cst_u && (S_u * L_u != P || S_i * L_i != P || S_u < 1 || S_i < 1))
    return fail_arg("coattn_bwd: S*L must equal P=%d on both sides (got %d*%d, %d*%d)", P, S_u, L_u, S_i, L_i);
  const size_t P4 = (size_t)((P + 3) & ~3);
  const size_t sm = sizeof(float) * (2 * P4 + 2 * D + 32) + sizeof(int) * (36 + 2 * (P4 + 4) + 8 * P4) + 2 * P4;
  if (sm > 200 * 1024) return fail_arg("coattn_bwd: P=%d too large", P);
  if (sm > 48 * 1024) cudaFuncSetAttribute(coattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  coattn_bwd_kernel<<<B, 256, sm, (cudaStream_t)stream>>>(gu, gi, giM, soft_u, soft_i, t_u, t_i, arg_u, arg_i, d_soft_u, d_soft_i,
                                                         d_atte_u, d_atte_i, P, cst_u, S_u, L_u, cst_i, S_i, L_i, add_u, add_i, dgu, dgi, dgiM);
  return check_launch("coattn_bwd");
}
