// ImprovedRnn (reference src/model.py:12-21): varlen bidirectional GRU with the reference's
// pack / total_length / double-un-sort semantics, forward and backward.
//
// Data layout in HBM:
//   xp  [n_slabs][R][64]          packed inputs: E embedding floats, 1.0 (bias column), zeros
//   G   [n_slabs][2][R][192]      input projections W_ih x + b_ih (+ b_hh for r,z) per direction
//   sv  [n_slabs][2][R][256]      saved r, z, n, (W_hn h + b_hn) for the backward pass
//   dG  [n_slabs][2][R][256]      gate-pre-activation gradients dr, dz, dn, dn*r
//   out [N][L][128]               ImprovedRnn result, rows in the reference's (double-permuted) order
// A "slab" is one time step of one tile of R length-sorted jobs; tiles early-exit at their own max length.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

// ------------------------------------------------------------------------------------------------
// 1. warp-vectorised embedding gather + length-aware pack
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_pack_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids,
                                                          const float* __restrict__ dense, Plan p, int L, int E,
                                                          float* __restrict__ xp) {
  const int sl = blockIdx.x;
  const int j = p.slab_tile[sl];
  const int t = sl - p.tile_off[j];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R;
  const bool vec2 = (E & 1) == 0;
  for (int r = warp; r < R; r += 8) {
    const int k = j * R + r;
    float2 v = make_float2(0.f, 0.f);
    if (t < p.len_of[k]) {
      const size_t tok = (size_t)p.seq_of[k] * L + t;
      const float* src = dense ? dense + tok * E : table + (size_t)ids[tok] * E;
      const int e0 = 2 * lane;
      if (vec2 && e0 + 1 < E) {
        v = *reinterpret_cast<const float2*>(src + e0);
      } else {
        if (e0 < E) v.x = src[e0]; else if (e0 == E) v.x = 1.f;
        if (e0 + 1 < E) v.y = src[e0 + 1]; else if (e0 + 1 == E) v.y = 1.f;
      }
    }
    reinterpret_cast<float2*>(xp + ((size_t)sl * R + r) * KP)[lane] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// 2. input projection for every packed token: G = xp · [W_ih | b]^T   (CUDA-core fp32 path)
// ------------------------------------------------------------------------------------------------
constexpr int AS_LD = KP + 4;    // 68 floats: keeps float4 alignment, breaks the row/bank aliasing
constexpr int WS_LD = G3 + 4;    // 196

template <int RPT>
__global__ void __launch_bounds__(256) gru_inproj_kernel(const float* __restrict__ xp, const float* __restrict__ w_ih_f,
                                                         const float* __restrict__ b_ih_f, const float* __restrict__ b_hh_f,
                                                         const float* __restrict__ w_ih_b, const float* __restrict__ b_ih_b,
                                                         const float* __restrict__ b_hh_b, int E, float* __restrict__ G) {
  constexpr int R = 16 * RPT;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // [R][68]   row-major, k contiguous
  float* Bs = smem + R * AS_LD;     // [64][196] k-major
  const int sl = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* w_ih = dir ? w_ih_b : w_ih_f;
  const float* b_ih = dir ? b_ih_b : b_ih_f;
  const float* b_hh = dir ? b_hh_b : b_hh_f;
  const int K4 = (E + 1 + 3) / 4;
  const float4* src = reinterpret_cast<const float4*>(xp + (size_t)sl * R * KP);
  for (int idx = tid; idx < R * 16; idx += 256) {
    const int r = idx >> 4, c4 = idx & 15;
    cp_async16(&As[r * AS_LD + c4 * 4], src + idx);
  }
  cp_async_commit();
  for (int idx = tid; idx < G3 * E; idx += 256) {
    const int g = idx / E, e = idx - g * E;
    Bs[e * WS_LD + g] = w_ih[idx];
  }
  for (int idx = tid; idx < G3 * (K4 * 4 - E); idx += 256) {
    const int e = E + idx / G3, g = idx % G3;
    Bs[e * WS_LD + g] = (e == E) ? (b_ih[g] + (g < 2 * H ? b_hh[g] : 0.f)) : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  float acc[RPT][12];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[i][c] = 0.f;
  for (int k4 = 0; k4 < K4; ++k4) {
    float4 a[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(&As[(ty * RPT + i) * AS_LD + k4 * 4]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* brow = &Bs[(k4 * 4 + kk) * WS_LD + tx * 4];
      const float4 b0 = *reinterpret_cast<const float4*>(brow);
      const float4 b1 = *reinterpret_cast<const float4*>(brow + H);
      const float4 b2 = *reinterpret_cast<const float4*>(brow + 2 * H);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] += av * b0.x; acc[i][1] += av * b0.y; acc[i][2] += av * b0.z; acc[i][3] += av * b0.w;
        acc[i][4] += av * b1.x; acc[i][5] += av * b1.y; acc[i][6] += av * b1.z; acc[i][7] += av * b1.w;
        acc[i][8] += av * b2.x; acc[i][9] += av * b2.y; acc[i][10] += av * b2.z; acc[i][11] += av * b2.w;
      }
    }
  }
  float* dst = G + ((size_t)sl * 2 + dir) * R * G3;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    float* row = dst + (size_t)(ty * RPT + i) * G3 + tx * 4;
    *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(row + H) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    *reinterpret_cast<float4*>(row + 2 * H) = make_float4(acc[i][8], acc[i][9], acc[i][10], acc[i][11]);
  }
}

// ------------------------------------------------------------------------------------------------
// 3. recurrence, forward.  One CTA per (tile of R jobs, direction); W_hh stays in shared memory for
//    the whole sequence; every job is masked by its own length; outputs land in the reference's row order.
// ------------------------------------------------------------------------------------------------
template <int RPT>
__global__ void __launch_bounds__(256, 1) gru_rec_fwd_kernel(const float* __restrict__ G, const float* __restrict__ w_hh_f,
                                                             const float* __restrict__ b_hh_f, const float* __restrict__ w_hh_b,
                                                             const float* __restrict__ b_hh_b, Plan p, int L,
                                                             float* __restrict__ out, float* __restrict__ hn,
                                                             float* __restrict__ sv, int N) {
  constexpr int R = 16 * RPT;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                  // [64][196]  Ws[k][g] = W_hh[g][k]
  float* hT = Ws + H * WS_LD;        // [64][R]    hT[k][job]
  float* Gs = hT + H * R;            // [R][192]   this step's input projections
  const int j = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* w_hh = dir ? w_hh_b : w_hh_f;
  const float* b_hh = dir ? b_hh_b : b_hh_f;
  for (int idx = tid; idx < G3 * H; idx += 256) {
    const int g = idx >> 6, k = idx & 63;
    Ws[k * WS_LD + g] = w_hh[idx];
  }
  for (int idx = tid; idx < H * R; idx += 256) hT[idx] = 0.f;

  int len[RPT], row[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int k = j * R + ty * RPT + i;
    len[i] = p.len_of[k];
    row[i] = p.row_of[k];
  }
  const int Lj = p.len_of[j * R];
  const int slab0 = p.tile_off[j];
  float bhn[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) bhn[c] = b_hh[2 * H + tx * 4 + c];
  float hreg[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) hreg[i][c] = 0.f;

  auto prefetch = [&](int s) {
    const int t = dir ? (Lj - 1 - s) : s;
    const float4* src = reinterpret_cast<const float4*>(G + ((size_t)(slab0 + t) * 2 + dir) * R * G3);
    for (int idx = tid; idx < R * (G3 / 4); idx += 256) cp_async16(&Gs[idx * 4], src + idx);
    cp_async_commit();
  };
  if (Lj > 0) prefetch(0);
  __syncthreads();

  for (int s = 0; s < Lj; ++s) {
    const int t = dir ? (Lj - 1 - s) : s;
    float acc[RPT][12];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int c = 0; c < 12; ++c) acc[i][c] = 0.f;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      float a[RPT];
      if constexpr (RPT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < RPT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&hT[k * R + ty * RPT + q * 4]);
          a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = hT[k * R + ty * RPT + i];
      }
      const float* brow = &Ws[k * WS_LD + tx * 4];
      const float4 b0 = *reinterpret_cast<const float4*>(brow);
      const float4 b1 = *reinterpret_cast<const float4*>(brow + H);
      const float4 b2 = *reinterpret_cast<const float4*>(brow + 2 * H);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        acc[i][0] += a[i] * b0.x; acc[i][1] += a[i] * b0.y; acc[i][2] += a[i] * b0.z; acc[i][3] += a[i] * b0.w;
        acc[i][4] += a[i] * b1.x; acc[i][5] += a[i] * b1.y; acc[i][6] += a[i] * b1.z; acc[i][7] += a[i] * b1.w;
        acc[i][8] += a[i] * b2.x; acc[i][9] += a[i] * b2.y; acc[i][10] += a[i] * b2.z; acc[i][11] += a[i] * b2.w;
      }
    }
    cp_async_wait_all();
    __syncthreads();   // G(t) landed; every thread finished reading hT
    float* svb = sv ? sv + ((size_t)(slab0 + t) * 2 + dir) * R * SV : nullptr;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = ty * RPT + i;
      if (row[i] < 0) continue;
      float* orow = out + ((size_t)row[i] * L + t) * D + dir * H + tx * 4;
      if (t < len[i]) {
        const float4 gr = *reinterpret_cast<const float4*>(&Gs[r * G3 + tx * 4]);
        const float4 gz = *reinterpret_cast<const float4*>(&Gs[r * G3 + H + tx * 4]);
        const float4 gn = *reinterpret_cast<const float4*>(&Gs[r * G3 + 2 * H + tx * 4]);
        const float grr[4] = {gr.x, gr.y, gr.z, gr.w}, gzz[4] = {gz.x, gz.y, gz.z, gz.w}, gnn[4] = {gn.x, gn.y, gn.z, gn.w};
        float rr[4], zz[4], nn[4], hh[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          rr[c] = sigmoidf_acc(grr[c] + acc[i][c]);
          zz[c] = sigmoidf_acc(gzz[c] + acc[i][4 + c]);
          hh[c] = acc[i][8 + c] + bhn[c];
          nn[c] = tanhf(gnn[c] + rr[c] * hh[c]);
          hreg[i][c] = (hreg[i][c] - nn[c]) * zz[c] + nn[c];    // ATen GRU cell: (h - n) * z + n
          hT[(tx * 4 + c) * R + r] = hreg[i][c];
        }
        *reinterpret_cast<float4*>(orow) = make_float4(hreg[i][0], hreg[i][1], hreg[i][2], hreg[i][3]);
        if (svb) {
          float* s4 = svb + (size_t)r * SV + tx * 4;
          *reinterpret_cast<float4*>(s4) = make_float4(rr[0], rr[1], rr[2], rr[3]);
          *reinterpret_cast<float4*>(s4 + H) = make_float4(zz[0], zz[1], zz[2], zz[3]);
          *reinterpret_cast<float4*>(s4 + 2 * H) = make_float4(nn[0], nn[1], nn[2], nn[3]);
          *reinterpret_cast<float4*>(s4 + 3 * H) = make_float4(hh[0], hh[1], hh[2], hh[3]);
        }
      } else {
        *reinterpret_cast<float4*>(orow) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();   // new hT visible; Gs consumed
    if (s + 1 < Lj) prefetch(s + 1);
  }
  // zero padding up to total_length (model.py:20) and the final hidden state in ORIGINAL sequence order
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    if (row[i] < 0) continue;
    for (int t = Lj; t < L; ++t)
      *reinterpret_cast<float4*>(out + ((size_t)row[i] * L + t) * D + dir * H + tx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hn) {
      const int seq = p.seq_of[j * R + ty * RPT + i];
      *reinterpret_cast<float4*>(hn + ((size_t)dir * N + seq) * H + tx * 4) =
          make_float4(hreg[i][0], hreg[i][1], hreg[i][2], hreg[i][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 4. recurrence, backward (reverse time).  Writes dG = [dr, dz, dn, dn*r] (pre-activation grads) per token.
// ------------------------------------------------------------------------------------------------
template <int RPT>
__global__ void __launch_bounds__(256, 1) gru_rec_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ d_hn,
                                                             const float* __restrict__ out, const float* __restrict__ sv,
                                                             const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_b,
                                                             Plan p, int L, float* __restrict__ dG, int N) {
  constexpr int R = 16 * RPT;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;               // [192][64]  natural W_hh layout: Ws[g][c]
  float* dT = Ws + G3 * H;        // [192][R]   dT[g][job] = d(hidden-side gate pre-activation)
  const int j = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* w_hh = dir ? w_hh_b : w_hh_f;
  for (int idx = tid; idx < G3 * H; idx += 256) Ws[idx] = w_hh[idx];

  int len[RPT], row[RPT];
  float carry[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int k = j * R + ty * RPT + i;
    len[i] = p.len_of[k];
    row[i] = p.row_of[k];
#pragma unroll
    for (int c = 0; c < 4; ++c) carry[i][c] = 0.f;
    if (d_hn && row[i] >= 0) {
      const float4 v = *reinterpret_cast<const float4*>(d_hn + ((size_t)dir * N + p.seq_of[k]) * H + tx * 4);
      carry[i][0] = v.x; carry[i][1] = v.y; carry[i][2] = v.z; carry[i][3] = v.w;
    }
  }
  const int Lj = p.len_of[j * R];
  const int slab0 = p.tile_off[j];
  __syncthreads();

  for (int s = 0; s < Lj; ++s) {
    const int t = dir ? s : (Lj - 1 - s);     // reverse of the forward kernel's order
    const size_t sbase = ((size_t)(slab0 + t) * 2 + dir) * R;
    float part[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = ty * RPT + i;
      float dr_[4] = {0, 0, 0, 0}, dz_[4] = {0, 0, 0, 0}, dn_[4] = {0, 0, 0, 0}, dnr_[4] = {0, 0, 0, 0};
      if (row[i] >= 0 && t < len[i]) {
        const float* s4 = sv + (sbase + r) * SV + tx * 4;
        const float4 r4 = *reinterpret_cast<const float4*>(s4);
        const float4 z4 = *reinterpret_cast<const float4*>(s4 + H);
        const float4 n4 = *reinterpret_cast<const float4*>(s4 + 2 * H);
        const float4 h4 = *reinterpret_cast<const float4*>(s4 + 3 * H);
        const float4 dy = *reinterpret_cast<const float4*>(d_out + ((size_t)row[i] * L + t) * D + dir * H + tx * 4);
        float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
        const int tp = dir ? t + 1 : t - 1;
        if (tp >= 0 && tp < len[i]) hp = *reinterpret_cast<const float4*>(out + ((size_t)row[i] * L + tp) * D + dir * H + tx * 4);
        const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w}, nn[4] = {n4.x, n4.y, n4.z, n4.w};
        const float hh[4] = {h4.x, h4.y, h4.z, h4.w}, dyy[4] = {dy.x, dy.y, dy.z, dy.w}, hpp[4] = {hp.x, hp.y, hp.z, hp.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float dh = carry[i][c] + dyy[c];
          const float dn = dh * (1.f - zz[c]);
          const float dz = dh * (hpp[c] - nn[c]);
          dn_[c] = dn * (1.f - nn[c] * nn[c]);
          dz_[c] = dz * zz[c] * (1.f - zz[c]);
          dr_[c] = dn_[c] * hh[c] * rr[c] * (1.f - rr[c]);
          dnr_[c] = dn_[c] * rr[c];
          part[i][c] = dh * zz[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) part[i][c] = carry[i][c];
      }
      float* g4 = dG + (sbase + r) * SV + tx * 4;
      *reinterpret_cast<float4*>(g4) = make_float4(dr_[0], dr_[1], dr_[2], dr_[3]);
      *reinterpret_cast<float4*>(g4 + H) = make_float4(dz_[0], dz_[1], dz_[2], dz_[3]);
      *reinterpret_cast<float4*>(g4 + 2 * H) = make_float4(dn_[0], dn_[1], dn_[2], dn_[3]);
      *reinterpret_cast<float4*>(g4 + 3 * H) = make_float4(dnr_[0], dnr_[1], dnr_[2], dnr_[3]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        dT[(tx * 4 + c) * R + r] = dr_[c];
        dT[(H + tx * 4 + c) * R + r] = dz_[c];
        dT[(2 * H + tx * 4 + c) * R + r] = dnr_[c];
      }
    }
    __syncthreads();
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll 4
    for (int g = 0; g < G3; ++g) {
      float a[RPT];
      if constexpr (RPT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < RPT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&dT[g * R + ty * RPT + q * 4]);
          a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = dT[g * R + ty * RPT + i];
      }
      const float4 b = *reinterpret_cast<const float4*>(&Ws[g * H + tx * 4]);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        acc[i][0] += a[i] * b.x; acc[i][1] += a[i] * b.y; acc[i][2] += a[i] * b.z; acc[i][3] += a[i] * b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) carry[i][c] = part[i][c] + acc[i][c];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// 5. weight gradients: C[256][128] = dG^T · [xp | h_prev], reduced over all packed tokens (split-K + atomics)
//    rows 0..191 x cols 0..E-1  -> dW_ih ; col E (the 1.0 column) -> db_ih and the r,z part of db_hh
//    rows 0..127 and 192..255 x cols 64..127 -> dW_hh ; row 192.., col E -> the n part of db_hh
// ------------------------------------------------------------------------------------------------
constexpr int WG_KK = 32;
__global__ void __launch_bounds__(256, 1) gru_wgrad_kernel(const float* __restrict__ dG, const float* __restrict__ xp,
                                                           const float* __restrict__ out, Plan p, int L, int E, int slabs_per_cta,
                                                           float* __restrict__ dw_ih_f, float* __restrict__ dw_hh_f,
                                                           float* __restrict__ db_ih_f, float* __restrict__ db_hh_f,
                                                           float* __restrict__ dw_ih_b, float* __restrict__ dw_hh_b,
                                                           float* __restrict__ db_ih_b, float* __restrict__ db_hh_b) {
  __shared__ __align__(16) float As[WG_KK][SV];       // dG rows
  __shared__ __align__(16) float Bs[WG_KK][2 * KP];   // [xp | h_prev]
  const int dir = blockIdx.y, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int R = p.R;
  const int sl0 = blockIdx.x * slabs_per_cta;
  const int sl1 = min(p.n_slabs, sl0 + slabs_per_cta);
  float acc[16][8];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;

  for (int sl = sl0; sl < sl1; ++sl) {
    const int j = p.slab_tile[sl];
    const int t = sl - p.tile_off[j];
    for (int r0 = 0; r0 < R; r0 += WG_KK) {
      // dG rows (contiguous), xp rows (contiguous), h_prev rows (gathered from `out`)
      const float4* ga = reinterpret_cast<const float4*>(dG + (((size_t)sl * 2 + dir) * R + r0) * SV);
      for (int idx = tid; idx < WG_KK * (SV / 4); idx += 256) cp_async16(&As[0][0] + idx * 4, ga + idx);
      const float4* xa = reinterpret_cast<const float4*>(xp + ((size_t)sl * R + r0) * KP);
      for (int idx = tid; idx < WG_KK * (KP / 4); idx += 256) {
        const int r = idx >> 4, c4 = idx & 15;
        cp_async16(&Bs[r][c4 * 4], xa + idx);
      }
      cp_async_commit();
      for (int idx = tid; idx < WG_KK * (H / 4); idx += 256) {
        const int r = idx >> 4, c4 = idx & 15;
        const int k = j * R + r0 + r;
        const int rw = p.row_of[k];
        const int tp = dir ? t + 1 : t - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rw >= 0 && tp >= 0 && tp < p.len_of[k])
          v = *reinterpret_cast<const float4*>(out + ((size_t)rw * L + tp) * D + dir * H + c4 * 4);
        *reinterpret_cast<float4*>(&Bs[r][KP + c4 * 4]) = v;
      }
      cp_async_wait_all();
      __syncthreads();
#pragma unroll 2
      for (int kk = 0; kk < WG_KK; ++kk) {
        float a[16], b[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * 16 + q * 4]);
          a[q * 4] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + q * 4]);
          b[q * 4] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] += a[i] * b[c];
      }
      __syncthreads();
    }
  }
  float* dw_ih = dir ? dw_ih_b : dw_ih_f;
  float* dw_hh = dir ? dw_hh_b : dw_hh_f;
  float* db_ih = dir ? db_ih_b : db_ih_f;
  float* db_hh = dir ? db_hh_b : db_hh_f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int g = ty * 16 + i;            // 0..255 : dr, dz, dn, dn*r
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = tx * 8 + c;         // 0..63 x side, 64..127 h side
      const float v = acc[i][c];
      if (col < KP) {
        if (g < G3) {
          if (col < E) atomicAdd(&dw_ih[g * E + col], v);
          else if (col == E) { atomicAdd(&db_ih[g], v); if (g < 2 * H) atomicAdd(&db_hh[g], v); }
        } else if (col == E) {
          atomicAdd(&db_hh[g - H], v);    // rows 192..255 -> b_hn at index 128..191
        }
      } else {
        const int hc = col - KP;
        if (g < 2 * H) atomicAdd(&dw_hh[g * H + hc], v);
        else if (g >= G3) atomicAdd(&dw_hh[(g - H) * H + hc], v);
      }
    }
  }
}

template <int RPT> static size_t inproj_smem() { return sizeof(float) * (16 * RPT * AS_LD + KP * WS_LD); }
template <int RPT> static size_t recf_smem() { return sizeof(float) * (H * WS_LD + H * 16 * RPT + 16 * RPT * G3); }
template <int RPT> static size_t recb_smem() { return sizeof(float) * (G3 * H + G3 * 16 * RPT); }

template <typename K> static int set_smem(K kern, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // namespace umpr

using namespace umpr;

#define DISPATCH_R(R, ...)                                           \
  switch (R) {                                                       \
    case 32: { constexpr int RPT = 2; __VA_ARGS__; } break;          \
    case 64: { constexpr int RPT = 4; __VA_ARGS__; } break;          \
    case 128: { constexpr int RPT = 8; __VA_ARGS__; } break;         \
    default: return fail_arg("tile rows R=%d not in {32,64,128}", R); \
  }

extern "C" int umpr_gather_pack(const float* table, const int64_t* ids, const float* dense, const int32_t* plan,
                                int n_tiles, int n_slabs, int R, int L, int E, float* xp, void* stream) {
  if ((!table || !ids) && !dense) return fail_arg("gather_pack: need (table, ids) or dense");
  if (E < 1 || E >= KP) return fail_arg("gather_pack: embedding width E=%d must be in [1, %d)", E, KP);
  if (R != 32 && R != 64 && R != 128) return fail_arg("gather_pack: R=%d", R);
  if (n_slabs == 0) return 0;
  Plan p = make_plan(plan, n_tiles, n_slabs, R);
  gather_pack_kernel<<<n_slabs, 256, 0, (cudaStream_t)stream>>>(table, ids, dense, p, L, E, xp);
  return check_launch("gather_pack");
}

extern "C" int umpr_gru_inproj(const float* xp, const float* const* w /*8 ptrs, nn.GRU order*/, int n_slabs, int R, int E,
                               float* G, void* stream) {
  if (E < 1 || E >= KP) return fail_arg("gru_inproj: E=%d", E);
  if (n_slabs == 0) return 0;
  DISPATCH_R(R, {
    const size_t sm = inproj_smem<RPT>();
    if (int e = set_smem(gru_inproj_kernel<RPT>, sm)) return e;
    gru_inproj_kernel<RPT><<<dim3(n_slabs, 2), 256, sm, (cudaStream_t)stream>>>(xp, w[0], w[2], w[3], w[4], w[6], w[7], E, G);
  });
  return check_launch("gru_inproj");
}

extern "C" int umpr_gru_recurrence_fwd(const float* G, const float* const* w, const int32_t* plan, int n_tiles, int n_slabs,
                                       int R, int N, int L, float* out, float* hn, float* sv, void* stream) {
  if (n_tiles == 0) return 0;
  Plan p = make_plan(plan, n_tiles, n_slabs, R);
  DISPATCH_R(R, {
    const size_t sm = recf_smem<RPT>();
    if (int e = set_smem(gru_rec_fwd_kernel<RPT>, sm)) return e;
    gru_rec_fwd_kernel<RPT><<<dim3(n_tiles, 2), 256, sm, (cudaStream_t)stream>>>(G, w[1], w[3], w[5], w[7], p, L, out, hn, sv, N);
  });
  return check_launch("gru_recurrence_fwd");
}

extern "C" int umpr_gru_recurrence_bwd(const float* d_out, const float* d_hn, const float* out, const float* sv,
                                       const float* const* w, const int32_t* plan, int n_tiles, int n_slabs, int R, int N,
                                       int L, float* dG, void* stream) {
  if (n_tiles == 0) return 0;
  Plan p = make_plan(plan, n_tiles, n_slabs, R);
  DISPATCH_R(R, {
    const size_t sm = recb_smem<RPT>();
    if (int e = set_smem(gru_rec_bwd_kernel<RPT>, sm)) return e;
    gru_rec_bwd_kernel<RPT><<<dim3(n_tiles, 2), 256, sm, (cudaStream_t)stream>>>(d_out, d_hn, out, sv, w[1], w[5], p, L, dG, N);
  });
  return check_launch("gru_recurrence_bwd");
}

extern "C" int umpr_gru_wgrad(const float* dG, const float* xp, const float* out, const int32_t* plan, int n_tiles,
                              int n_slabs, int R, int L, int E, float* const* dw /*8 ptrs, nn.GRU order, pre-zeroed or accumulating*/,
                              int n_ctas, void* stream) {
  if (n_slabs == 0) return 0;
  if (R % WG_KK) return fail_arg("gru_wgrad: R=%d", R);
  Plan p = make_plan(plan, n_tiles, n_slabs, R);
  if (n_ctas < 1) n_ctas = 1;
  const int per = (n_slabs + n_ctas - 1) / n_ctas;
  const int grid = (n_slabs + per - 1) / per;
  gru_wgrad_kernel<<<dim3(grid, 2), 256, 0, (cudaStream_t)stream>>>(dG, xp, out, p, L, E, per, dw[0], dw[1], dw[2], dw[3], dw[4],
                                                                  dw[5], dw[6], dw[7]);
  return check_launch("gru_wgrad");
}
