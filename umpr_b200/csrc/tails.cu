// Small per-sample stages around the big kernels (all HBM/launch bound, one pass over their inputs):
//   fusion Linear+ReLU (model.py:242-245,252-255,268,274), losses (model.py:269,275-277),
//   ControlNet tail: SSNet + Eq.18 + quadratic gates (model.py:142-143,185-197),
//   VisualNet tail (model.py:219-228), and the tanh backward of the text-matching layer (model.py:168).
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

// ---------------------------------------------------------------- fusion: pred = relu(w · [repr, fpos, fneg] + b)
__global__ void __launch_bounds__(128) fusion_fwd_kernel(const float* __restrict__ repr, const float* __restrict__ fpos,
                                                         const float* __restrict__ fneg, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int V, float* __restrict__ pred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  const float4 r = *reinterpret_cast<const float4*>(repr + (size_t)b * D + lane * 4);
  const float4 ww = *reinterpret_cast<const float4*>(w + lane * 4);
  float a = r.x * ww.x + r.y * ww.y + r.z * ww.z + r.w * ww.w;
  for (int v = lane; v < V; v += 32) a += fpos[(size_t)b * V + v] * w[D + v] + fneg[(size_t)b * V + v] * w[D + V + v];
  a = warp_sum(a);
  if (lane == 0) pred[b] = fmaxf(a + bias[0], 0.f);
}

__global__ void __launch_bounds__(128) fusion_bwd_kernel(const float* __restrict__ repr, const float* __restrict__ fpos,
                                                         const float* __restrict__ fneg, const float* __restrict__ w,
                                                         const float* __restrict__ pred, const float* __restrict__ d_pred, int B, int V,
                                                         float* __restrict__ d_repr, float* __restrict__ d_fpos,
                                                         float* __restrict__ d_fneg, float* __restrict__ d_w, float* __restrict__ d_b) {
  // one CTA handles a strip of samples; thread c owns feature column c (0..127), columns 128.. handled by threads < 2V
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * 8, b1 = min(B, b0 + 8);
  float dwc = 0.f, dwx = 0.f, dwy = 0.f, dbb = 0.f;   // x/y: extra columns tid and tid+128 (2V <= 256)
  const float wc = w[tid];
#pragma unroll 8
  for (int b = b0; b < b1; ++b) {
    const float dz = pred[b] > 0.f ? d_pred[b] : 0.f;
    d_repr[(size_t)b * D + tid] = dz * wc;
    dwc += dz * repr[(size_t)b * D + tid];
    if (tid < V) {
      d_fpos[(size_t)b * V + tid] = dz * w[D + tid];
      d_fneg[(size_t)b * V + tid] = dz * w[D + V + tid];
      dwx += dz * fpos[(size_t)b * V + tid];
      dwy += dz * fneg[(size_t)b * V + tid];
    }
    if (tid == 0) dbb += dz;
  }
  atomicAdd(&d_w[tid], dwc);
  if (tid < V) { atomicAdd(&d_w[D + tid], dwx); atomicAdd(&d_w[D + V + tid], dwy); }
  if (tid == 0) atomicAdd(d_b, dbb);
}

// ---------------------------------------------------------------- loss = mean((pred-label)^2) + rate * mean_{VxV}(pp^T pm + pn^T nm)
__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ labels,
                                                       const float* __restrict__ pp, const float* __restrict__ pn,
                                                       const float* __restrict__ pm, const float* __restrict__ nm, int B, int V,
                                                       float rate, float* __restrict__ loss /* zero-initialised */) {
  __shared__ float red[32];
  const int b = blockIdx.x * 256 + threadIdx.x;
  float a = 0.f;
  if (b < B) {
    const float d = pred[b] - labels[b];
    a = d * d / (float)B;
    if (pp) {
      float spp = 0.f, spn = 0.f, spm = 0.f, snm = 0.f;
      for (int v = 0; v < V; ++v) {
        spp += pp[(size_t)b * V + v]; spn += pn[(size_t)b * V + v];
        spm += pm[(size_t)b * V + v]; snm += nm[(size_t)b * V + v];
      }
      a += rate * (spp * spm + spn * snm) / (float)(V * V);
    }
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) atomicAdd(loss, a);
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ labels,
                                                       const float* __restrict__ pp, const float* __restrict__ pn,
                                                       const float* __restrict__ pm, const float* __restrict__ nm,
                                                       const float* __restrict__ d_loss, int B, int V, float rate,
                                                       float* __restrict__ d_pred, float* __restrict__ d_pp, float* __restrict__ d_pn,
                                                       float* __restrict__ d_pm, float* __restrict__ d_nm) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b >= B) return;
  const float dl = d_loss[0];
  d_pred[b] = dl * 2.f * (pred[b] - labels[b]) / (float)B;
  if (pp) {
    float spp = 0.f, spn = 0.f, spm = 0.f, snm = 0.f;
    for (int v = 0; v < V; ++v) {
      spp += pp[(size_t)b * V + v]; spn += pn[(size_t)b * V + v];
      spm += pm[(size_t)b * V + v]; snm += nm[(size_t)b * V + v];
    }
    const float k = dl * rate / (float)(V * V);
    for (int v = 0; v < V; ++v) {
      d_pp[(size_t)b * V + v] = k * spm; d_pm[(size_t)b * V + v] = k * spp;
      d_pn[(size_t)b * V + v] = k * snm; d_nm[(size_t)b * V + v] = k * spn;
    }
  }
}

// ---------------------------------------------------------------- dx = dy * (1 - y^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long n, float* __restrict__ dx) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float t = y[i]; dx[i] = dy[i] * (1.f - t * t); }
}

// ---------------------------------------------------------------- standalone SSNet (model.py:129-143): y = sigmoid(x · w + b), x (rows, 128)
__global__ void __launch_bounds__(128) ssnet_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                        long rows, float* __restrict__ y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 4 + warp;
  if (r >= rows) return;
  const float4 xv = *reinterpret_cast<const float4*>(x + r * D + lane * 4);
  const float4 wv = *reinterpret_cast<const float4*>(w + lane * 4);
  const float a = warp_sum(xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w);
  if (lane == 0) y[r] = sigmoidf_acc(a + b[0]);
}
// dpre = dy y (1-y);  dx = dpre w;  dw += sum_r dpre x_r;  db += sum_r dpre.  One CTA per strip of 32 rows, thread = feature column.
__global__ void __launch_bounds__(128) ssnet_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ y,
                                                        const float* __restrict__ dy, long rows, float* __restrict__ dx,
                                                        float* __restrict__ dw, float* __restrict__ db) {
  const int tid = threadIdx.x;
  const long r0 = (long)blockIdx.x * 32, r1 = r0 + 32 < rows ? r0 + 32 : rows;
  const float wc = w[tid];
  float dwc = 0.f, dbb = 0.f;
  for (long r = r0; r < r1; ++r) {
    const float yy = y[r], dp = dy[r] * yy * (1.f - yy);
    if (dx) dx[r * D + tid] = dp * wc;
    dwc += dp * x[r * D + tid];
    dbb += dp;
  }
  atomicAdd(&dw[tid], dwc);
  if (tid == 0) atomicAdd(db, dbb);
}

// ---------------------------------------------------------------- ControlNet tail (model.py:186-197)
__global__ void __launch_bounds__(128) control_tail_fwd_kernel(const float* __restrict__ s, const float* __restrict__ view_p,
                                                               const float* __restrict__ c_out, const float* __restrict__ ss_w,
                                                               const float* __restrict__ ss_b, float eps, int Su, int V,
                                                               float* __restrict__ senti, float* __restrict__ score,
                                                               float* __restrict__ prefer_pos, float* __restrict__ prefer_neg) {
  extern __shared__ float sm[];   // senti [Su]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float4 w4 = *reinterpret_cast<const float4*>(ss_w + lane * 4);
  for (int q = warp; q < Su; q += 4) {
    const float4 x = *reinterpret_cast<const float4*>(s + ((size_t)b * Su + q) * D + lane * 4);
    float a = x.x * w4.x + x.y * w4.y + x.z * w4.z + x.w * w4.w;
    a = warp_sum(a);
    if (lane == 0) { const float v = sigmoidf_acc(a + ss_b[0]); sm[q] = v; senti[(size_t)b * Su + q] = v; }
  }
  __syncthreads();
  for (int v = tid; v < V; v += 128) {
    float num = 0.f, den = 0.f;
    for (int q = 0; q < Su; ++q) {
      const float p = view_p[((size_t)b * Su + q) * V + v];
      num += sm[q] * p * p;
      den += p * p;
    }
    const float sc = num / (den + eps);                      // Eq. 18
    const float co = c_out[(size_t)b * V + v];
    const float qpos = sc < 0.5f ? 0.f : 4.f * (sc - 0.5f) * (sc - 0.5f);
    const float qneg = sc > 0.5f ? 0.f : 4.f * (0.5f - sc) * (0.5f - sc);
    const float qp = sc > 0.5f ? 1.f : 0.f;
    score[(size_t)b * V + v] = sc;
    prefer_pos[(size_t)b * V + v] = co * qp * qpos;
    prefer_neg[(size_t)b * V + v] = co * (1.f - qp) * qneg;
  }
}

__global__ void __launch_bounds__(128) control_tail_bwd_kernel(const float* __restrict__ s, const float* __restrict__ view_p,
                                                               const float* __restrict__ c_out, const float* __restrict__ ss_w,
                                                               const float* __restrict__ senti, const float* __restrict__ score,
                                                               const float* __restrict__ d_pp, const float* __restrict__ d_pn, float eps,
                                                               int Su, int V, float* __restrict__ d_s, float* __restrict__ d_view_p,
                                                               float* __restrict__ d_c_out, float* __restrict__ d_ss_w,
                                                               float* __restrict__ d_ss_b) {
  extern __shared__ float sm[];   // dnum [V], dden [V], dsenti [Su]
  float* dnum = sm;
  float* dden = sm + V;
  float* dsen = sm + 2 * V;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int v = tid; v < V; v += 128) {
    const float sc = score[(size_t)b * V + v], co = c_out[(size_t)b * V + v];
    const float gpp = d_pp[(size_t)b * V + v], gpn = d_pn[(size_t)b * V + v];
    const float qpos = sc < 0.5f ? 0.f : 4.f * (sc - 0.5f) * (sc - 0.5f);
    const float qneg = sc > 0.5f ? 0.f : 4.f * (0.5f - sc) * (0.5f - sc);
    const float qp = sc > 0.5f ? 1.f : 0.f;
    d_c_out[(size_t)b * V + v] = gpp * qp * qpos + gpn * (1.f - qp) * qneg;
    float dsc = 0.f;
    if (sc > 0.5f) dsc += gpp * co * 8.f * (sc - 0.5f);
    if (sc <= 0.5f) dsc += gpn * co * (-8.f) * (0.5f - sc);
    float num = 0.f, den = eps;
    for (int q = 0; q < Su; ++q) {
      const float p = view_p[((size_t)b * Su + q) * V + v];
      num += senti[(size_t)b * Su + q] * p * p;
      den += p * p;
    }
    dnum[v] = dsc / den;
    dden[v] = -dsc * num / (den * den);
  }
  __syncthreads();
  for (int q = tid; q < Su; q += 128) {
    float a = 0.f;
    const float se = senti[(size_t)b * Su + q];
    for (int v = 0; v < V; ++v) {
      const float p = view_p[((size_t)b * Su + q) * V + v];
      a += dnum[v] * p * p;
      d_view_p[((size_t)b * Su + q) * V + v] = 2.f * p * (dnum[v] * se + dden[v]);
    }
    dsen[q] = a * se * (1.f - se);     // through the sigmoid of SSNet
  }
  __syncthreads();
  float dw = 0.f, db = 0.f;
  const float wc = ss_w[tid];
  for (int q = 0; q < Su; ++q) {
    const float dp = dsen[q];
    d_s[((size_t)b * Su + q) * D + tid] = dp * wc;
    dw += dp * s[((size_t)b * Su + q) * D + tid];
    db += dp;
  }
  atomicAdd(&d_ss_w[tid], dw);
  if (tid == 0) atomicAdd(d_ss_b, db);
}

// ---------------------------------------------------------------- VisualNet tail (model.py:219-228)
// emb[0][v] = w · pos_v_emb[v] + b ; emb[1][v] = w · neg_v_emb[v] + b
__global__ void __launch_bounds__(128) visual_emb_kernel(const float* __restrict__ pos_e, const float* __restrict__ neg_e,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int V, int F,
                                                         float* __restrict__ emb) {
  __shared__ float red[32];
  const int v = blockIdx.x % V, which = blockIdx.x / V;
  const float* e = (which ? neg_e : pos_e) + (size_t)v * F;
  float a = 0.f;
  for (int k = threadIdx.x; k < F; k += 128) a += e[k] * w[k];
  a = block_sum(a, red);
  if (threadIdx.x == 0) emb[which * V + v] = a + bias[0];
}

__global__ void __launch_bounds__(128) visual_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ emb,
                                                         const float* __restrict__ c_u, const float* __restrict__ c_i, int BV, int V,
                                                         int Pc, int F, float* __restrict__ img_emb, float* __restrict__ pos_match,
                                                         float* __restrict__ neg_match, float* __restrict__ final_pos,
                                                         float* __restrict__ final_neg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bv = blockIdx.x * 4 + warp;
  if (bv >= BV) return;
  const int v = bv % V;
  const float* f = feat + (size_t)bv * Pc * F;
  float a = 0.f;
  if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
    // 16-byte loads, 8 of them in flight per lane (the 4 KB feature row of a (sample, view) is one round trip instead of 32)
    const int F4 = F >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float inv = 1.f / (float)Pc;
#pragma unroll 8
    for (int k = lane; k < F4; k += 32) {
      float4 m = f4[k];
      for (int p = 1; p < Pc; ++p) { const float4 t = f4[(size_t)p * F4 + k]; m.x += t.x; m.y += t.y; m.z += t.z; m.w += t.w; }
      const float4 ww = w4[k];
      a += (m.x * inv) * ww.x + (m.y * inv) * ww.y + (m.z * inv) * ww.z + (m.w * inv) * ww.w;
    }
  } else {
    for (int k = lane; k < F; k += 32) {
      float m = 0.f;
      for (int p = 0; p < Pc; ++p) m += f[(size_t)p * F + k];
      a += (m / (float)Pc) * w[k];
    }
  }
  a = warp_sum(a);
  if (lane == 0) {
    const float ie = a + bias[0];
    const float pm = tanhf(fabsf(emb[v] - ie)), nmv = tanhf(fabsf(emb[V + v] - ie));
    const float cc = c_u[bv] * c_i[bv];
    img_emb[bv] = ie; pos_match[bv] = pm; neg_match[bv] = nmv;
    final_pos[bv] = cc * (1.f - pm); final_neg[bv] = cc * (1.f - nmv);
  }
}

// per (b, v): gradients of the scalars; writes d(img_emb), d(pos_emb - img_emb), d(neg_emb - img_emb)
__global__ void __launch_bounds__(256) visual_bwd_scalar_kernel(const float* __restrict__ emb, const float* __restrict__ img_emb,
                                                                const float* __restrict__ pos_match, const float* __restrict__ neg_match,
                                                                const float* __restrict__ c_u, const float* __restrict__ c_i,
                                                                const float* __restrict__ d_pm, const float* __restrict__ d_nm,
                                                                const float* __restrict__ d_fp, const float* __restrict__ d_fn, int BV, int V,
                                                                float* __restrict__ d_cu, float* __restrict__ d_ci,
                                                                float* __restrict__ d_img, float* __restrict__ d_dp, float* __restrict__ d_dn) {
  const int bv = blockIdx.x * 256 + threadIdx.x;
  if (bv >= BV) return;
  const int v = bv % V;
  const float pm = pos_match[bv], nmv = neg_match[bv], cu = c_u[bv], ci = c_i[bv];
  const float gfp = d_fp ? d_fp[bv] : 0.f, gfn = d_fn ? d_fn[bv] : 0.f;
  const float gpm = (d_pm ? d_pm[bv] : 0.f) - gfp * cu * ci;
  const float gnm = (d_nm ? d_nm[bv] : 0.f) - gfn * cu * ci;
  const float t = gfp * (1.f - pm) + gfn * (1.f - nmv);
  d_cu[bv] = ci * t;
  d_ci[bv] = cu * t;
  const float dfp_ = emb[v] - img_emb[bv], dfn_ = emb[V + v] - img_emb[bv];
  const float sp = dfp_ > 0.f ? 1.f : (dfp_ < 0.f ? -1.f : 0.f);
  const float sn = dfn_ > 0.f ? 1.f : (dfn_ < 0.f ? -1.f : 0.f);
  const float ddp = gpm * (1.f - pm * pm) * sp;
  const float ddn = gnm * (1.f - nmv * nmv) * sn;
  d_dp[bv] = ddp; d_dn[bv] = ddn; d_img[bv] = -ddp - ddn;
}

// d_w[k] += sum_{rows} d_img[row] * mean_p feat[row][p][k]   (rows = B*V strips)
__global__ void __launch_bounds__(256) visual_bwd_w_kernel(const float* __restrict__ feat, const float* __restrict__ d_img, int BV, int Pc,
                                                           int F, int rows_per_cta, float* __restrict__ d_w) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(BV, r0 + rows_per_cta);
  if (k >= F) return;
  float a = 0.f;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    float m = 0.f;
    for (int p = 0; p < Pc; ++p) m += feat[((size_t)r * Pc + p) * F + k];
    a += d_img[r] * (m / (float)Pc);
  }
  atomicAdd(&d_w[k], a);
}

// per view v: d_pos_emb = sum_b d_dp[b,v]; d_pos_v_emb[v] = d_pos_emb * w; d_w += d_pos_emb * pos_v_emb[v] (same for neg); d_b
__global__ void __launch_bounds__(256) visual_bwd_views_kernel(const float* __restrict__ pos_e, const float* __restrict__ neg_e,
                                                               const float* __restrict__ w, const float* __restrict__ d_dp,
                                                               const float* __restrict__ d_dn, const float* __restrict__ d_img, int B, int V,
                                                               int F, float* __restrict__ d_pos_e, float* __restrict__ d_neg_e,
                                                               float* __restrict__ d_w, float* __restrict__ d_b) {
  __shared__ float red[32];
  const int v = blockIdx.x, tid = threadIdx.x;
  float sp = 0.f, sn = 0.f, si = 0.f;
  for (int b = tid; b < B; b += 256) { sp += d_dp[(size_t)b * V + v]; sn += d_dn[(size_t)b * V + v]; si += d_img[(size_t)b * V + v]; }
  sp = block_sum(sp, red);
  sn = block_sum(sn, red);
  si = block_sum(si, red);
  for (int k = tid; k < F; k += 256) {
    const float wk = w[k];
    d_pos_e[(size_t)v * F + k] = sp * wk;
    d_neg_e[(size_t)v * F + k] = sn * wk;
    atomicAdd(&d_w[k], sp * pos_e[(size_t)v * F + k] + sn * neg_e[(size_t)v * F + k]);
  }
  if (tid == 0) atomicAdd(d_b, sp + sn + si);
}

}  // namespace umpr

using namespace umpr;
#define ST (cudaStream_t)stream

extern "C" int umpr_fusion_fwd(const float* repr, const float* fpos, const float* fneg, const float* w, const float* bias, int B, int V,
                               float* pred, void* stream) {
  if (B <= 0) return 0;
  if (V < 0 || V > 128) return fail_arg("fusion: V=%d", V);
  fusion_fwd_kernel<<<(B + 3) / 4, 128, 0, ST>>>(repr, fpos, fneg, w, bias, B, fpos ? V : 0, pred);
  return check_launch("fusion_fwd");
}
extern "C" int umpr_fusion_bwd(const float* repr, const float* fpos, const float* fneg, const float* w, const float* pred,
                               const float* d_pred, int B, int V, float* d_repr, float* d_fpos, float* d_fneg, float* d_w, float* d_b,
                               void* stream) {
  if (B <= 0) return 0;
  if (V < 0 || V > 128) return fail_arg("fusion: V=%d", V);
  fusion_bwd_kernel<<<(B + 7) / 8, 128, 0, ST>>>(repr, fpos, fneg, w, pred, d_pred, B, fpos ? V : 0, d_repr, d_fpos, d_fneg, d_w, d_b);
  return check_launch("fusion_bwd");
}
extern "C" int umpr_loss_fwd(const float* pred, const float* labels, const float* pp, const float* pn, const float* pm, const float* nm,
                             int B, int V, float rate, float* loss, void* stream) {
  if (B <= 0) return fail_arg("loss: empty batch");
  loss_fwd_kernel<<<(B + 255) / 256, 256, 0, ST>>>(pred, labels, pp, pn, pm, nm, B, V, rate, loss);
  return check_launch("loss_fwd");
}
extern "C" int umpr_loss_bwd(const float* pred, const float* labels, const float* pp, const float* pn, const float* pm, const float* nm,
                             const float* d_loss, int B, int V, float rate, float* d_pred, float* d_pp, float* d_pn, float* d_pm,
                             float* d_nm, void* stream) {
  if (B <= 0) return 0;
  loss_bwd_kernel<<<(B + 255) / 256, 256, 0, ST>>>(pred, labels, pp, pn, pm, nm, d_loss, B, V, rate, d_pred, d_pp, d_pn, d_pm, d_nm);
  return check_launch("loss_bwd");
}
extern "C" int umpr_tanh_bwd(const float* y, const float* dy, long n, float* dx, void* stream) {
  if (n <= 0) return 0;
  tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST>>>(y, dy, n, dx);
  return check_launch("tanh_bwd");
}
extern "C" int umpr_ssnet_fwd(const float* x, const float* w, const float* b, long rows, float* y, void* stream) {
  if (rows <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) return fail_arg("ssnet_fwd: x and w must be 16-byte aligned");
  ssnet_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(x, w, b, rows, y);
  return check_launch("ssnet_fwd");
}
extern "C" int umpr_ssnet_bwd(const float* x, const float* w, const float* y, const float* dy, long rows, float* dx /* may be NULL */,
                              float* dw /* += */, float* db /* += */, void* stream) {
  if (rows <= 0) return 0;
  ssnet_bwd_kernel<<<(unsigned)((rows + 31) / 32), 128, 0, (cudaStream_t)stream>>>(x, w, y, dy, rows, dx, dw, db);
  return check_launch("ssnet_bwd");
}

extern "C" int umpr_control_tail_fwd(const float* s, const float* view_p, const float* c_out, const float* ss_w, const float* ss_b,
                                     float eps, int B, int Su, int V, float* senti, float* score, float* prefer_pos, float* prefer_neg,
                                     void* stream) {
  if (B <= 0) return 0;
  control_tail_fwd_kernel<<<B, 128, sizeof(float) * Su, ST>>>(s, view_p, c_out, ss_w, ss_b, eps, Su, V, senti, score, prefer_pos, prefer_neg);
  return check_launch("control_tail_fwd");
}
extern "C" int umpr_control_tail_bwd(const float* s, const float* view_p, const float* c_out, const float* ss_w, const float* senti,
                                     const float* score, const float* d_pp, const float* d_pn, float eps, int B, int Su, int V,
                                     float* d_s, float* d_view_p, float* d_c_out, float* d_ss_w, float* d_ss_b, void* stream) {
  if (B <= 0) return 0;
  control_tail_bwd_kernel<<<B, 128, sizeof(float) * (2 * V + Su), ST>>>(s, view_p, c_out, ss_w, senti, score, d_pp, d_pn, eps, Su, V, d_s,
                                                                      d_view_p, d_c_out, d_ss_w, d_ss_b);
  return check_launch("control_tail_bwd");
}
extern "C" int umpr_visual_fwd(const float* feat, const float* pos_e, const float* neg_e, const float* w, const float* bias,
                               const float* c_u, const float* c_i, int B, int V, int Pc, int F, float* emb /*2V*/, float* img_emb,
                               float* pos_match, float* neg_match, float* final_pos, float* final_neg, void* stream) {
  if (B <= 0) return 0;
  visual_emb_kernel<<<2 * V, 128, 0, ST>>>(pos_e, neg_e, w, bias, V, F, emb);
  if (int e = check_launch("visual_emb")) return e;
  visual_fwd_kernel<<<(B * V + 3) / 4, 128, 0, ST>>>(feat, w, bias, emb, c_u, c_i, B * V, V, Pc, F, img_emb, pos_match, neg_match,
                                                   final_pos, final_neg);
  return check_launch("visual_fwd");
}
extern "C" int umpr_visual_bwd(const float* feat, const float* pos_e, const float* neg_e, const float* w, const float* emb,
                               const float* img_emb, const float* pos_match, const float* neg_match, const float* c_u, const float* c_i,
                               const float* d_pm, const float* d_nm, const float* d_fp, const float* d_fn, int B, int V, int Pc, int F,
                               float* scratch /*3*B*V*/, float* d_cu, float* d_ci, float* d_pos_e, float* d_neg_e, float* d_w, float* d_b,
                               void* stream) {
  if (B <= 0) return 0;
  const int BV = B * V;
  float* d_img = scratch; float* d_dp = scratch + BV; float* d_dn = scratch + 2 * BV;
  visual_bwd_scalar_kernel<<<(BV + 255) / 256, 256, 0, ST>>>(emb, img_emb, pos_match, neg_match, c_u, c_i, d_pm, d_nm, d_fp, d_fn, BV, V,
                                                           d_cu, d_ci, d_img, d_dp, d_dn);
  if (int e = check_launch("visual_bwd_scalar")) return e;
  const int rows_per = 16;
  visual_bwd_w_kernel<<<dim3((F + 255) / 256, (BV + rows_per - 1) / rows_per), 256, 0, ST>>>(feat, d_img, BV, Pc, F, rows_per, d_w);
  if (int e = check_launch("visual_bwd_w")) return e;
  visual_bwd_views_kernel<<<V, 256, 0, ST>>>(pos_e, neg_e, w, d_dp, d_dn, d_img, B, V, F, d_pos_e, d_neg_e, d_w, d_b);
  return check_launch("visual_bwd_views");
}
