// One C-ABI call per train / eval step: the whole UMPR forward and backward (reference src/model.py:257-278 under main.py:32-36,
// evaluate.py:8-11) issued from native host code.  The Python autograd path (umpr_b200/functional.py) launches the same kernels
// through ~60 ctypes calls, tensor allocations and autograd nodes per step - 4-5 ms of host time, which bounds the step at
// small batches and makes eight ranks wait for the slowest host.  Here the sequence is straight-line C++: one bump-allocated
// workspace, ~80 kernel launches, a few hundred microseconds of host time.
//
// What the host still does (Python, off the critical path on a worker thread): the reference's own torch.sort call and the
// integer pack plans / valid-row tables / tile schedules derived from it (umpr_b200/plan.py), uploaded in one buffer.
// Parameter gradients are ACCUMULATED into the caller's gradient bucket (zeroed by the caller, all-reduced and consumed by
// umpr_adam_step afterwards), exactly as functional._sinks does under FlatTrainer.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

// ---- optional per-entry-point timing (bench.py's kernel table and roofline): CUDA-event pairs on the launching stream around the
// entry points of one umpr_step call - all of them, or only the one named in umpr_step_profile_begin.  Off by default: no events.
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
struct Profiler {
  bool on = false;
  std::string only;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t event() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
};
static Profiler g_prof;
// gradient exchange overlapped with the backward (umpr_step_comm): the flat bucket is [early | late]; `early` (everything but R-Net's
// GRU gradients) is final before the last kernel of the backward - the fused R-Net GRU backward - starts, and is all-reduced on a
// communication stream WHILE that kernel runs; `late` follows when it has finished (two hand-placed buckets)
struct Overlap { void* comm = nullptr; float* bucket = nullptr; long n_early = 0, n_total = 0; };
static Overlap g_ov;
static bool g_single_stream = false;      // UMPR_STEP_SINGLE_STREAM=1 / umpr_step_streams(1): never fork (debugging / timing)
static bool g_streams_set = false;
struct ProfScope {
  cudaEvent_t e1 = nullptr;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t s) : st(s) {
    if (!g_prof.on || (!g_prof.only.empty() && g_prof.only != name)) return;
    ProfRec r{name, g_prof.event(), g_prof.event()};
    cudaEventRecord(r.e0, st);
    e1 = r.e1;
    g_prof.recs.push_back(r);
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, st); }
};

struct Arena {
  unsigned char* base;
  size_t off, cap;
  template <class T> T* get(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

#define UMPR_TRY(call) do { if (int rc_ = (call)) return rc_; } while (0)
#define STEP_CALL(name, call) do { ProfScope ps_(name, st); if (int rc_ = (call)) return rc_; } while (0)
#define STEP_CALL_C(name, call) do { ProfScope ps_(name, stc); if (int rc_ = (call)) return rc_; } while (0)      // on the C-Net branch's stream

__global__ void expand_rows_kernel(const float* __restrict__ src, long n_rows, int L, float* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows * L) dst[i] = src[i / L];
}
__global__ void set_scalar_kernel(float* p, float v) { *p = v; }
__global__ void add_inplace_kernel(float4* __restrict__ a, const float4* __restrict__ b, long n4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) { float4 x = a[i]; const float4 y = b[i]; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; a[i] = x; }
}

// everything one review side keeps between forward and backward
struct SideBufs {
  void* xq; void* hq_r; void* hq_c;            // token images; hidden images of the R-Net / C-Net GRU (training)
  float* out_r; float* out_c;                  // ImprovedRnn results (B, S*L, 128)
  float* self_atte; float* wsum; float* senti; // S-Net
  float* cfeat; int32_t* cidx; float* view_p; float* fin;      // C-Net tail
  float* dx_r; float* dx_c; float* dx_s;       // gradients of out_r (co-attention), out_c (convolution), and S-Net's share
  float* dcfeat; float* d_sa; float* d_wsum; float* d_soft;
};

static size_t side_tokens_rows(const umpr_step_side& s) { return (size_t)s.B * s.S * s.L; }

// the step proper; with a.base == nullptr nothing is launched and only the workspace size is computed
static int run_step(const umpr_step_model& m, const umpr_step_side* sd, const float* photos, const float* labels,
                    const int32_t* sched_r, int nq_r, const int32_t* sched_c, int nq_c, const void* zero_img, Arena& a, float* pred_out,
                    float* loss_out, int train, int n_ctas, void* stream) {
  const bool dry = a.base == nullptr;
  const int order_f[3] = {2, 0, 1};          // the order control_net calls c_net in: ui, user, item
  const bool full = !m.review_net_only;
  const int B = sd[0].B, P = sd[0].S * sd[0].L, V = m.V, KC = m.KC, E = m.E, Dm = 128;
  const int n_sides = full ? 3 : 2;
  if (sd[1].B != B || sd[1].S * sd[1].L != P) return fail_arg("step: user and item sides must share (B, S*L)");
  cudaStream_t st = (cudaStream_t)stream;
  SideBufs sb[3];
  memset(sb, 0, sizeof(sb));
  // ------------------------------------------------------------------------------------------------ workspace layout
  for (int k = 0; k < n_sides; ++k) {
    const umpr_step_side& s = sd[k];
    const size_t rows = side_tokens_rows(s), N = (size_t)s.B * s.S;
    sb[k].xq = a.get<unsigned char>((size_t)s.n_slabs * 32768);
    if (k < 2) {
      sb[k].out_r = a.get<float>(rows * Dm);
      if (train) sb[k].hq_r = a.get<unsigned char>((size_t)s.n_slabs * 2 * 32768);
      sb[k].self_atte = a.get<float>(N * Dm);
      sb[k].wsum = a.get<float>(N);
      sb[k].senti = a.get<float>((size_t)B * Dm);
      if (train) {
        sb[k].dx_r = a.get<float>(rows * Dm);
        sb[k].dx_s = a.get<float>(rows * Dm);
        sb[k].d_sa = a.get<float>(N * Dm);
        sb[k].d_wsum = a.get<float>(N);
        sb[k].d_soft = a.get<float>(rows);
      }
    }
    if (full) {
      sb[k].out_c = a.get<float>(rows * Dm);
      if (train) sb[k].hq_c = a.get<unsigned char>((size_t)s.n_slabs * 2 * 32768);
      sb[k].cfeat = a.get<float>(N * KC);
      sb[k].cidx = a.get<int32_t>(N * KC);
      sb[k].view_p = a.get<float>(N * V);
      sb[k].fin = a.get<float>((size_t)B * V);
      if (train) {
        sb[k].dx_c = a.get<float>(rows * Dm);
        sb[k].dcfeat = a.get<float>(N * KC);
      }
    }
  }
  const size_t BP = (size_t)B * P;
  float* giM = a.get<float>(BP * Dm);
  float* soft = a.get<float>(4 * BP);               // soft_u, soft_i, t_u, t_i
  int32_t* arg = a.get<int32_t>(2 * BP);
  float* atte = a.get<float>(2 * (size_t)B * Dm);
  long long co_bytes = 0;
  UMPR_TRY(umpr_workspace_bytes("coattn_fwd_tc", B, P, &co_bytes));
  void* co_scratch = a.get<unsigned char>((size_t)co_bytes);
  float* repr = a.get<float>((size_t)B * Dm);
  float* pred = pred_out ? pred_out : a.get<float>(B);
  float* loss = loss_out ? loss_out : a.get<float>(1);
  // full model
  int Nmax = 0;
  for (int k = 0; k < n_sides; ++k) Nmax = sd[k].B * sd[k].S > Nmax ? sd[k].B * sd[k].S : Nmax;
  const int cap = ((Nmax * KC + 7) / 8) > 4096 ? ((Nmax * KC + 7) / 8) : 4096;      // one 2-byte re-scoring record per (sentence, filter)
  void* conv_scratch = nullptr; float* s_ui = nullptr; float* senti_ss = nullptr; float* ct_out = nullptr; float* vis_emb = nullptr; float* vis_out = nullptr;
  void* dx_scratch[3] = {nullptr, nullptr, nullptr};
  void* fix_rec[3] = {nullptr, nullptr, nullptr};
  if (full) {
    long long cb = 0;
    UMPR_TRY(umpr_workspace_bytes("cnet_conv_fwd_tc", cap, 0, &cb));
    conv_scratch = a.get<unsigned char>((size_t)cb);
    for (int k = 0; k < 3; ++k) fix_rec[k] = a.get<unsigned char>(2 * (size_t)sd[k].B * sd[k].S * KC + 16);      // re-scoring records per side (the sides run concurrently)
    s_ui = a.get<float>((size_t)sd[2].B * sd[2].S * Dm);
    senti_ss = a.get<float>((size_t)sd[2].B * sd[2].S);
    ct_out = a.get<float>(3 * (size_t)B * V);       // score, prefer_pos, prefer_neg
    vis_emb = a.get<float>(2 * (size_t)V);
    vis_out = a.get<float>(5 * (size_t)B * V);      // img_emb, pos_match, neg_match, final_pos, final_neg
    if (train) for (int k = 0; k < 3; ++k) dx_scratch[k] = a.get<unsigned char>(196608);
  }
  // backward temporaries
  float* d_pred = nullptr, *g4 = nullptr, *d_repr = nullptr, *d_f = nullptr, *vis_scr = nullptr, *d_c = nullptr, *d_s = nullptr, *d_vp = nullptr,
        *d_co = nullptr, *dpre = nullptr, *dins = nullptr, *dgiM = nullptr, *dx_s_ui = nullptr, *one = nullptr;
  if (train) {
    d_pred = a.get<float>(B);
    d_repr = a.get<float>((size_t)B * Dm);
    dpre = a.get<float>((size_t)B * Dm);
    dins = a.get<float>(4 * (size_t)B * Dm);
    dgiM = a.get<float>(BP * Dm);
    one = a.get<float>(1);
    if (full) {
      g4 = a.get<float>(4 * (size_t)B * V);
      d_f = a.get<float>(2 * (size_t)B * V);
      vis_scr = a.get<float>(3 * (size_t)B * V);
      d_c = a.get<float>(2 * (size_t)B * V);
      d_s = a.get<float>((size_t)sd[2].B * sd[2].S * Dm);
      d_vp = a.get<float>((size_t)sd[2].B * sd[2].S * V);
      d_co = a.get<float>((size_t)B * V);
      dx_s_ui = a.get<float>(side_tokens_rows(sd[2]) * Dm);
    }
  }
  if (dry) return 0;
  if (a.off > a.cap) return fail_arg("step: workspace of %zu bytes needed, %zu given", a.off, a.cap);

  // ------------------------------------------------------------------------------------------------ forward
  {                                          // model.py:262-264 fused into the pack half of model.py:18, all sides in one launch
    const int64_t* g_ids[3]; const int32_t* g_plan[3]; int g_nt[3], g_ns[3], g_L[3]; void* g_xq[3];
    for (int k = 0; k < n_sides; ++k) {
      g_ids[k] = sd[k].ids; g_plan[k] = sd[k].plan; g_nt[k] = sd[k].n_tiles; g_ns[k] = sd[k].n_slabs; g_L[k] = sd[k].L; g_xq[k] = sb[k].xq;
    }
    STEP_CALL("umpr_gather_pack_tc", gather_pack_tc_sides(m.table, n_sides, g_ids, g_plan, g_nt, g_ns, g_L, E, g_xq, stream));
  }
  // The C-Net branch (its GRU, convolution tails, ControlNet tail) runs on a second stream beside the R-Net branch, forward and
  // backward.  Small batches leave most SMs idle inside one GRU launch (a tile is a serial chain of time steps); at large batches every
  // kernel is a persistent one-CTA-per-SM grid whose last wave and prologue leave SMs idle - the other branch's kernels fill them
  // (batch 1024: 3.89 -> 3.59 ms per step).
  int tiles_all = 0;
  for (int k = 0; k < n_sides; ++k) tiles_all += sd[k].n_tiles * (k < 2 && full ? 2 : 1);      // user, item tiles are in both launches
  // the fourth stream pays for its event traffic only when the kernels are long (small batches are bound by the issuing thread)
  const bool par = !g_single_stream && 2 * tiles_all > n_ctas + n_ctas / 4;
  const bool two = full && !g_single_stream;
  cudaStream_t stc = st;
  void* cstream = stream;
  // the library's side streams and events belong to a device: one set per device of the process
  struct DevStreams { cudaStream_t side = nullptr, side3 = nullptr, side4 = nullptr, comm = nullptr; cudaEvent_t ev[12], cev[3]; };
  static DevStreams dev_streams[64];
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (cur_dev < 0 || cur_dev >= 64) return fail_arg("step: device %d", cur_dev);
  DevStreams& ds = dev_streams[cur_dev];
  cudaStream_t& side = ds.side; cudaStream_t& side3 = ds.side3; cudaStream_t& side4 = ds.side4;
  cudaEvent_t* ev = ds.ev;
  cudaStream_t st3 = st;                     // third stream: the item side of the C-Net tails (convolution, heads) beside ui + user
  cudaStream_t st4 = st;                     // fourth: S-Net beside the co-attention (forward), the item side's S-Net backward, dM
  if (par || two) {
    if (!side) {
      if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) { side = nullptr; return fail_arg("step: cannot create the side stream"); }
      if (cudaStreamCreateWithFlags(&side3, cudaStreamNonBlocking) != cudaSuccess) { side3 = nullptr; return fail_arg("step: cannot create the side stream"); }
      if (cudaStreamCreateWithFlags(&side4, cudaStreamNonBlocking) != cudaSuccess) { side4 = nullptr; return fail_arg("step: cannot create the side stream"); }
      for (int i = 0; i < 12; ++i) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    if (par) st4 = side4;
  }
  if (two) {
    stc = side;
    cstream = side;
    st3 = side3;
    cudaEventRecord(ev[0], st);              // fork: the token images are packed
    cudaStreamWaitEvent(stc, ev[0], 0);
  }
  {                                          // R-Net's GRU over user + item, one launch (model.py:45-46)
    umpr_gru_seg segs[2];
    for (int k = 0; k < 2; ++k)
      segs[k] = umpr_gru_seg{sb[k].xq, sd[k].plan, sb[k].out_r, nullptr, sb[k].hq_r, sd[k].n_tiles, sd[k].n_slabs, sd[k].B * sd[k].S, sd[k].L};
    STEP_CALL("umpr_gru_fwd_tc", umpr_gru_fwd_tc(segs, 2, m.rnet_gru, E, sched_r, nq_r, stream));
  }
  const float* gu = sb[0].out_r, *gi = sb[1].out_r;
  // S-Net of each side (model.py:162-163) needs the GRU output only: it runs on the fourth stream beside giM and the co-attention
  if (par) { cudaEventRecord(ev[7], st); cudaStreamWaitEvent(st4, ev[7], 0); }
  for (int k = 0; k < 2; ++k) {
    const float* Ms = k ? m.snet_i_Ms : m.snet_u_Ms, *Ws = k ? m.snet_i_Ws : m.snet_u_Ws;
    ProfScope ps_("umpr_snet_fwd_tc", st4);
    UMPR_TRY(umpr_snet_fwd_tc(sb[k].out_r, sd[k].snet_table, sd[k].snet_tiles, Ms, Ws, sd[k].B * sd[k].S, sd[k].L, sb[k].self_atte, n_ctas, st4));
  }
  if (par) cudaEventRecord(ev[8], st4);
  // co-attention (model.py:50-55): giM = gi · M over the valid rows, flash-style affinity on tcgen05
  if (BP >= 1024)
    STEP_CALL("umpr_tc_gemm_ws", umpr_tc_gemm_ws(gi, Dm, m.M, Dm, giM, Dm, (int)BP, Dm, Dm, 0, nullptr, 0, 1, sd[1].snet_table, sd[1].snet_tiles, sd[1].L, n_ctas, stream));
  else
    STEP_CALL("umpr_tc_gemm_nt", umpr_tc_gemm_nt(gi, Dm, m.M, Dm, giM, Dm, (int)BP, Dm, Dm, 0, nullptr, 0, 1, stream));
  const int32_t* cst_u = sd[0].snet_table + sd[0].snet_tiles + 1, *cst_i = sd[1].snet_table + sd[1].snet_tiles + 1;
  const int pv_max = sd[0].pv_max > sd[1].pv_max ? sd[0].pv_max : sd[1].pv_max;
  STEP_CALL("umpr_coattn_fwd_tc", umpr_coattn_fwd_tc(gu, gi, giM, B, P, cst_u, sd[0].S, sd[0].L, cst_i, sd[1].S, sd[1].L, pv_max, co_scratch, soft, soft + BP,
                              soft + 2 * BP, soft + 3 * BP, arg, arg + BP, atte, atte + (size_t)B * Dm, stream));
  if (m.routing_coattn) cudaMemcpyAsync(m.routing_coattn, arg, 2 * BP * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
  if (par) cudaStreamWaitEvent(st, ev[8], 0);          // join: self_atte of both sides
  for (int k = 0; k < 2; ++k)
    STEP_CALL("umpr_snet_sentiment_fwd", umpr_snet_sentiment_fwd(sb[k].self_atte, soft + k * BP, B, sd[k].S, sd[k].L, sb[k].wsum, sb[k].senti, stream));
  STEP_CALL("umpr_text_match_fwd", umpr_text_match_fwd(atte, sb[0].senti, atte + (size_t)B * Dm, sb[1].senti, m.lin_u, m.lin_i, B, repr, stream));      // model.py:166-168
  const float* pp = nullptr, *pn = nullptr, *pm = nullptr, *nm = nullptr, *fpos = nullptr, *fneg = nullptr;
  if (full) {
    {                                        // C-Net's GRU over ui + user + item, one launch (model.py:182-184)
      umpr_gru_seg segs[3];
      const int order[3] = {2, 0, 1};
      for (int j = 0; j < 3; ++j) {
        const int k = order[j];
        segs[j] = umpr_gru_seg{sb[k].xq, sd[k].plan, sb[k].out_c, nullptr, sb[k].hq_c, sd[k].n_tiles, sd[k].n_slabs, sd[k].B * sd[k].S, sd[k].L};
      }
      STEP_CALL_C("umpr_gru_fwd_tc", umpr_gru_fwd_tc(segs, 3, m.cnet_gru, E, sched_c, nq_c, cstream));
    }
    // conv + ReLU + max-pool + view head (model.py:118-125) per side: ui and user on the C-Net stream, item on a third one
    STEP_CALL_C("umpr_cnet_conv_fwd_tc", cnet_conv_fwd_tc_impl(sb[0].out_c, m.conv_w, m.conv_b, 1, sd[0].L, KC, m.ksize, nullptr, 0, conv_scratch, cap, nullptr, nullptr,
                                   n_ctas, 2, nullptr, cstream));      // the weight image, once for the three sides
    if (two) { cudaEventRecord(ev[4], stc); cudaStreamWaitEvent(st3, ev[4], 0); }
    for (int j = 0; j < 3; ++j) {
      const int k = order_f[j];
      const int N = sd[k].B * sd[k].S;
      cudaStream_t sk = k == 1 ? st3 : stc;
      { ProfScope ps_("umpr_cnet_conv_fwd_tc", sk);
        UMPR_TRY(cnet_conv_fwd_tc_impl(sb[k].out_c, m.conv_w, m.conv_b, N, sd[k].L, KC, m.ksize, sd[k].cnet_table, sd[k].cnet_tiles, conv_scratch, cap,
                                       sb[k].cfeat, sb[k].cidx, n_ctas, 0, fix_rec[k], sk)); }
      { ProfScope ps_("umpr_cnet_head_fwd", sk);
        UMPR_TRY(umpr_cnet_head_fwd(sb[k].cfeat, m.clin_w, m.clin_b, m.threshold, sd[k].B, sd[k].S, V, KC, sb[k].view_p, sb[k].fin, sk)); }
      if (m.routing_cnet[k]) cudaMemcpyAsync(m.routing_cnet[k], sb[k].cidx, (size_t)N * KC * sizeof(int32_t), cudaMemcpyDeviceToDevice, sk);
    }
    if (two) { cudaEventRecord(ev[5], st3); cudaStreamWaitEvent(st, ev[5], 0); }      // the visual tail needs c_i
    // ControlNet tail (model.py:185-197): S-Net on the user->item review (its `sentiment` output is unused), SSNet + Eq.18 + gates
    STEP_CALL_C("umpr_snet_fwd_tc", umpr_snet_fwd_tc(sb[2].out_c, sd[2].snet_table, sd[2].snet_tiles, m.csnet_Ms, m.csnet_Ws, sd[2].B * sd[2].S, sd[2].L, s_ui, n_ctas, cstream));
    STEP_CALL_C("umpr_control_tail_fwd", umpr_control_tail_fwd(s_ui, sb[2].view_p, sb[2].fin, m.ss_w, m.ss_b, m.eq18_eps, B, sd[2].S, V, senti_ss, ct_out, ct_out + (size_t)B * V,
                                   ct_out + 2 * (size_t)B * V, cstream));
    if (two) { cudaEventRecord(ev[1], stc); cudaStreamWaitEvent(st, ev[1], 0); }      // join: the visual tail needs c_u, c_i
    // VisualNet tail (model.py:219-228)
    STEP_CALL("umpr_visual_fwd", umpr_visual_fwd(photos, m.pos_e, m.neg_e, m.vis_w, m.vis_b, sb[0].fin, sb[1].fin, B, V, m.Pc, m.F, vis_emb, vis_out, vis_out + (size_t)B * V,
                             vis_out + 2 * (size_t)B * V, vis_out + 3 * (size_t)B * V, vis_out + 4 * (size_t)B * V, stream));
    pp = ct_out + (size_t)B * V; pn = ct_out + 2 * (size_t)B * V;
    pm = vis_out + (size_t)B * V; nm = vis_out + 2 * (size_t)B * V;
    fpos = vis_out + 3 * (size_t)B * V; fneg = vis_out + 4 * (size_t)B * V;
  }
  STEP_CALL("umpr_fusion_fwd", umpr_fusion_fwd(repr, fpos, fneg, m.fus_w, m.fus_b, B, full ? V : 0, pred, stream));          // model.py:268,274
  cudaMemsetAsync(loss, 0, sizeof(float), st);
  STEP_CALL("umpr_loss_fwd", umpr_loss_fwd(pred, labels, pp, pn, pm, nm, B, full ? V : 0, m.loss_v_rate, loss, stream));   // model.py:269,275-277
  if (!train) return 0;

  // ------------------------------------------------------------------------------------------------ backward (main.py:36)
  set_scalar_kernel<<<1, 1, 0, st>>>(one, 1.0f);                                 // d(loss) = 1
  STEP_CALL("umpr_loss_bwd", umpr_loss_bwd(pred, labels, pp, pn, pm, nm, one, B, full ? V : 0, m.loss_v_rate, d_pred, full ? g4 : nullptr, full ? g4 + (size_t)B * V : nullptr,
                         full ? g4 + 2 * (size_t)B * V : nullptr, full ? g4 + 3 * (size_t)B * V : nullptr, stream));
  STEP_CALL("umpr_fusion_bwd", umpr_fusion_bwd(repr, fpos, fneg, m.fus_w, pred, d_pred, B, full ? V : 0, d_repr, full ? d_f : nullptr, full ? d_f + (size_t)B * V : nullptr,
                           m.g_fus_w, m.g_fus_b, stream));
  if (full) {
    STEP_CALL("umpr_visual_bwd", umpr_visual_bwd(photos, m.pos_e, m.neg_e, m.vis_w, vis_emb, vis_out, pm, nm, sb[0].fin, sb[1].fin, g4 + 2 * (size_t)B * V, g4 + 3 * (size_t)B * V,
                             d_f, d_f + (size_t)B * V, B, V, m.Pc, m.F, vis_scr, d_c, d_c + (size_t)B * V, m.g_pos_e, m.g_neg_e, m.g_vis_w, m.g_vis_b, stream));
    if (two) { cudaEventRecord(ev[2], st); cudaStreamWaitEvent(stc, ev[2], 0); }      // fork: d(c_u), d(c_i), d(prefer_*) are there
    STEP_CALL_C("umpr_control_tail_bwd", umpr_control_tail_bwd(s_ui, sb[2].view_p, sb[2].fin, m.ss_w, senti_ss, ct_out, g4, g4 + (size_t)B * V, m.eq18_eps, B, sd[2].S, V, d_s, d_vp, d_co,
                                   m.g_ss_w, m.g_ss_b, cstream));
    // S-Net on the user->item review: only self_atte was used, so d(self_atte) = d_s as it is
    STEP_CALL_C("umpr_snet_bwd_tc", umpr_snet_bwd_tc(sb[2].out_c, sd[2].snet_table, sd[2].snet_tiles, d_s, m.csnet_Ms, m.csnet_Ws, sd[2].B * sd[2].S, sd[2].L, dx_s_ui,
                              m.g_csnet_Ms, m.g_csnet_Ws, n_ctas, cstream));
    if (two) cudaStreamWaitEvent(st3, ev[2], 0);                                   // the item side's chain runs on the third stream
    for (int j = 0; j < 3; ++j) {
      const int k = order_f[j];
      const int N = sd[k].B * sd[k].S;
      cudaStream_t sk = k == 1 ? st3 : stc;
      const float* d_view = k == 2 ? d_vp : nullptr;
      const float* d_fin = k == 2 ? d_co : d_c + (size_t)k * B * V;              // c_u, c_i feed the visual tail; c_net_out the control tail
      { ProfScope ps_("umpr_cnet_head_bwd", sk);
        UMPR_TRY(umpr_cnet_head_bwd(sb[k].cfeat, sb[k].cidx, sb[k].view_p, m.clin_w, d_view, d_fin, sd[k].B, sd[k].S, V, KC, sb[k].dcfeat, m.g_clin_w, m.g_clin_b,
                                    m.g_conv_b, sk)); }
      { ProfScope ps_("umpr_cnet_conv_bwd_dx_tc", sk);
        UMPR_TRY(umpr_cnet_conv_bwd_dx_tc(sb[k].dcfeat, sb[k].cidx, m.conv_w, N, sd[k].L, KC, sd[k].cnet_table, sd[k].cnet_tiles, dx_scratch[k], sb[k].dx_c, n_ctas, sk)); }
      { ProfScope ps_("umpr_cnet_conv_bwd_dw_tc", sk);
        UMPR_TRY(umpr_cnet_conv_bwd_dw_tc(sb[k].out_c, sb[k].dcfeat, sb[k].cidx, N, sd[k].L, KC, sd[k].cnet_table, sd[k].cnet_tiles, m.g_conv_w, n_ctas, sk)); }
    }
    if (two) { cudaEventRecord(ev[6], st3); cudaStreamWaitEvent(stc, ev[6], 0); }      // join: the C-Net GRU backward needs all three dx
    {                                        // the user->item GRU output feeds the convolution AND S-Net: sum of both gradients
      const long n4 = (long)(side_tokens_rows(sd[2]) * Dm / 4);
      add_inplace_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stc>>>(reinterpret_cast<float4*>(sb[2].dx_c), reinterpret_cast<const float4*>(dx_s_ui), n4);
      UMPR_TRY(check_launch("step add"));
    }
    umpr_gru_bwd_seg segs[3];
    const int order[3] = {2, 0, 1};
    for (int j = 0; j < 3; ++j) {
      const int k = order[j];
      segs[j] = umpr_gru_bwd_seg{sb[k].dx_c, nullptr, sb[k].xq, sb[k].hq_c, sd[k].plan, sd[k].n_tiles, sd[k].n_slabs, sd[k].B * sd[k].S, sd[k].L};
    }
    STEP_CALL_C("umpr_gru_bwd_tc", umpr_gru_bwd_tc(segs, 3, m.cnet_gru, m.g_cnet_gru, E, zero_img, sched_c, nq_c, cstream));
  }
  // text matching (model.py:166-168)
  STEP_CALL("umpr_tanh_bwd", umpr_tanh_bwd(repr, d_repr, (long)B * Dm, dpre, stream));
  STEP_CALL("umpr_text_match_bwd", umpr_text_match_bwd(dpre, m.lin_u, m.lin_i, B, dins, dins + (size_t)B * Dm, dins + 2 * (size_t)B * Dm, dins + 3 * (size_t)B * Dm, stream));
  STEP_CALL("umpr_text_match_wgrad", umpr_text_match_wgrad(dpre, atte, sb[0].senti, atte + (size_t)B * Dm, sb[1].senti, B, m.g_lin_u, m.g_lin_i, stream));
  // S-Net of each side; its input gradient is handed to the co-attention backward (add_u / add_i).  User side here, item side on the
  // fourth stream
  if (par) { cudaEventRecord(ev[9], st); cudaStreamWaitEvent(st4, ev[9], 0); }
  for (int k = 0; k < 2; ++k) {
    cudaStream_t sk = k ? st4 : st;
    const float* Ms = k ? m.snet_i_Ms : m.snet_u_Ms, *Ws = k ? m.snet_i_Ws : m.snet_u_Ws;
    float* gMs = k ? m.g_snet_i_Ms : m.g_snet_u_Ms, *gWs = k ? m.g_snet_i_Ws : m.g_snet_u_Ws;
    const int N = sd[k].B * sd[k].S;
    { ProfScope ps_("umpr_snet_sentiment_bwd", sk);
      UMPR_TRY(umpr_snet_sentiment_bwd(sb[k].self_atte, sb[k].wsum, dins + (size_t)(2 * k + 1) * B * Dm, nullptr, B, sd[k].S, sb[k].d_sa, sb[k].d_wsum, sk)); }
    { ProfScope ps_("umpr_snet_bwd_tc", sk);
      UMPR_TRY(umpr_snet_bwd_tc(sb[k].out_r, sd[k].snet_table, sd[k].snet_tiles, sb[k].d_sa, Ms, Ws, N, sd[k].L, sb[k].dx_s, gMs, gWs, n_ctas, sk)); }
    const long n = (long)N * sd[k].L;        // d(word_soft)[n][l] = d(sum_l word_soft)[n]  (model.py:79)
    expand_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sk>>>(sb[k].d_wsum, N, sd[k].L, sb[k].d_soft);
    UMPR_TRY(check_launch("step expand"));
  }
  if (par) { cudaEventRecord(ev[10], st4); cudaStreamWaitEvent(st, ev[10], 0); }
  STEP_CALL("umpr_coattn_bwd", umpr_coattn_bwd(gu, gi, giM, soft, soft + BP, soft + 2 * BP, soft + 3 * BP, arg, arg + BP, sb[0].d_soft, sb[1].d_soft, dins, dins + 2 * (size_t)B * Dm,
                           B, P, cst_u, sd[0].S, sd[0].L, cst_i, sd[1].S, sd[1].L, sb[0].dx_s, sb[1].dx_s, sb[0].dx_r, sb[1].dx_r, dgiM, stream));
  // dgi += dgiM · M^T (feeds the GRU backward) ;  dM = gi^T · dgiM (a parameter gradient only: fourth stream)
  if (par) { cudaEventRecord(ev[9], st); cudaStreamWaitEvent(st4, ev[9], 0); }
  if (BP >= 4096) {
    ProfScope ps_("umpr_tc_gemm_tn", st4);
    UMPR_TRY(umpr_tc_gemm_tn(gi, Dm, dgiM, Dm, m.g_M, Dm, Dm, Dm, (long)BP, n_ctas, st4));
  } else {
    int splits = (int)(BP / 256);
    splits = splits < 1 ? 1 : (splits > n_ctas ? n_ctas : splits);
    ProfScope ps_("umpr_sgemm", st4);
    UMPR_TRY(umpr_sgemm(gi, 1, Dm, dgiM, Dm, 1, m.g_M, Dm, Dm, Dm, (int)BP, splits, 1, nullptr, 0, st4));
  }
  if (par) cudaEventRecord(ev[11], st4);
  if (BP >= 1024)
    STEP_CALL("umpr_tc_gemm_ws", umpr_tc_gemm_ws(dgiM, Dm, m.M, Dm, sb[1].dx_r, Dm, (int)BP, Dm, Dm, 1, nullptr, 0, 0, sd[1].snet_table, sd[1].snet_tiles, sd[1].L, n_ctas, stream));
  else
    STEP_CALL("umpr_tc_gemm_nt", umpr_tc_gemm_nt(dgiM, Dm, m.M, Dm, sb[1].dx_r, Dm, (int)BP, Dm, Dm, 1, nullptr, 0, 0, stream));
  cudaStream_t& comm_st = ds.comm;
  cudaEvent_t* cev = ds.cev;
  const bool overlap = g_ov.comm != nullptr;
  if (overlap) {
    if (!comm_st) {
      if (cudaStreamCreateWithFlags(&comm_st, cudaStreamNonBlocking) != cudaSuccess) { comm_st = nullptr; return fail_arg("step: cannot create the communication stream"); }
      for (int i = 0; i < 3; ++i) cudaEventCreateWithFlags(&cev[i], cudaEventDisableTiming);
    }
    // early bucket: every gradient but R-Net's GRU tensors has been accumulated (on this stream, and on the C-Net branch's)
    cudaEventRecord(cev[0], st);
    cudaStreamWaitEvent(comm_st, cev[0], 0);
    if (two) { cudaEventRecord(ev[3], stc); cudaStreamWaitEvent(comm_st, ev[3], 0); }
    if (par) cudaStreamWaitEvent(comm_st, ev[11], 0);      // dM
    UMPR_TRY(umpr_allreduce(g_ov.comm, g_ov.bucket, g_ov.n_early, comm_st));
  }
  {
    umpr_gru_bwd_seg segs[2];
    for (int k = 0; k < 2; ++k)
      segs[k] = umpr_gru_bwd_seg{sb[k].dx_r, nullptr, sb[k].xq, sb[k].hq_r, sd[k].plan, sd[k].n_tiles, sd[k].n_slabs, sd[k].B * sd[k].S, sd[k].L};
    STEP_CALL("umpr_gru_bwd_tc", umpr_gru_bwd_tc(segs, 2, m.rnet_gru, m.g_rnet_gru, E, zero_img, sched_r, nq_r, stream));
  }
  if (overlap) {
    cudaEventRecord(cev[1], st);             // late bucket: R-Net's GRU gradients (and the shard count behind them)
    cudaStreamWaitEvent(comm_st, cev[1], 0);
    UMPR_TRY(umpr_allreduce(g_ov.comm, g_ov.bucket + g_ov.n_early, g_ov.n_total - g_ov.n_early, comm_st));
    cudaEventRecord(cev[2], comm_st);
    cudaStreamWaitEvent(st, cev[2], 0);      // whatever follows on the caller's stream (umpr_adam_step) sees the reduced bucket
  } else {
    if (two) { cudaEventRecord(ev[3], stc); cudaStreamWaitEvent(st, ev[3], 0); }      // join: every gradient is in the bucket
    if (par) cudaStreamWaitEvent(st, ev[11], 0);
  }
  return 0;
}

static int check_args(const umpr_step_model* m, const umpr_step_side* sd, int train) {
  if (!m || !sd) return fail_arg("step: NULL model or sides");
  const int n_sides = m->review_net_only ? 2 : 3;
  for (int k = 0; k < n_sides; ++k) {
    const umpr_step_side& s = sd[k];
    if (s.B < 1 || s.S < 1 || s.L < 1 || s.n_tiles < 1 || s.n_slabs < 1) return fail_arg("step: side %d is empty", k);
    if (s.L > 128 || (!m->review_net_only && s.L + 2 > 128)) return fail_arg("step: padded sentence length %d exceeds this path's limit", s.L);
  }
  if (m->E < 1 || m->E >= KP) return fail_arg("step: embedding width E=%d must be in [1,%d)", m->E, KP);
  if (!m->review_net_only && (m->KC < 1 || m->KC > 128 || m->ksize != 3 || m->V < 1 || m->V > 128)) return fail_arg("step: KC=%d ksize=%d V=%d outside this path's limits", m->KC, m->ksize, m->V);
  (void)train;
  return 0;
}
}  // namespace umpr

using namespace umpr;

extern "C" int umpr_step_workspace_bytes(const umpr_step_model* model, const umpr_step_side* sides, int train, long long* bytes) {
  if (!bytes) return fail_arg("step_workspace_bytes: NULL output");
  UMPR_TRY(check_args(model, sides, train));
  Arena a{nullptr, 0, 0};
  UMPR_TRY(run_step(*model, sides, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, a, nullptr, nullptr, train, 148, nullptr));
  *bytes = (long long)((a.off + 255) & ~size_t(255));
  return 0;
}

extern "C" int umpr_step(const umpr_step_model* model, const umpr_step_side* sides, const float* photos, const float* labels,
                         const int32_t* sched_r, int nq_r, const int32_t* sched_c, int nq_c, const void* zero_img, void* workspace,
                         long long workspace_bytes, float* pred, float* loss, int train, void* stream) {
  UMPR_TRY(check_args(model, sides, train));
  if (!workspace || !labels || !sched_r || !pred || !loss) return fail_arg("step: NULL workspace / labels / schedule / outputs");
  if (!model->review_net_only && (!photos || !sched_c)) return fail_arg("step: the full model needs photo features and the C-Net schedule");
  if (train && !zero_img) return fail_arg("step: zero_img (32 KB of zeros) is required for training");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail_arg("step: workspace must be 256-byte aligned");
  static int n_ctas = 0;
  if (!n_ctas) {
    const char* e = getenv("UMPR_STEP_SINGLE_STREAM");
    if (!g_streams_set) g_single_stream = e && e[0] == '1';
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_ctas, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_ctas < 1) n_ctas = 148;
  }
  Arena a{reinterpret_cast<unsigned char*>(workspace), 0, (size_t)workspace_bytes};
  return run_step(*model, sides, photos, labels, sched_r, nq_r, sched_c, nq_c, zero_img, a, pred, loss, train, n_ctas, stream);
}

// per-entry-point timing of the following umpr_step calls: every entry point (only == NULL) or just the named one
// serialize = 1: the whole step on the caller's stream (clean per-kernel timing); 0: branches on the library's side streams
extern "C" int umpr_step_streams(int serialize) {
  g_single_stream = serialize != 0;
  g_streams_set = true;
  return 0;
}

extern "C" int umpr_step_profile_begin(const char* only) {
  g_prof.on = true;
  g_prof.only = only ? only : "";
  g_prof.recs.clear();
  g_prof.used = 0;
  return 0;
}
// stops the timing, synchronises, and returns per entry point (aggregated by name, first-launch order): names[i] (48 bytes each,
// NUL-terminated), total milliseconds, number of calls; *n_out entries (at most max_entries)
extern "C" int umpr_step_profile_end(int max_entries, char* names, float* ms, int* calls, int* n_out) {
  g_prof.on = false;
  if (!names || !ms || !calls || !n_out) return fail_arg("step_profile_end: NULL output");
  int n = 0;
  for (const ProfRec& r : g_prof.recs) {
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e != cudaSuccess) { set_error("step_profile_end: %s", cudaGetErrorString(e)); return (int)e; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    int k = 0;
    for (; k < n; ++k) if (!strcmp(names + 48 * k, r.name)) break;
    if (k == n) {
      if (n == max_entries) continue;
      strncpy(names + 48 * k, r.name, 47);
      names[48 * k + 47] = 0;
      ms[k] = 0.f; calls[k] = 0;
      ++n;
    }
    ms[k] += t; calls[k] += 1;
  }
  *n_out = n;
  g_prof.recs.clear();
  g_prof.used = 0;
  return 0;
}

// Overlap the gradient exchange with the backward of the following umpr_step(train) calls: comm = umpr_comm_init handle (NULL switches
// the overlap off), bucket = the flat gradient buffer the g_* pointers of the model point into, laid out [early (n_early floats) | late]
// with ONLY R-Net's GRU gradients (and trailing extras such as a shard counter) in the late part.  Both all-reduces run on a stream
// of the library; the caller's stream waits for them at the end of umpr_step.
extern "C" int umpr_step_comm(void* comm, float* bucket, long n_early, long n_total) {
  if (comm && (!bucket || n_early < 0 || n_total < n_early)) return fail_arg("step_comm: bucket=%p n_early=%ld n_total=%ld", (void*)bucket, n_early, n_total);
  g_ov.comm = comm; g_ov.bucket = bucket; g_ov.n_early = n_early; g_ov.n_total = n_total;
  return 0;
}
