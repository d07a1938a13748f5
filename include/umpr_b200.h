/* umpr_b200 — C-ABI of the B200-native UMPR review-network hot path.
 *
 * This is the drop-in boundary: plain C, raw device pointers, explicit sizes, the CUDA stream as
 * void*.  No torch types.  Every function cites the reference code it replaces (paths relative to the reference
 * repository iamwinter/UMPR).  The reference ships no native code: each entry point below replaces the stock
 * PyTorch/ATen/cuDNN/cuBLAS operators that the cited Python lines dispatch to.
 *
 * Conventions
 *   - return value: 0 ok; UMPR_ERR_ARG (<0) bad argument/shape (see umpr_last_error()); >0 a cudaError_t.
 *   - all tensors are dense row-major fp32 unless stated; token ids and nothing else are int64; plans are int32.
 *   - launches are asynchronous on `stream`; no hidden synchronisation; the library owns no device memory.
 *   - fixed hot-path sizes: gru_size 64, self_atte_size 64, conv kernel_size 3, kernel_count <= 128 (config.py:34-37).
 *   - "d_*" / "*grad*" outputs marked (+=) are accumulated with atomics and must be initialised by the caller.
 *
 * Pack plan (int32 device buffer, built on the host from the reference's own torch.sort call, model.py:18):
 *   [seq_of (Rp) | row_of (Rp) | len_of (Rp) | tile_off (n_tiles+1) | slab_tile (n_slabs)],  Rp = n_tiles*R
 *   job k (descending length): reads input sequence seq_of[k], fills ImprovedRnn output row row_of[k] (model.py:21),
 *   has length len_of[k] (0 = padding job).  Tile j = jobs [j*R, (j+1)*R) owns slabs tile_off[j] .. +len_of[j*R].
 */
#ifndef UMPR_B200_H
#define UMPR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMPR_B200_VERSION 100
#define UMPR_ERR_ARG (-1)

int umpr_version(void);
/* Scratch bytes of the entry points that take a caller-owned workspace (PyTorch owns every buffer):
 *   "coattn_fwd_tc": a = B, b = P     "cnet_conv_fwd_tc": a = cap (>= N*KC/8)      "cnet_conv_bwd_dx": a = kernel_count
 *   "cnet_conv_bwd_dx_tc": no size argument */
int umpr_workspace_bytes(const char* entry, long a, long b, long long* bytes);

/* ---- text matching (model.py:166-168): y = tanh(linear_u([atte_u | senti_u]) + linear_i([atte_i | senti_i])), Wu / Wi (128,256),
 * bias-free; the backward takes dpre = dy * (1 - y^2) (umpr_tanh_bwd) and returns the four input gradients (weight gradients:
 * umpr_tc_gemm_tn / umpr_sgemm reductions over the batch) ---- */
int umpr_text_match_fwd(const float* atte_u, const float* senti_u, const float* atte_i, const float* senti_i, const float* Wu,
                        const float* Wi, int B, float* y /*(B,128)*/, void* stream);
int umpr_text_match_bwd(const float* dpre, const float* Wu, const float* Wi, int B, float* d_atte_u, float* d_senti_u, float* d_atte_i,
                        float* d_senti_i, void* stream);
int umpr_text_match_wgrad(const float* dpre, const float* atte_u, const float* senti_u, const float* atte_i, const float* senti_i, int B,
                          float* dWu /*(128,256) +=*/, float* dWi /*(+=)*/, void* stream);

/* ---- R-Net pre-training head (pretrain/pretrain_rnet.py:148-168): sigmoid(Linear(256 -> 1)([att_u | att_i])) + BCELoss (mean) ---- */
int umpr_bce_head_fwd(const float* att_u /*(B,128)*/, const float* att_i, const float* w /*(1,256)*/, const float* bias /*(1)*/,
                      const float* target /*(B)*/, int B, float* result /*(B)*/, float* loss /*scalar, zero-initialised*/, void* stream);
int umpr_bce_head_bwd(const float* att_u, const float* att_i, const float* w, const float* result, const float* target,
                      const float* d_loss /*scalar or NULL*/, const float* d_result /*(B) or NULL*/, int B, float* d_att_u, float* d_att_i,
                      float* d_w /*(+=)*/, float* d_b /*(+=)*/, void* stream);

/* ---- data-parallel gradient exchange (replaces nn.DataParallel, main.py:81-82): one process per GPU, one in-place sum all-reduce
 * of the flat fp32 gradient bucket per step.  NCCL is resolved at run time from the libnccl.so.2 already in the process. ---- */
int umpr_comm_unique_id(void* id128 /* 128 bytes, created on rank 0 and handed to every rank by the host */);
int umpr_comm_init(int rank, int world, const void* id128, void** comm);
int umpr_allreduce(void* comm, float* flat, long n_floats, void* stream);
int umpr_comm_destroy(void* comm);

const char* umpr_last_error(void);              /* thread-local message of the last failing call */
int umpr_sm_count(int device, int* out);

/* ---- ImprovedRnn: src/model.py:12-21 (pack_padded_sequence -> nn.GRU -> pad_packed_sequence -> 2nd un-sort),
 *      embedding gather src/model.py:262-264.  `w` / `dw` are 8 pointers in nn.GRU order:
 *      weight_ih_l0 (192,E), weight_hh_l0 (192,64), bias_ih_l0, bias_hh_l0, then the *_reverse four. ---- */
/* xp[n_slabs][R][64] <- embedding rows (table[ids], or dense (Nseq,L,E) when dense != NULL), a 1.0 bias column, zeros */
int umpr_gather_pack(const float* table, const int64_t* ids, const float* dense, const int32_t* plan, int n_tiles,
                     int n_slabs, int R, int L, int E, float* xp, void* stream);
/* G[n_slabs][2][R][192] = xp · [W_ih | b_ih (+ b_hh for r,z)]^T for both directions */
int umpr_gru_inproj(const float* xp, const float* const* w, int n_slabs, int R, int E, float* G, void* stream);
/* out (N,L,128) fully written (zeros for t >= length, total_length = L, model.py:17,20); hn (2,N,64) or NULL;
 * sv[n_slabs][2][R][256] saved gates for backward, or NULL for inference (evaluate.py:8-11) */
int umpr_gru_recurrence_fwd(const float* G, const float* const* w, const int32_t* plan, int n_tiles, int n_slabs, int R,
                            int N, int L, float* out, float* hn, float* sv, void* stream);
/* dG[n_slabs][2][R][256] = d(gate pre-activations) [dr, dz, dn, dn*r]; d_hn may be NULL */
int umpr_gru_recurrence_bwd(const float* d_out, const float* d_hn, const float* out, const float* sv, const float* const* w,
                            const int32_t* plan, int n_tiles, int n_slabs, int R, int N, int L, float* dG, void* stream);
/* dw[8] (+=): gradients of the 8 GRU tensors.  No input gradient: the embedding is frozen (model.py:237). */
int umpr_gru_wgrad(const float* dG, const float* xp, const float* out, const int32_t* plan, int n_tiles, int n_slabs, int R,
                   int L, int E, float* const* dw, int n_ctas, void* stream);

/* tensor-core form of umpr_gru_wgrad (tcgen05 with MN-major operands, TMEM accumulation over each CTA's slot range) */
int umpr_gru_wgrad_tc(const float* dG, const float* xp, const float* out, const int32_t* plan, int n_tiles, int n_slabs, int R,
                      int L, int E, float* const* dw, int n_ctas, void* stream);

/* ---- fused tensor-core ImprovedRnn forward: input projection + recurrence of model.py:19 in one persistent tcgen05 kernel
 *      (W_ih, W_hh resident in shared memory, two 128-sequence tiles per CTA ping-ponging through TMEM, rows masked by their
 *      own length, no G round trip).  Up to 3 calls that share the GRU weights (user+item of model.py:45-46; ui+user+item of
 *      model.py:182-184) run as segments of one launch.  Plans must use R = 128.  sched (int32, device):
 *      [q_off (n_queues+1) | q_tile (sum of n_tiles)]: queue q lists global tile ids (segment tile bases are cumulative
 *      n_tiles in segment order), CTA c walks queues 2c and 2c+1 (plan.py builds it longest-first). ---- */
typedef struct umpr_gru_seg {
  const void* xq;         /* [n_slabs][hi|lo][128][64 bf16] packed token images (umpr_gather_pack_tc) */
  const int32_t* plan;    /* pack plan with R = 128 */
  float* out;             /* (N,L,128), fully written */
  float* hn;              /* (2,N,64) or NULL */
  void* hq;               /* [n_slabs][2][hi|lo][128][64 bf16] h_t operand images: with xq ALL the backward kernel reads (it recomputes the
                             gates); NULL = inference, nothing is kept */
  int32_t n_tiles, n_slabs, N, L;
} umpr_gru_seg;
/* xq[n_slabs][hi|lo][128][64 bf16]: gathered + packed tokens as SWIZZLE_128B bf16 hi/lo operand images (E values, 1.0, zeros) */
int umpr_gather_pack_tc(const float* table, const int64_t* ids, const float* dense, const int32_t* plan, int n_tiles, int n_slabs,
                        int L, int E, void* xq, void* stream);
int umpr_gru_fwd_tc(const umpr_gru_seg* segs /*host array*/, int n_seg, const float* const* w, int E, const int32_t* sched,
                    int n_queues, void* stream);

/* backward of umpr_gru_fwd_tc (same tiles, queues and segments, reverse time), recurrence AND weight gradients in one kernel:
 * the gates are recomputed from xq[t] and hq[t-1] with the forward's own MMAs; the gate gradients go into a shared-memory tile
 * that is the A operand of the carry product [dr, dz, dn*r]·W_hh and the B operand of the weight-gradient product
 * [xq | hq]^T·[dr, dz, dn*r, dn], which accumulates in tensor memory over each CTA's queue - neither gates nor gate gradients
 * ever reach HBM.  xq, hq: what the forward launch read / wrote.  dw[8] (+=): gradients of the 8 GRU tensors (nn.GRU order).
 * zero_img: 32 KB of zeros (h_{t-1} of a sequence's first step). */
typedef struct umpr_gru_bwd_seg {
  const float* d_out;     /* (N,L,128) gradient of the ImprovedRnn result */
  const float* d_hn;      /* (2,N,64) or NULL */
  const void* xq;         /* packed token images (forward input) */
  const void* hq;         /* h_t operand images written by the forward */
  const int32_t* plan;
  int32_t n_tiles, n_slabs, N, L;
} umpr_gru_bwd_seg;
int umpr_gru_bwd_tc(const umpr_gru_bwd_seg* segs /*host array*/, int n_seg, const float* const* w, float* const* dw, int E,
                    const void* zero_img, const int32_t* sched, int n_queues, void* stream);
/* ---- generic strided fp32 GEMM: gi·M (model.py:50), text matching (model.py:168) and their gradients ----
 * C[m][n] = act(accumulate*C + sum_k A[m*ars+k*acs] * B[k*brs+n*bcs] + bias[n]); act 0 none, 1 tanh, 2 relu, 3 sigmoid.
 * splits > 1: split-K with atomic accumulation into C (C must be initialised; bias/act not allowed). */
int umpr_sgemm(const float* A, long ars, long acs, const float* B, long brs, long bcs, float* C, long ldc, int M, int N, int K,
               int splits, int accumulate, const float* bias, int act, void* stream);

/* ---- tcgen05 (5th-gen tensor core) variants: fp32 operands split into 3xBF16 on the fly, fp32 accumulation in TMEM ----
 * C[m][n] = act(accumulate*C + sum_k A[m*lda+k] * B[n*ldb+k] + bias[n])  ("NT": both operands K-contiguous); act 0|1 tanh|2 relu.
 * Pointers 16-byte aligned, leading dimensions multiples of 4. */
int umpr_tc_gemm_nt(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K, int accumulate,
                    const float* bias, int act, int b_kn /* B stored [K][N] instead of [N][K] */, void* stream);
/* same contract as umpr_gru_inproj, on the tensor cores: one GEMM over every packed token (model.py:19, input half) */
int umpr_gru_inproj_tc(const float* xp, const float* const* w, int n_slabs, int R, int E, float* G, int n_ctas, void* stream);
/* persistent weight-stationary form of umpr_tc_gemm_nt for N <= 128, K <= 128 (B stays in shared memory, A streams).
 * table (optional) = [tile_sent_off (n_row_tiles+1) | cstart (M/L+1)] for an A whose rows are sentences of L positions with
 * exactly-zero rows at and beyond each sentence's length: only the valid rows are read and written, the others left untouched. */
int umpr_tc_gemm_ws(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K, int accumulate,
                    const float* bias, int act, int b_kn, const int32_t* table, int n_row_tiles, int L, int n_ctas, void* stream);

/* reduction GEMM over a huge K with both operands row-major in k ("TN"): C[m][n] (+=) sum_k A[k*lda+m] * B[k*ldb+n], M, N <= 128.
 * dM = gi^T · dgiM of the co-attention backward (model.py:50) runs here.  C must be initialised (atomic accumulation). */
int umpr_tc_gemm_tn(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, long K, int n_ctas, void* stream);

/* ---- RNet co-attention: src/model.py:50-55.  gu, gi, giM (=gi·M): (B,P,128).  The (P,P) affinity matrix is never
 *      materialised.  rowkey/colkey: (B,P) uint64 scratch, colkey zero-initialised.  t_*: tanh of the row/col maxima,
 *      arg_*: their positions (saved for backward). ---- */
int umpr_coattn_fwd(const float* gu, const float* gi, const float* giM, int B, int P, unsigned long long* rowkey,
                    unsigned long long* colkey, float* soft_u, float* soft_i, float* t_u, float* t_i, int32_t* arg_u,
                    int32_t* arg_i, float* atte_u, float* atte_i, void* stream);
/* tensor-core form of umpr_coattn_fwd (at most 2560 valid positions per sample and side): operands pre-split into bf16 hi|lo images; one CTA per sample issues two tcgen05
 * products per tile pair (S and S^T, so row and column maxima are per-thread scans) in two passes - maxima, then every value
 * within the 3xBF16 error bound of its final maximum - and the surviving near-ties are re-scored with exact fp32 dot products
 * (the arg-max routes the gradient).  scratch: umpr_workspace_bytes("coattn_fwd_tc", B, P), 16-byte aligned.
 * cst_u / cst_i (both or neither): exclusive prefix sums (int32, device, B*S_x+1 entries) of the sentence lengths of the user / item
 * side, S_x sentences of L_x positions per sample (S_x*L_x == P), for inputs produced by ImprovedRnn (rows at or beyond a
 * sentence's length are exactly zero, model.py:20): only the valid rows are multiplied, their exact 0 enters each maximum
 * analytically.  NULL: all rows are treated as valid. */
int umpr_coattn_fwd_tc(const float* gu, const float* gi, const float* giM, int B, int P, const int32_t* cst_u, int S_u, int L_u,
                       const int32_t* cst_i, int S_i, int L_i, int pv_max /* most valid rows of any sample and side; the limit of 2560
                       applies to it (to P without tables) */, void* scratch, float* soft_u, float* soft_i, float* t_u, float* t_i,
                       int32_t* arg_u, int32_t* arg_i, float* atte_u, float* atte_i, void* stream);
/* dgu, dgi (without the dgiM·M^T term), dgiM: (B,P,128).  d_* inputs may be NULL.  cst_* as in umpr_coattn_fwd_tc: with them, rows
 * at or beyond a sentence's length are neither read nor written (dgiM alone gets explicit zeros there); NULL: all rows written.
 * add_u / add_i (optional, (B,P,128)): gradients of gu / gi from their other consumer (S-Net), added into dgu / dgi on the way out. */
int umpr_coattn_bwd(const float* gu, const float* gi, const float* giM, const float* soft_u, const float* soft_i, const float* t_u,
                    const float* t_i, const int32_t* arg_u, const int32_t* arg_i, const float* d_soft_u, const float* d_soft_i,
                    const float* d_atte_u, const float* d_atte_i, int B, int P, const int32_t* cst_u, int S_u, int L_u,
                    const int32_t* cst_i, int S_i, int L_i, const float* add_u, const float* add_i, float* dgu, float* dgi, float* dgiM,
                    void* stream);

/* ---- SNet: src/model.py:71-81.  x = gru_repr viewed (N, L, 128), N = B*S. ---- */
int umpr_snet_fwd(const float* x, const float* Ms, const float* Ws, int N, int L, float* self_atte /*(N,128)*/,
                  float* soft /*(N,L) or NULL*/, float* th /*(N,L,64) or NULL*/, int n_ctas, void* stream);
int umpr_snet_sentiment_fwd(const float* self_atte, const float* word_soft /*(B,S,Wd)*/, int B, int S, int Wd, float* wsum /*(B,S)*/,
                            float* sentiment /*(B,128)*/, void* stream);
int umpr_snet_sentiment_bwd(const float* self_atte, const float* wsum, const float* d_sentiment, const float* d_self_atte_up, int B,
                            int S, float* d_self_atte, float* d_wsum, void* stream);
int umpr_snet_bwd(const float* x, const float* th, const float* soft, const float* d_self_atte, const float* Ms, const float* Ws, int N,
                  int L, float* dx /*(N,L,128) written*/, float* dMs /*(+=)*/, float* dWs /*(+=)*/, int n_ctas, void* stream);

/* Tensor-core S-Net for inputs produced by ImprovedRnn (rows at or beyond a sentence's length are exactly zero, model.py:20): only the
 * valid rows are multiplied.  table = [tile_sent_off (n_tiles+1) | cstart (N+1)] (int32, device): sentences tile_sent_off[k] ..
 * tile_sent_off[k+1]-1 form tile k (<= 128 valid rows), cstart = exclusive prefix sum of the per-sentence lengths.  The backward
 * recomputes the scores from x (nothing else is saved) and writes dx only for the valid rows. */
int umpr_snet_fwd_tc(const float* x, const int32_t* table, int n_tiles, const float* Ms, const float* Ws, int N, int L,
                     float* self_atte /*(N,128)*/, int n_ctas, void* stream);
int umpr_snet_bwd_tc(const float* x, const int32_t* table, int n_tiles, const float* d_self_atte, const float* Ms, const float* Ws, int N,
                     int L, float* dx /*(N,L,128): valid rows written*/, float* dMs /*(+=)*/, float* dWs /*(+=)*/, int n_ctas, void* stream);

/* ---- CNet tail: src/model.py:118-125 ---- */
int umpr_cnet_prep(const float* conv_w /*(KC,128,3)*/, int KC, int ksize, float* wt /*(384,128)*/, void* stream);
int umpr_cnet_conv_fwd(const float* x, const float* wt, const float* conv_b, int N, int L, int KC, float* cfeat /*(N,KC)*/,
                       int32_t* cidx /*(N,KC) arg-max position, -1 if clipped by ReLU*/, int n_ctas, void* stream);
/* tensor-core form of umpr_cnet_prep + umpr_cnet_conv_fwd (implicit GEMM on tcgen05, weights streamed by bulk copies); maxima
 * that are near-tied (or next to the ReLU threshold) are re-scored in exact fp32 because the arg-max routes the gradient.
 * scratch: 197632 + 16*cap bytes, 16-byte aligned; it holds one 2-byte re-scoring record per (sentence, filter): cap >= ceil(N*KC/8).
 * table (optional, int32 device) = [tile_sent_off (n_tiles+1) | cstart (N+1)], cstart = exclusive prefix sum of (len+2) per sentence,
 * for inputs produced by ImprovedRnn (rows at or beyond a sentence's length exactly zero): only the valid rows are laid out and
 * multiplied, the all-zero windows enter the max as the bias.  NULL: every sentence is processed at its full length L. */
int umpr_cnet_conv_fwd_tc(const float* x, const float* conv_w, const float* conv_b, int N, int L, int KC, int ksize,
                          const int32_t* table, int n_tiles, void* scratch, int cap, float* cfeat, int32_t* cidx, int n_ctas,
                          void* stream);
int umpr_cnet_head_fwd(const float* cfeat, const float* lin_w, const float* lin_b, float threshold, int B, int S, int V, int KC,
                       float* view_p /*(B,S,V)*/, float* final_repr /*(B,V)*/, void* stream);
int umpr_cnet_head_bwd(const float* cfeat, const int32_t* cidx, const float* view_p, const float* lin_w, const float* d_view_p,
                       const float* d_final, int B, int S, int V, int KC, float* dcfeat /*(N,KC) written*/, float* d_lin_w /*(+=)*/,
                       float* d_lin_b /*(+=)*/, float* d_conv_b /*(+=)*/, void* stream);
/* Backward of conv + ReLU + max (model.py:118-120): sparse, one arg-max position per (sentence, filter).
 * cst (optional): exclusive prefix sum (N+1) of the sentence lengths for an x produced by ImprovedRnn (rows at or beyond a sentence's
 * length exactly zero): those rows are not staged and their dx rows are not written.  NULL: all rows. */
int umpr_cnet_conv_bwd_dx(const float* dcfeat, const int32_t* cidx, const float* conv_w, int N, int L, int KC, const int32_t* cst,
                          float* wt_scratch /*KC*3*128 floats (tap-major weight copy), or NULL for the scatter kernel*/,
                          float* dx /*(N,L,128): rows below each length written*/, int n_ctas, void* stream);
int umpr_cnet_conv_bwd_dw(const float* x, const float* dcfeat, const int32_t* cidx, int N, int L, int KC, const int32_t* cst,
                          float* d_conv_w /*(+=)*/, int n_ctas, void* stream);
/* the same weight gradient on tcgen05: per tap, dW_j^T = x^T . G_j with the gradients scattered into a one-hot tile G_j (3xBF16);
 * table = the convolution's tile table of umpr_cnet_conv_fwd_tc (required) */
int umpr_cnet_conv_bwd_dw_tc(const float* x, const float* dcfeat, const int32_t* cidx, int N, int L, int KC, const int32_t* table,
                             int n_tiles, float* d_conv_w /*(+=)*/, int n_ctas, void* stream);
/* the input gradient on tcgen05: dX = sum_j G_j . W_j with the same one-hot tiles, the tap weights streamed by TMA bulk copies;
 * scratch: umpr_workspace_bytes("cnet_conv_bwd_dx_tc") = 196608 bytes, 16-byte aligned; rows below each length written */
int umpr_cnet_conv_bwd_dx_tc(const float* dcfeat, const int32_t* cidx, const float* conv_w, int N, int L, int KC, const int32_t* table,
                             int n_tiles, void* scratch, float* dx, int n_ctas, void* stream);

/* ---- ControlNet tail: SSNet (model.py:142-143), Eq.18 (model.py:188, eps 1e-4 in code), gates (model.py:189-197) ---- */
/* standalone SSNet (model.py:129-143): y[r] = sigmoid(x[r,:128] . w + b); backward adds into dw (128) / db (1), dx may be NULL */
int umpr_ssnet_fwd(const float* x, const float* w, const float* b, long rows, float* y, void* stream);
int umpr_ssnet_bwd(const float* x, const float* w, const float* y, const float* dy, long rows, float* dx, float* dw, float* db,
                   void* stream);
int umpr_control_tail_fwd(const float* s, const float* view_p, const float* c_out, const float* ss_w, const float* ss_b, float eps,
                          int B, int Su, int V, float* senti, float* score, float* prefer_pos, float* prefer_neg, void* stream);
int umpr_control_tail_bwd(const float* s, const float* view_p, const float* c_out, const float* ss_w, const float* senti,
                          const float* score, const float* d_prefer_pos, const float* d_prefer_neg, float eps, int B, int Su, int V,
                          float* d_s, float* d_view_p, float* d_c_out, float* d_ss_w /*(+=)*/, float* d_ss_b /*(+=)*/, void* stream);

/* ---- VisualNet tail: src/model.py:219-228 on VGG16 *features* (B,V,Pc,F); the backbone (model.py:204-207,217) is out of scope ---- */
int umpr_visual_fwd(const float* feat, const float* pos_v_emb, const float* neg_v_emb, const float* w, const float* bias,
                    const float* c_u, const float* c_i, int B, int V, int Pc, int F, float* emb /*(2,V)*/, float* img_emb,
                    float* pos_match, float* neg_match, float* final_pos, float* final_neg, void* stream);
int umpr_visual_bwd(const float* feat, const float* pos_v_emb, const float* neg_v_emb, const float* w, const float* emb,
                    const float* img_emb, const float* pos_match, const float* neg_match, const float* c_u, const float* c_i,
                    const float* d_pos_match, const float* d_neg_match, const float* d_final_pos, const float* d_final_neg, int B, int V,
                    int Pc, int F, float* scratch /*3*B*V*/, float* d_c_u, float* d_c_i, float* d_pos_v_emb, float* d_neg_v_emb,
                    float* d_w /*(+=)*/, float* d_b /*(+=)*/, void* stream);

/* ---- fusion Linear+ReLU (model.py:242-245,252-255,268,274) and losses (model.py:269,275-277) ---- */
int umpr_fusion_fwd(const float* repr, const float* final_pos, const float* final_neg, const float* w, const float* bias, int B, int V,
                    float* pred, void* stream);
int umpr_fusion_bwd(const float* repr, const float* final_pos, const float* final_neg, const float* w, const float* pred,
                    const float* d_pred, int B, int V, float* d_repr, float* d_final_pos, float* d_final_neg, float* d_w /*(+=)*/,
                    float* d_b /*(+=)*/, void* stream);
int umpr_loss_fwd(const float* pred, const float* labels, const float* prefer_pos, const float* prefer_neg, const float* pos_match,
                  const float* neg_match, int B, int V, float loss_v_rate, float* loss /*zero-initialised scalar*/, void* stream);
int umpr_loss_bwd(const float* pred, const float* labels, const float* prefer_pos, const float* prefer_neg, const float* pos_match,
                  const float* neg_match, const float* d_loss, int B, int V, float loss_v_rate, float* d_pred, float* d_prefer_pos,
                  float* d_prefer_neg, float* d_pos_match, float* d_neg_match, void* stream);
int umpr_tanh_bwd(const float* y, const float* dy, long n, float* dx, void* stream);

/* ---- train step tail: main.py:22-26,37 Adam with L2 on non-bias tensors, fused over one flat parameter buffer ---- */
int umpr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* weight_decay /*per element*/,
                   long n, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                   const float* shard_count /* device scalar or NULL: when given, the gradient scale is 1 / max(1, *shard_count) -
                                               the number of replicas that received a chunk (DataParallel's loss.mean(), main.py:34) */,
                   void* stream);

/* ---- host-side integer bookkeeping of the packing (model.py:18-21), no CUDA: pack plan, valid-row tile tables and tile schedules from
 * the reference's own sort result (sorted_idx / sorted_len = torch.sort(lengths.cpu(), descending=True)); see umpr_b200/plan.py ---- */
int umpr_plan_build(const int64_t* sorted_idx, const int64_t* sorted_len, long n, int R, int32_t* plan, long plan_ints, int64_t* tokens_out);
int umpr_plan_table(const int64_t* sorted_idx, const int64_t* lengths, long n, int L, int extra /* 0: S-Net rows, 2: convolution guard rows */,
                    int32_t* table /* capacity 2*(n+1) */, int32_t* n_tiles_out);
int umpr_plan_schedule(const int64_t* const* tile_lens, const int32_t* n_tiles, int n_seg, int n_ctas, int32_t* sched, long sched_ints,
                       int32_t* n_queues_out);

/* ---- the data formats either side of the path ----
 * collate on device (dataset.py:122-131,163-171): expands ragged token lists - flat int32 tokens + the (n_sent + 1) exclusive prefix sum
 * of the per-slot token counts - into the padded (n_sent, L) int64 id tensor the reference's collate emits (pad_id beyond each
 * sentence, empty slots all pad). */
int umpr_collate_ids(const int32_t* flat_tokens, const int32_t* sent_off, long n_sent, int L, long pad_id, int64_t* ids, void* stream);
/* VGG16 feature cache (model.py:204-219, dataset.py:134-151): out[n] = table[idx[n]] (F floats per photo); idx < 0 or >= rows selects
 * missing_row (the features of the reference's zero image; < 0: zeros). */
int umpr_feature_gather(const float* table, const int32_t* idx, long n_photos, long rows, int F, long missing_row, float* out, void* stream);

/* ---- the whole step as ONE call (model.py:257-278 forward; with train != 0 also its backward, main.py:36): the same kernels as
 * above, issued from native host code out of one caller-owned workspace - no per-kernel host round trips.  Tensor-core path only:
 * every side needs a pack plan with R = 128 and its valid-row tables (plan.py: snet_table / cnet_table), L <= 126 for the full
 * model (<= 128 review-net only), at most 512 valid positions per sample.  Parameter gradients are accumulated (+=) into g_*:
 * the caller zeroes its bucket before the call and all-reduces / applies it afterwards (umpr_allreduce, umpr_adam_step). ---- */
typedef struct umpr_step_side {        /* one review side: user, item, user->item (dataset.py:173-182) */
  const int64_t* ids;                  /* (B, S, L) token ids */
  const int32_t* plan;                 /* pack plan, R = 128 */
  const int32_t* snet_table;           /* [tile_sent_off (snet_tiles+1) | cstart (B*S+1)] */
  const int32_t* cnet_table;           /* [tile_sent_off (cnet_tiles+1) | cstart (B*S+1)], full model only */
  int32_t snet_tiles, cnet_tiles, n_tiles, n_slabs, B, S, L, pv_max /* most valid positions of any sample */;
} umpr_step_side;
typedef struct umpr_step_model {
  int32_t review_net_only, V, Pc, F /* photo feature width */, KC, ksize, E, reserved;
  float threshold, loss_v_rate, eq18_eps, reserved_f;
  const float* table;                  /* embedding.weight (rows, E), frozen */
  /* parameters, reference state_dict order; GRU tensors in nn.GRU order (see umpr_gru_fwd_tc) */
  const float* rnet_gru[8]; const float* M; const float* snet_u_Ms; const float* snet_u_Ws; const float* snet_i_Ms; const float* snet_i_Ws;
  const float* lin_u; const float* lin_i; const float* fus_w; const float* fus_b;
  const float* cnet_gru[8]; const float* conv_w; const float* conv_b; const float* clin_w; const float* clin_b;
  const float* csnet_Ms; const float* csnet_Ws; const float* ss_w; const float* ss_b;
  const float* pos_e; const float* neg_e; const float* vis_w; const float* vis_b;
  /* their gradients (+=), same order; unused for train == 0 */
  float* g_rnet_gru[8]; float* g_M; float* g_snet_u_Ms; float* g_snet_u_Ws; float* g_snet_i_Ms; float* g_snet_i_Ws;
  float* g_lin_u; float* g_lin_i; float* g_fus_w; float* g_fus_b;
  float* g_cnet_gru[8]; float* g_conv_w; float* g_conv_b; float* g_clin_w; float* g_clin_b;
  float* g_csnet_Ms; float* g_csnet_Ws; float* g_ss_w; float* g_ss_b;
  float* g_pos_e; float* g_neg_e; float* g_vis_w; float* g_vis_b;
  /* optional taps (tests): the arg-max positions the kernels chose - co-attention (2, B, P), max-pool (B*S, KC) of ui / user / item */
  int32_t* routing_coattn; int32_t* routing_cnet[3];
} umpr_step_model;
int umpr_step_workspace_bytes(const umpr_step_model* model, const umpr_step_side* sides /* 2 (review net only) or 3: user, item, ui */,
                              int train, long long* bytes);
int umpr_step(const umpr_step_model* model, const umpr_step_side* sides, const float* photos /* (B,V,Pc,F) features or NULL */,
              const float* labels /* (B) */, const int32_t* sched_r, int nq_r /* tile schedule of the user+item GRU launch */,
              const int32_t* sched_c, int nq_c /* ... of the ui+user+item launch (full model) */, const void* zero_img /* 32 KB of zeros */,
              void* workspace /* 256-byte aligned */, long long workspace_bytes, float* pred /* (B) */, float* loss /* scalar */,
              int train, void* stream);

/* gradient exchange overlapped with the backward of the following umpr_step(train) calls (replaces DataParallel's reduce, main.py:81-82):
 * the bucket is [early | late] with only R-Net's GRU gradients (+ trailing extras) late; `early` is all-reduced on a library stream while
 * the last backward kernel runs, `late` after it; the caller's stream waits for both before umpr_step returns control to it. */
int umpr_step_comm(void* comm /* umpr_comm_init handle, NULL = off */, float* bucket, long n_early, long n_total);
/* optional timing of the entry points inside the following umpr_step calls (CUDA-event pairs on the launching stream): all of them
 * (only == NULL) or just the named one.  _end synchronises and returns the totals aggregated by entry-point name. */
int umpr_step_profile_begin(const char* only);
int umpr_step_profile_end(int max_entries, char* names /* max_entries x 48 bytes */, float* ms, int* calls, int* n_out);
/* umpr_step issues independent branches of the step (R-Net / C-Net, the item side of the C-Net tails, S-Net beside the co-attention)
 * on up to three library streams beside the caller's so that one branch fills the idle SMs of the other's kernel tails.
 * serialize = 1: everything on the caller's stream (per-kernel timing without co-scheduled kernels); 0: default. */
int umpr_step_streams(int serialize);

#ifdef __cplusplus
}
#endif
#endif
