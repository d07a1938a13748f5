// Host-side integer bookkeeping of ImprovedRnn's packing (reference src/model.py:18-21), in C: everything umpr_b200/plan.py derives
// from the reference's own torch.sort call - pack plan, the valid-row tile tables of S-Net / co-attention / GEMM rows and of the
// C-Net convolution, and the tile schedules of the fused GRU launches.  No CUDA here: these run on a data-loader worker thread (ctypes
// releases the GIL), ~50 us per review side instead of ~0.6 ms of numpy, so that eight ranks sharing one host do not wait for plans.
// The numpy forms in plan.py remain the specification (tests/test_plan.py holds both to the same arrays).
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "../../include/umpr_b200.h"

using namespace umpr;

// plan (int32, 3*Rp + n_tiles + 1 + n_slabs entries, Rp = n_tiles*R): [seq_of | row_of | len_of | tile_off | slab_tile]
extern "C" int umpr_plan_build(const int64_t* sorted_idx, const int64_t* sorted_len, long n, int R, int32_t* plan, long plan_ints,
                               int64_t* tokens_out) {
  if (!sorted_idx || !sorted_len || !plan || n < 1 || R < 1) return fail_arg("plan_build: bad arguments");
  const long n_tiles = (n + R - 1) / R, Rp = n_tiles * R;
  long n_slabs = 0;
  for (long j = 0; j < n_tiles; ++j) n_slabs += sorted_len[j * R];
  if (plan_ints != 3 * Rp + n_tiles + 1 + n_slabs) return fail_arg("plan_build: buffer of %ld ints, %ld needed", plan_ints, 3 * Rp + n_tiles + 1 + n_slabs);
  int32_t* seq_of = plan, *row_of = plan + Rp, *len_of = plan + 2 * Rp, *tile_off = plan + 3 * Rp, *slab_tile = tile_off + n_tiles + 1;
  int64_t tokens = 0;
  for (long k = 0; k < n; ++k) {
    const int64_t s = sorted_idx[k];
    if (s < 0 || s >= n) return fail_arg("plan_build: sorted index %lld out of range", (long long)s);
    seq_of[k] = (int32_t)s;
    row_of[k] = (int32_t)sorted_idx[s];            // output row fed by job k: the un-sort applied twice (model.py:21)
    len_of[k] = (int32_t)sorted_len[k];
    tokens += sorted_len[k];
  }
  for (long k = n; k < Rp; ++k) { seq_of[k] = 0; row_of[k] = -1; len_of[k] = 0; }
  long o = 0;
  for (long j = 0; j < n_tiles; ++j) {
    tile_off[j] = (int32_t)o;
    const long len = sorted_len[j * R];
    for (long t = 0; t < len; ++t) slab_tile[o + t] = (int32_t)j;
    o += len;
  }
  tile_off[n_tiles] = (int32_t)o;
  if (tokens_out) *tokens_out = tokens;
  return 0;
}

// Valid-row tile table (plan.py: snet_table with extra = 0, cnet_table with extra = 2): sentence n (an OUTPUT row of ImprovedRnn: row
// sorted_idx[j] holds sequence j) takes len + extra tile rows; consecutive sentences are grouped into tiles of at most 128 rows by the
// window index cstart // (129 - (L + extra)).  table (capacity 2*(n+1) ints) <- [tile_sent_off (n_tiles+1) | cstart (n+1)].
extern "C" int umpr_plan_table(const int64_t* sorted_idx, const int64_t* lengths, long n, int L, int extra, int32_t* table,
                               int32_t* n_tiles_out) {
  if (!sorted_idx || !lengths || !table || !n_tiles_out || n < 1) return fail_arg("plan_table: bad arguments");
  const int win = 129 - (L + extra);
  if (win < 1) return fail_arg("plan_table: sentence length %d (+%d) exceeds a 128-row tile", L, extra);
  std::vector<int32_t> row_len((size_t)n);
  for (long j = 0; j < n; ++j) row_len[(size_t)sorted_idx[j]] = (int32_t)lengths[j] + extra;
  int32_t* cstart = table + (n + 1);             // tail half first, moved behind the tile offsets below
  long run = 0;
  for (long i = 0; i < n; ++i) { cstart[i] = (int32_t)run; run += row_len[(size_t)i]; }
  cstart[n] = (int32_t)run;
  long nt = 0;
  long prev = -1;
  for (long i = 0; i < n; ++i) {
    const long w = cstart[i] / win;
    if (w != prev) { table[nt++] = (int32_t)i; prev = w; }      // a new tile wherever the window index changes: no empty tiles
  }
  table[nt] = (int32_t)n;
  memmove(table + nt + 1, cstart, sizeof(int32_t) * (size_t)(n + 1));
  *n_tiles_out = (int32_t)nt;
  return 0;
}

// Tile schedule of one fused GRU launch (plan.py: build_schedule): tiles of all segments, longest first (stable), dealt in
// boustrophedon order to the 2*G slot queues of G = min(n_ctas, T) CTAs.  sched (2*G + 1 + T ints) <- [q_off | q_tile]; returns G.
extern "C" int umpr_plan_schedule(const int64_t* const* tile_lens, const int32_t* n_tiles, int n_seg, int n_ctas, int32_t* sched,
                                  long sched_ints, int32_t* n_queues_out) {
  if (!tile_lens || !n_tiles || !sched || !n_queues_out || n_seg < 1 || n_ctas < 1) return fail_arg("plan_schedule: bad arguments");
  long T = 0;
  for (int s = 0; s < n_seg; ++s) T += n_tiles[s];
  if (T < 1) return fail_arg("plan_schedule: empty launch");
  const long G = std::min<long>(n_ctas, T), Q = 2 * G;
  if (sched_ints < Q + 1 + T) return fail_arg("plan_schedule: buffer of %ld ints, %ld needed", sched_ints, Q + 1 + T);
  std::vector<int64_t> lens((size_t)T);
  long o = 0;
  for (int s = 0; s < n_seg; ++s) for (long j = 0; j < n_tiles[s]; ++j) lens[(size_t)o++] = tile_lens[s][j];
  std::vector<int32_t> order((size_t)T);
  for (long i = 0; i < T; ++i) order[(size_t)i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return lens[(size_t)a] > lens[(size_t)b]; });
  std::vector<int32_t> queue((size_t)T), counts((size_t)Q, 0);
  for (long i = 0; i < T; ++i) {
    const long p = i / Q, pos = i % Q;
    const long sq = (p % 2 == 0) ? pos : Q - 1 - pos;            // slot queue in dealing order: [slot 0 of CTA 0..G-1 | slot 1 of CTA 0..G-1]
    const long q = 2 * (sq % G) + sq / G;                        // kernel order: CTA c reads queues 2c (slot 0) and 2c+1 (slot 1)
    queue[(size_t)i] = (int32_t)q;
    counts[(size_t)q]++;
  }
  int32_t* q_off = sched, *q_tile = sched + Q + 1;
  long run = 0;
  for (long q = 0; q < Q; ++q) { q_off[q] = (int32_t)run; run += counts[(size_t)q]; }
  q_off[Q] = (int32_t)run;
  std::vector<int32_t> cur(q_off, q_off + Q);
  for (long i = 0; i < T; ++i) q_tile[cur[(size_t)queue[(size_t)i]]++] = order[(size_t)i];      // dealing order = longest first inside a queue
  *n_queues_out = (int32_t)Q;
  return 0;
}
