// The two data formats either side of the hot path:
//   * collate on device (reference src/dataset.py:122-131,163-171 pad_reviews / batch_loader): the host ships the RAGGED token lists -
//     one flat int32 array plus a prefix sum of the per-sentence token counts, ~2.4x fewer bytes than the padded int64 tensors - and
//     this kernel expands them into the (B, S, L) int64 id tensors the reference's collate would have produced (pad id beyond each
//     sentence; an empty slot is an all-PAD sentence whose length the host reports as 1, dataset.py:127);
//   * VGG16 feature cache (src/model.py:204-219 backbone output, src/dataset.py:134-151 image loading): the 1000-d backbone
//     features of every photo live in one device-resident table; a batch carries photo ROW INDICES (B, V, Pc) and this kernel
//     gathers the (B, V, Pc, F) feature tensor VisualNet's tail consumes.  A negative index = photo missing / unreadable: the
//     reference substitutes a zero image (dataset.py:147-148), whose features are whatever the table's `missing_row` holds.
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {

__global__ void __launch_bounds__(256) collate_ids_kernel(const int32_t* __restrict__ flat, const int32_t* __restrict__ off, long n_sent,
                                                          int L, int64_t pad, int64_t* __restrict__ ids) {
  // one warp per sentence slot, lanes stride over the L positions (coalesced 8-byte stores)
  const long n = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_sent) return;
  const int o = off[n], cnt = off[n + 1] - o;
  for (int t = lane; t < L; t += 32) ids[n * L + t] = t < cnt ? (int64_t)flat[o + t] : pad;
}

__global__ void __launch_bounds__(256) feature_gather_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx, long n_photos,
                                                             long rows, int F, long missing_row, float* __restrict__ out) {
  // one warp per photo: float4 copies of its F-float row
  const long n = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_photos) return;
  long r = idx[n];
  if (r < 0 || r >= rows) r = missing_row;
  float* dst = out + n * F;
  if (r < 0) {                                  // no stand-in row: zeros
    for (int c = lane; c < F; c += 32) dst[c] = 0.f;
    return;
  }
  const float* src = table + r * F;
  if ((F & 3) == 0) {
    for (int c = lane * 4; c < F; c += 128) *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(src + c);
  } else {
    for (int c = lane; c < F; c += 32) dst[c] = src[c];
  }
}

}  // namespace umpr

using namespace umpr;

extern "C" int umpr_collate_ids(const int32_t* flat_tokens, const int32_t* sent_off, long n_sent, int L, long pad_id, int64_t* ids, void* stream) {
  if (n_sent <= 0) return 0;
  if (!sent_off || !ids || L < 1) return fail_arg("collate_ids: NULL argument or L=%d", L);
  const long threads = n_sent * 32;
  collate_ids_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(flat_tokens, sent_off, n_sent, L, (int64_t)pad_id, ids);
  return check_launch("collate_ids");
}

extern "C" int umpr_feature_gather(const float* table, const int32_t* idx, long n_photos, long rows, int F, long missing_row, float* out,
                                   void* stream) {
  if (n_photos <= 0) return 0;
  if (!table || !idx || !out || F < 1 || rows < 1) return fail_arg("feature_gather: NULL argument, F=%d or rows=%ld", F, rows);
  if (missing_row >= rows) return fail_arg("feature_gather: missing_row=%ld outside the table (%ld rows)", missing_row, rows);
  if ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) return fail_arg("feature_gather: table and out must be 16-byte aligned");
  const long threads = n_photos * 32;
  feature_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, idx, n_photos, rows, F, missing_row, out);
  return check_launch("feature_gather");
}
