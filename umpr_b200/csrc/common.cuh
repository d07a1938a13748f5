// Shared device/host helpers for the UMPR sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace umpr {

// fixed hot-path dimensions (config.py:34-37 defaults; the reference's published configs never change them)
constexpr int H = 64;        // gru_size
constexpr int D = 128;       // 2*gru_size
constexpr int G3 = 192;      // 3*gru_size, gate rows [r; z; n]
constexpr int ATT = 64;      // self_atte_size
constexpr int KP = 64;       // packed input width: E embedding floats, a 1.0 bias column, zero padding
constexpr int SV = 256;      // saved per (token, direction): r, z, n, (W_hn h + b_hn)

void set_error(const char* fmt, ...);
int fail_arg(const char* fmt, ...);
int check_launch(const char* what);
int gather_pack_tc_sides(const float* table, int n_sides, const int64_t* const* ids, const int32_t* const* plan, const int* n_tiles,
                         const int* n_slabs, const int* L, int E, void* const* xq, void* stream);
int dbg_flags();      // UMPR_DBG (development only): role-ablation switches of the tile kernels, 0 in normal operation
// internal form of umpr_cnet_conv_fwd_tc (csrc/cnet_tc.cu): prep = 0 reuses the weight image already in `wimg`, 2 = only build it;
// fix_records (optional): 2*N*KC bytes for the re-scoring records instead of the area behind the image
int cnet_conv_fwd_tc_impl(const float* x, const float* conv_w, const float* conv_b, int N, int L, int KC, int ksize, const int32_t* table,
                          int table_tiles, void* wimg, int cap, float* cfeat, int32_t* cidx, int n_ctas, int prep, void* fix_records, void* stream);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// block-wide sum for blockDim.x <= 1024; `red` is >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) { r = warp_sum(r); if (lane == 0) red[0] = r; }
  __syncthreads();
  r = red[0];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) { r = warp_max(r); if (lane == 0) red[0] = r; }
  __syncthreads();
  r = red[0];
  return r;
}

// order-preserving float -> uint32 (for packed (value, index) atomicMax)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- pack plan (device int32 buffer) -------------------------------------------------
// [seq_of (Rp) | row_of (Rp) | len_of (Rp) | tile_off (n_tiles+1) | slab_tile (n_slabs)],  Rp = n_tiles*R.
// Entry k of the first three arrays describes one GRU job in descending-length order:
//   seq_of[k] : which input sequence feeds it            (= sorted_indices[k] of the reference's torch.sort)
//   row_of[k] : which OUTPUT row of ImprovedRnn it fills (= sorted_indices[sorted_indices[k]], model.py:21)
//   len_of[k] : its length (0 for the padding jobs of the last tile)
// tile_off[j] is the first slab of tile j (a slab = R jobs at one time step); tile j owns len_of[j*R] slabs.
struct Plan {
  const int* seq_of; const int* row_of; const int* len_of; const int* tile_off; const int* slab_tile;
  int n_tiles, n_slabs, R;
};
inline Plan make_plan(const int32_t* buf, int n_tiles, int n_slabs, int R) {
  Plan p; const int Rp = n_tiles * R;
  p.seq_of = buf; p.row_of = buf + Rp; p.len_of = buf + 2 * Rp; p.tile_off = buf + 3 * Rp;
  p.slab_tile = p.tile_off + n_tiles + 1; p.n_tiles = n_tiles; p.n_slabs = n_slabs; p.R = R;
  return p;
}

// valid-row tables, software-pipelined by the loader threads: tile boundaries (tso) two tiles ahead, sentence offsets (cst) one tile
// ahead, so that no table load sits in front of a tile's row loads
struct TabPipe {
  int s0c, s1c, s0n, s1n, cbc, cec, c0c, cendc;         // current tile, next tile's boundaries
  int s0nn, s1nn, cbn, cen, c0n, cendn;                 // in flight
  __device__ __forceinline__ void ld_tso(const int* __restrict__ tso, int tile, int n_tiles, int& s0, int& s1) {
    if (tile < n_tiles) { s0 = tso[tile]; s1 = tso[tile + 1]; } else { s0 = s1 = 0; }
  }
  __device__ __forceinline__ void ld_cst(const int* __restrict__ cst, int s0, int s1, int tid, int& cb, int& ce, int& c0, int& cend) {
    c0 = cst[s0]; cend = cst[s1];
    cb = ce = 0;
    if (tid < s1 - s0) { cb = cst[s0 + tid]; ce = cst[s0 + tid + 1]; }
  }
  __device__ __forceinline__ void init(const int* tso, const int* cst, int tile, int stride, int n_tiles, int tid) {
    ld_tso(tso, tile, n_tiles, s0c, s1c);
    ld_tso(tso, tile + stride, n_tiles, s0n, s1n);
    ld_cst(cst, s0c, s1c, tid, cbc, cec, c0c, cendc);
  }
  __device__ __forceinline__ void prefetch(const int* tso, const int* cst, int tile, int stride, int n_tiles, int tid) {
    ld_cst(cst, s0n, s1n, tid, cbn, cen, c0n, cendn);       // (past the end s0n == s1n == 0: a harmless read of cst[0])
    ld_tso(tso, tile + 2 * stride, n_tiles, s0nn, s1nn);
  }
  __device__ __forceinline__ void rotate() {
    s0c = s0n; s1c = s1n; s0n = s0nn; s1n = s1nn;
    cbc = cbn; cec = cen; c0c = c0n; cendc = cendn;
  }
};


}  // namespace umpr
