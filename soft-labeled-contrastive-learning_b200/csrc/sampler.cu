// sampler.cu -- class-balanced pixel sampler building blocks (north_star item 1):
// stable per-class compaction of pixel indices, gather + L2-normalise of the
// sampled rows out of an NCHW map (fp32 and/or bf16, K-major rows for the
// tensor-core kernel), and the matching scatter-add backward.
//
// The reference fork has no sampler; the
// deterministic samplers it does have are LocalConLoss' stride slice
// (utils/loss.py:401-404) and BlockConLoss' tiles (:430-437).  Contract of this
// file: for every class k the compacted index list equals
// torch.nonzero(labels == k).squeeze(1) bit for bit; which of those indices are
// *used* is decided by the caller from PyTorch's RNG stream (randperm), so the
// selection is bit-exact by construction.
//
// Roofline: HBM (integer / gather work; no tensor cores).
#include <algorithm>

#include "common.cuh"

#include <cuda_bf16.h>
#include <limits.h>
#include <math.h>

namespace slcl {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int KM = SLCL_MAX_CLASSES;
// pixels per block of the compaction kernels (a multiple of 256): 4096 for large maps, 1024 below 2M pixels so that a
// 64K-pixel map (cfg3) still spreads over 64 blocks
inline int compact_chunk(int64_t n) { return n >= ((int64_t)1 << 21) ? 4096 : 1024; }

// phase 1: per-block class histogram -> hist[block][K]
// `skip`: optional device flag -- the launch does nothing when *skip == 0 (the sampler's second-phase compactions are
// only needed when some class is short; the decision is taken on the device so the launch sequence stays static)
__global__ void __launch_bounds__(kThreads) compact_count_kernel(const int64_t* labels, int64_t n, int K, int* hist, int kChunk,
                                                                 const int64_t* skip) {
  if (skip != nullptr && *skip == 0) return;
  __shared__ int s_cnt[KM];
  if (threadIdx.x < KM) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t start = (int64_t)blockIdx.x * kChunk;
  int local[KM];
#pragma unroll
  for (int k = 0; k < KM; ++k) local[k] = 0;
  for (int r = 0; r < kChunk / kThreads; ++r) {
    int64_t i = start + r * kThreads + threadIdx.x;
    if (i < n) {
      long long lab = labels[i];
#pragma unroll
      for (int k = 0; k < KM; ++k) local[k] += (lab == k) ? 1 : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    int t = local[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && k < K && t) atomicAdd(&s_cnt[k], t);     // integer: order-independent
  }
  __syncthreads();
  if (threadIdx.x < K) hist[(int64_t)blockIdx.x * K + threadIdx.x] = s_cnt[threadIdx.x];
}

// phase 2 (one block): counts, class offsets, and per-block start offsets (in place in hist)
__global__ void __launch_bounds__(kThreads) compact_scan_kernel(int* hist, int n_blocks, int K, int64_t* counts,
                                                                int64_t* offsets, int64_t* block_base, const int64_t* skip) {
  if (skip != nullptr && *skip == 0) return;
  __shared__ long long s_total[KM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp k scans class k over the blocks (sequential over 32-wide strips: n_blocks is small)
  if (warp < K) {
    long long run = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 32) {
      int b = b0 + lane;
      int v = (b < n_blocks) ? hist[(int64_t)b * K + warp] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (b < n_blocks) block_base[(int64_t)b * K + warp] = run + (incl - v);
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_total[warp] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long off = 0;
    for (int k = 0; k < K; ++k) { counts[k] = s_total[k]; offsets[k] = off; off += s_total[k]; }
    offsets[K] = off;
  }
}

// phase 3: stable scatter
__global__ void __launch_bounds__(kThreads) compact_write_kernel(const int64_t* labels, int64_t n, int K,
                                                                 const int64_t* offsets, const int64_t* block_base,
                                                                 int64_t* index, int kChunk, const int64_t* skip) {
  if (skip != nullptr && *skip == 0) return;
  __shared__ int s_warp[KM][kWarps];
  __shared__ long long s_run[KM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < K) s_run[threadIdx.x] = offsets[threadIdx.x] + block_base[(int64_t)blockIdx.x * K + threadIdx.x];
  __syncthreads();
  const int64_t start = (int64_t)blockIdx.x * kChunk;
  for (int r = 0; r < kChunk / kThreads; ++r) {
    const int64_t i = start + r * kThreads + threadIdx.x;
    const long long lab = (i < n) ? labels[i] : -1;
    int my_rank = 0;
    for (int k = 0; k < K; ++k) {
      unsigned m = __ballot_sync(0xffffffffu, lab == k);
      if (lane == 0) s_warp[k][warp] = __popc(m);
      if (lab == k) my_rank = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (lab >= 0 && lab < K) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += s_warp[lab][w];
      index[s_run[lab] + before + my_rank] = i;
    }
    __syncthreads();
    if (threadIdx.x < K) {
      int t = 0;
      for (int w = 0; w < kWarps; ++w) t += s_warp[threadIdx.x][w];
      s_run[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// ---- class-balanced two-phase pick (the sampler of SURVEY.md 8(c)-3; test-side restatement: sample_class_balanced) ----
// lp[i] = labels[perm[i]]: the label map read in permutation order
__global__ void __launch_bounds__(kThreads) gather_labels_kernel(const int64_t* perm, const int64_t* labels, int64_t n, int64_t* lp) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int64_t p = perm[i];
  lp[i] = (p >= 0 && p < n) ? labels[p] : -1;
}

struct PickArgs {
  const int64_t* perm; const int64_t* counts; const int64_t* offsets; const int64_t* index;
  int K;
  int64_t per[2];          // picks per class of the two quotas (0 = quota unused)
  int64_t* out[2];         // [K * per] pixel indices, class-major
  int64_t* flag[2];        // [N] in permutation order, preset to -1: 0 marks a labelled pixel phase 1 left unpicked
  int64_t* need[2];        // [1] slots phase 1 leaves open
};
// One thread per entry e of the stable class-major compaction of lp: class k, rank r within the class (in permutation
// order).  Phase 1: the first `per` pixels of every class, class-major.
__global__ void __launch_bounds__(kThreads) balanced_pick_kernel(const PickArgs a) {
  __shared__ long long s_off[KM + 1], s_base[2][KM];
  if (threadIdx.x == 0) {
    for (int k = 0; k <= a.K; ++k) s_off[k] = a.offsets[k];
    for (int q = 0; q < 2; ++q) {
      long long base = 0;
      for (int k = 0; k < a.K; ++k) { s_base[q][k] = base; base += a.counts[k] < a.per[q] ? a.counts[k] : a.per[q]; }
      if (blockIdx.x == 0 && a.out[q] != nullptr) *a.need[q] = a.K * a.per[q] - base;
    }
  }
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= s_off[a.K]) return;
  int k = 0;
  while (e >= s_off[k + 1]) ++k;
  const int64_t r = e - s_off[k], p = a.index[e], pix = a.perm[p];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (a.out[q] == nullptr) continue;
    if (r < a.per[q]) a.out[q][s_base[q][k] + r] = pix;
    else a.flag[q][p] = 0;
  }
}
// Phase 2: the slots a short class leaves open take the next unpicked labelled pixels in permutation order (`index2`:
// their positions, compacted in order); slots that stay open hold 0 and `filled` < K * per tells the caller.
__global__ void __launch_bounds__(kThreads) balanced_fill_kernel(const int64_t* perm, const int64_t* need_p, int64_t total,
                                                                 const int64_t* counts2, const int64_t* index2, int64_t* out,
                                                                 int64_t* filled) {
  const int64_t need = *need_p;
  if (need == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) *filled = total; return; }
  const int64_t n1 = total - need, avail = counts2[0], m = need < avail ? need : avail;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < need; r += (int64_t)gridDim.x * kThreads)
    out[n1 + r] = r < m ? perm[index2[r]] : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *filled = n1 + m;
}

// ---- self-pair maps for ids in [0, n_ids) (pixel indices): two lookup tables instead of a sort ----
__global__ void __launch_bounds__(kThreads) self_table_kernel(const int64_t* id_a, int64_t na, const int64_t* id_b, int64_t m,
                                                              int64_t n_ids, int32_t* table_a, int32_t* table_b) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < na) { const int64_t v = id_a[i]; if (v >= 0 && v < n_ids) table_a[v] = (int32_t)i; }
  if (i < m) { const int64_t v = id_b[i]; if (v >= 0 && v < n_ids) table_b[v] = (int32_t)i; }
}
__global__ void __launch_bounds__(kThreads) self_lookup_kernel(const int64_t* id_a, int64_t na, const int64_t* id_b, int64_t m,
                                                               int64_t n_ids, const int32_t* table_a, const int32_t* table_b,
                                                               int32_t* selfcol, int32_t* selfrow) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < na) { const int64_t v = id_a[i]; selfcol[i] = (v >= 0 && v < n_ids) ? table_b[v] : -1; }
  if (i < m) { const int64_t v = id_b[i]; selfrow[i] = (v >= 0 && v < n_ids) ? table_a[v] : -1; }
}

// ---- {label, id} metadata rows of the tensor-core sweeps straight from the label map ----
// meta[r] = {labels[idx[r]], (int)idx[r]} for r < n_rows, {INT_MIN, INT_MIN} for the pad rows up to n_pad
__global__ void __launch_bounds__(kThreads) rows_meta_kernel(const int64_t* labels, int64_t n_pixels, const int64_t* idx,
                                                             int64_t n_rows, int64_t n_pad, int2* meta) {
  const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (r >= n_pad) return;
  int2 v = make_int2(INT_MIN, INT_MIN);
  if (r < n_rows) {
    const int64_t p = idx[r];
    if (p >= 0 && p < n_pixels) {
      const long long lab = labels[p];
      v = make_int2(lab >= INT_MIN + 2 && lab <= INT_MAX ? (int)lab : INT_MIN + 1, (int)p);
    }
  }
  meta[r] = v;
}

// ---- row weights of the SupCon family from the metadata rows: fg_r / sum_tile fg / #tiles with foreground ----
// (utils/loss.py:382-384 for one problem; :445-448 / :465 folded into the rows for BlockConLoss' tiles.)
// pass 1: tile_fg[t] = number of rows of tile t whose label is not 0 (background); one block per tile, fixed order
__global__ void __launch_bounds__(kThreads) tile_fg_kernel(const int2* meta, int64_t rows_per_tile, float* tile_fg) {
  __shared__ int red[kWarps];
  const int2* m = meta + (int64_t)blockIdx.x * rows_per_tile;
  int cnt = 0;
  for (int64_t r = threadIdx.x; r < rows_per_tile; r += kThreads) cnt += m[r].x != 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kWarps; ++w) t += red[w];
    tile_fg[blockIdx.x] = (float)t;
  }
}
// pass 2: weight_r = fg_r / tile_fg / n_keep, n_keep = tiles with foreground (every block counts them itself).
// zero_if_empty: a tile (or the whole problem) without foreground contributes 0 (LocalConLoss / BlockConLoss early-outs);
// otherwise the division is left as it is (0/0 = NaN, what SupConLoss returns there).
__global__ void __launch_bounds__(kThreads) tile_weight_kernel(const int2* meta, int64_t n_tiles, int64_t rows_per_tile,
                                                               const float* tile_fg, int zero_if_empty, float* weight) {
  __shared__ int red[kWarps];
  __shared__ float s_keep;
  int cnt = 0;
  for (int64_t t = threadIdx.x; t < n_tiles; t += kThreads) cnt += tile_fg[t] > 0.f ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kWarps; ++w) t += red[w];
    s_keep = (float)t;
  }
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (r >= n_tiles * rows_per_tile) return;
  const float fg = meta[r].x != 0 ? 1.f : 0.f;
  float tf = tile_fg[r / rows_per_tile], keep = s_keep;
  if (zero_if_empty) {
    const float k = tf > 0.f ? 1.f : 0.f;
    weight[r] = fg / fmaxf(tf, 1.f) * k / fmaxf(keep, 1.f);
  } else {
    weight[r] = n_tiles == 1 ? fg / tf : fg / tf / keep;
  }
}

// One warp per sampled row: gather C strided values, L2-normalise, write row-major.
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const float* feat, int64_t C, int64_t HW,
                                                               const int64_t* pixel_idx, int64_t n_rows, int normalize,
                                                               __nv_bfloat16* out_bf16, int64_t bf16_stride, float* out_f32,
                                                               float* inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t pix = pixel_idx[row];
  const int64_t b = pix / HW, p = pix - b * HW;
  const float* base = feat + b * C * HW + p;
  float ss = 0.f;
  constexpr int kRegs = 8;                                   // C <= 256: the row stays in registers (one touch, 8 loads in flight)
  float v[kRegs];
  const bool in_regs = C <= 32 * kRegs;
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
      const int64_t c = lane + 32 * j;
      v[j] = c < C ? __ldg(base + c * HW) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kRegs; ++j) ss = fmaf(v[j], v[j], ss);
  } else {
    for (int64_t c = lane; c < C; c += 32) { float x = base[c * HW]; ss = fmaf(x, x, ss); }
  }
  ss = warp_sum(ss);
  const float inv_n = 1.0f / fmaxf(sqrtf(ss), 1e-12f);      // always reported: callers derive the exp shift from it
  const float inv = normalize ? inv_n : 1.0f;
  if (lane == 0 && inv_norm) inv_norm[row] = inv_n;
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
      const int64_t c = lane + 32 * j;
      if (c < C) {
        const float x = v[j] * inv;
        if (out_f32) out_f32[row * C + c] = x;
        if (out_bf16) out_bf16[row * bf16_stride + c] = __float2bfloat16_rn(x);
      }
    }
  } else {
    for (int64_t c = lane; c < C; c += 32) {
      float x = base[c * HW] * inv;       // second touch hits L1/L2 (the row's sectors were just loaded)
      if (out_f32) out_f32[row * C + c] = x;
      if (out_bf16) out_bf16[row * bf16_stride + c] = __float2bfloat16_rn(x);
    }
  }
  if (out_bf16) {
    for (int64_t c = C + lane; c < bf16_stride; c += 32) out_bf16[row * bf16_stride + c] = __float2bfloat16_rn(0.f);
  }
}

// The same gather for narrow maps (C <= 64, rows padded to 64 bf16): ONE THREAD per sampled row, so a warp's load of one
// channel touches 32 pixels -- four sectors when the rows are neighbouring pixels (dense tiles: BlockConLoss, strided
// grids) and never more sectors than the warp-per-row layout when they are scattered.  The row stays in registers; the
// block's 256 x 128-byte output region is written through shared memory with coalesced 16-byte stores.
constexpr int kDenseC = 64;
constexpr int kDenseRows = 256;
constexpr int kDensePitch = kDenseC + 8;          // bf16 elements: 144-byte pitch, conflict-free 16-byte accesses
__global__ void __launch_bounds__(kDenseRows) gather_rows_dense_kernel(const float* feat, int64_t C, int64_t HW,
                                                                       const int64_t* pixel_idx, int64_t n_rows, int normalize,
                                                                       __nv_bfloat16* out_bf16, float* inv_norm) {
  __shared__ __align__(16) __nv_bfloat16 tile[kDenseRows][kDensePitch];
  const int64_t row0 = (int64_t)blockIdx.x * kDenseRows;
  const int64_t row = row0 + threadIdx.x;
  const bool ok = row < n_rows;
  const int64_t pix = ok ? pixel_idx[row] : 0;
  const int64_t b = pix / HW, p = pix - b * HW;
  const float* base = feat + b * C * HW + p;
  float v[kDenseC];
#pragma unroll
  for (int c = 0; c < kDenseC; ++c) v[c] = (ok && c < C) ? __ldg(base + (int64_t)c * HW) : 0.f;
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < kDenseC; ++c) ss = fmaf(v[c], v[c], ss);
  const float inv_n = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  const float inv = normalize ? inv_n : 1.0f;
  if (ok && inv_norm) inv_norm[row] = inv_n;
#pragma unroll
  for (int k = 0; k < kDenseC / 8; ++k) {
    uint4 q;
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * k + 2 * u] * inv, v[8 * k + 2 * u + 1] * inv);
      w[u] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(&tile[threadIdx.x][8 * k]) = q;
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(out_bf16 + row0 * kDenseC);
#pragma unroll
  for (int k = 0; k < kDenseC / 8; ++k) {
    const int q = threadIdx.x + kDenseRows * k, r = q >> 3, ch = q & 7;
    if (row0 + r < n_rows) out[q] = *reinterpret_cast<const uint4*>(&tile[r][8 * ch]);
  }
}

// Scatter for narrow maps, one thread per row (see gather_rows_dense_kernel): the block's row gradients come in through
// shared memory (coalesced), every thread then owns its row, and a warp's atomic adds of one channel hit 32 pixels.
constexpr int kScatterRows = 128;
__global__ void __launch_bounds__(kScatterRows) scatter_rows_dense_kernel(const float* feat, int64_t C, int64_t HW,
                                                                          const int64_t* pixel_idx, int64_t n_rows,
                                                                          int normalize, const float* d_rows,
                                                                          const float* inv_norm, float* dfeat) {
  extern __shared__ float s_g[];                    // [kScatterRows][C + 1]
  const int pitch = (int)C + 1;
  const int64_t row0 = (int64_t)blockIdx.x * kScatterRows;
  const int64_t n_here = min((int64_t)kScatterRows, n_rows - row0);
  for (int64_t e = threadIdx.x; e < n_here * C; e += kScatterRows) {
    const int r = (int)(e / C), c = (int)(e - (int64_t)r * C);
    s_g[r * pitch + c] = d_rows[row0 * C + e];
  }
  __syncthreads();
  const int64_t row = row0 + threadIdx.x;
  if (row >= n_rows) return;
  const int64_t pix = pixel_idx[row];
  const int64_t b = pix / HW, p = pix - b * HW;
  const int64_t base = b * C * HW + p;
  const float* g = s_g + threadIdx.x * pitch;
  if (!normalize) {
    for (int c = 0; c < (int)C; ++c) atomicAdd(dfeat + base + (int64_t)c * HW, g[c]);
    return;
  }
  const float inv = inv_norm[row];
  float dotv = 0.f;
  for (int c = 0; c < (int)C; ++c) dotv = fmaf(__ldg(feat + base + (int64_t)c * HW) * inv, g[c], dotv);
  for (int c = 0; c < (int)C; ++c)
    atomicAdd(dfeat + base + (int64_t)c * HW, (g[c] - __ldg(feat + base + (int64_t)c * HW) * inv * dotv) * inv);
}

// exp shift of un-normalised rows: shift_i = |a_i| max_j |b_j| / T >= S_ij / T (the losses are shift-invariant; the bound
// keeps exp() in range).  Every block finds the maximum itself (inv_b is a few hundred KB, L2-resident) -- one launch, no
// grid barrier -- and writes its slice.  min over 1/|b| = 1 / max |b|.
__global__ void __launch_bounds__(1024) p2p_shift_kernel(const float* inv_a, int64_t na, const float* inv_b, int64_t m, float inv_t,
                                                         float* shift) {
  __shared__ float red[32];
  float lo = INFINITY;
  for (int64_t j = threadIdx.x; j < m; j += 1024) lo = fminf(lo, __ldg(inv_b + j));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lo;
  __syncthreads();
  lo = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  const float scale = (1.0f / lo) * inv_t;
  for (int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x; i < na; i += (int64_t)gridDim.x * 1024)
    shift[i] = (1.0f / inv_a[i]) * scale;
}

// dx = (g - xhat (xhat . g)) * inv_norm, scattered (atomic add) into NCHW dfeat
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(const float* feat, int64_t C, int64_t HW,
                                                                const int64_t* pixel_idx, int64_t n_rows, int normalize,
                                                                const float* d_rows, const float* inv_norm,
                                                                float* dfeat) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t pix = pixel_idx[row];
  const int64_t b = pix / HW, p = pix - b * HW;
  const int64_t base = b * C * HW + p;
  const float inv = normalize ? inv_norm[row] : 1.0f;
  // C <= 256: feature row and gradient row stay in registers -- all the sparse loads of the row are in flight at once.
  // (The adds stay fire-and-forget reductions even where every pixel is known to occur once: a plain read-modify-write
  // was measured slower, 88 vs 74 us per launch at cfg3 -- the load puts a second L2 round trip on every element.)
  constexpr int kRegs = 8;
  if (C <= 32 * kRegs) {
    float x[kRegs], g[kRegs];
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
      const int64_t c = lane + 32 * j;
      x[j] = (normalize && c < C) ? __ldg(feat + base + c * HW) * inv : 0.f;
      g[j] = c < C ? __ldg(d_rows + row * C + c) : 0.f;
    }
    float dotv = 0.f;
    if (normalize) {
#pragma unroll
      for (int j = 0; j < kRegs; ++j) dotv = fmaf(x[j], g[j], dotv);
      dotv = warp_sum(dotv);
    }
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
      const int64_t c = lane + 32 * j;
      if (c < C) atomicAdd(dfeat + base + c * HW, normalize ? (g[j] - x[j] * dotv) * inv : g[j]);
    }
    return;
  }
  float dotv = 0.f;
  if (normalize) {
    for (int64_t c = lane; c < C; c += 32) dotv = fmaf(feat[base + c * HW] * inv, d_rows[row * C + c], dotv);
    dotv = warp_sum(dotv);
  }
  for (int64_t c = lane; c < C; c += 32) {
    float g = d_rows[row * C + c];
    float v = normalize ? (g - feat[base + c * HW] * inv * dotv) * inv : g;
    atomicAdd(dfeat + base + c * HW, v);
  }
}

// Backward of the gather driven from the PIXEL side: map_s[pixel] = row of set s that holds the pixel, or -1.  Threads
// own pixels and walk channels, so loads of the feature map and stores of the gradient map are coalesced and EVERY
// element of dfeat is written exactly once (zeros where no row holds the pixel): no pre-zeroed map, no atomics, no
// 4-byte accesses scattered over 32-byte sectors.  Pays off when the sampled rows cover a good share of the map
// (cfg3: 20 480 rows of 65 536 pixels touch ~95 % of the map's sectors anyway).  Up to two row sets (anchors and
// contrast rows) are summed on the fly.
__global__ void __launch_bounds__(kThreads) scatter_by_map_kernel(const float* feat, int64_t C, int64_t HW, int64_t n_pix,
                                                                  int normalize, const int32_t* map0, const float* d0,
                                                                  const float* inv0, const int32_t* map1, const float* d1,
                                                                  const float* inv1, float* dfeat) {
  // block = 32 consecutive pixels (lane) x 8 channel groups (warp): a warp's load / store of one channel is one full
  // line, a thread's slice of its gradient row is contiguous, and 8 x more threads share the serial channel walk
  __shared__ float s_dot[2][kWarps][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * 32 + lane;
  const bool live = p < n_pix;
  const int64_t pc = live ? p : n_pix - 1;
  const int64_t b = pc / HW, q = pc - b * HW;
  const int64_t base = b * C * HW + q;
  const int r0 = live ? map0[pc] : -1;
  const int r1 = (live && map1 != nullptr) ? map1[pc] : -1;
  const bool any = r0 >= 0 || r1 >= 0;
  const int cg = (int)((C + kWarps - 1) / kWarps);
  const int c_begin = grp * cg, c_end = (int)min((int64_t)(grp + 1) * cg, C);
  if (!__any_sync(0xffffffffu, any)) {                 // (the same 32 pixels in every warp of the block: uniform exit)
    if (live)
      for (int c = c_begin; c < c_end; ++c) dfeat[base + (int64_t)c * HW] = 0.f;
    return;
  }
  const float* g0 = r0 >= 0 ? d0 + (int64_t)r0 * C : nullptr;
  const float* g1 = r1 >= 0 ? d1 + (int64_t)r1 * C : nullptr;
  const float i0 = (normalize && r0 >= 0) ? inv0[r0] : 1.f;
  const float i1 = (normalize && r1 >= 0) ? inv1[r1] : 1.f;
  float dot0 = 0.f, dot1 = 0.f;
  if (normalize) {
    float t0 = 0.f, t1 = 0.f;
    if (any) {
#pragma unroll 8
      for (int c = c_begin; c < c_end; ++c) {
        const float x = __ldg(feat + base + (int64_t)c * HW);
        if (g0) t0 = fmaf(x * i0, __ldg(g0 + c), t0);
        if (g1) t1 = fmaf(x * i1, __ldg(g1 + c), t1);
      }
    }
    s_dot[0][grp][lane] = t0; s_dot[1][grp][lane] = t1;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { dot0 += s_dot[0][w][lane]; dot1 += s_dot[1][w][lane]; }      // fixed order
  }
#pragma unroll 8
  for (int c = c_begin; c < c_end; ++c) {              // all lanes together: full-line stores, zeros where nothing was sampled
    float v = 0.f;
    if (any) {
      const float x = normalize ? __ldg(feat + base + (int64_t)c * HW) : 0.f;
      if (g0) { const float g = __ldg(g0 + c); v += normalize ? (g - x * i0 * dot0) * i0 : g; }
      if (g1) { const float g = __ldg(g1 + c); v += normalize ? (g - x * i1 * dot1) * i1 : g; }
    }
    if (live) dfeat[base + (int64_t)c * HW] = v;
  }
}

}  // namespace
}  // namespace slcl

using namespace slcl;

extern "C" size_t slcl_compact_workspace_bytes(int64_t n_pixels, int n_class) {
  if (n_pixels <= 0 || n_class < 1 || n_class > KM) return 0;
  size_t blocks = (size_t)ceil_div<int64_t>(n_pixels, compact_chunk(n_pixels));
  return align_up(blocks * n_class * sizeof(int), 256) + align_up(blocks * n_class * sizeof(int64_t), 256);
}

namespace slcl {
namespace {
// the three launches of the stable compaction; `skip` (device flag, may be null): do nothing when *skip == 0
void launch_compact(const int64_t* labels, int64_t n_pixels, int n_class, int64_t* counts, int64_t* offsets, int64_t* index,
                    void* workspace, const int64_t* skip, cudaStream_t stream) {
  const int chunk = compact_chunk(n_pixels);
  const int blocks = (int)ceil_div<int64_t>(n_pixels, chunk);
  int* hist = reinterpret_cast<int*>(workspace);
  int64_t* block_base =
      reinterpret_cast<int64_t*>((char*)workspace + align_up((size_t)blocks * n_class * sizeof(int), 256));
  compact_count_kernel<<<blocks, kThreads, 0, stream>>>(labels, n_pixels, n_class, hist, chunk, skip);
  compact_scan_kernel<<<1, kThreads, 0, stream>>>(hist, blocks, n_class, counts, offsets, block_base, skip);
  compact_write_kernel<<<blocks, kThreads, 0, stream>>>(labels, n_pixels, n_class, offsets, block_base, index, chunk, skip);
}
}  // namespace
}  // namespace slcl

extern "C" int slcl_compact_by_class(const int64_t* labels, int64_t n_pixels, int n_class, int64_t* counts,
                                     int64_t* offsets, int64_t* index, void* workspace, size_t workspace_bytes,
                                     slcl_stream_t stream_) {
  if (!labels || n_pixels <= 0 || !counts || !offsets || !index || !workspace) return SLCL_ERR_INVALID_ARGUMENT;
  if (n_class < 1 || n_class > KM) return SLCL_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < slcl_compact_workspace_bytes(n_pixels, n_class) || !aligned16(workspace))
    return SLCL_ERR_WORKSPACE;
  launch_compact(labels, n_pixels, n_class, counts, offsets, index, workspace, nullptr, (cudaStream_t)stream_);
  return check_launch("slcl_compact_by_class");
}

// workspace of slcl_sample_balanced: lp [N] | counts [K] offsets [K+1] | index [N] | flags 2 x [N] | need 2 | per quota:
// counts2 [1] offsets2 [2] index2 [N] | compaction scratch (shared by the three compactions, they run one after the other)
namespace slcl {
namespace {
struct SampleWs { int64_t *lp, *counts, *offsets, *index, *flag[2], *need, *counts2[2], *offsets2[2], *index2[2]; void* scratch; size_t total; };
SampleWs carve_sample(void* ws, int64_t n, int K) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t nb = (size_t)n * sizeof(int64_t);
  size_t o_lp = take(nb), o_c = take((size_t)K * 8), o_o = take((size_t)(K + 1) * 8), o_i = take(nb), o_f = take(2 * nb), o_n = take(16);
  size_t o_c2[2], o_o2[2], o_i2[2];
  for (int q = 0; q < 2; ++q) { o_c2[q] = take(8); o_o2[q] = take(16); o_i2[q] = take(nb); }
  size_t o_s = take(slcl_compact_workspace_bytes(n, K > 1 ? K : 1));
  SampleWs w{};
  char* b = reinterpret_cast<char*>(ws);
  w.lp = (int64_t*)(b + o_lp); w.counts = (int64_t*)(b + o_c); w.offsets = (int64_t*)(b + o_o); w.index = (int64_t*)(b + o_i);
  w.flag[0] = (int64_t*)(b + o_f); w.flag[1] = w.flag[0] + n; w.need = (int64_t*)(b + o_n);
  for (int q = 0; q < 2; ++q) { w.counts2[q] = (int64_t*)(b + o_c2[q]); w.offsets2[q] = (int64_t*)(b + o_o2[q]); w.index2[q] = (int64_t*)(b + o_i2[q]); }
  w.scratch = b + o_s;
  w.total = off;
  return w;
}
}  // namespace
}  // namespace slcl

extern "C" size_t slcl_sample_balanced_workspace_bytes(int64_t n_pixels, int n_class) {
  if (n_pixels <= 0 || n_class < 1 || n_class > KM) return 0;
  return carve_sample(nullptr, n_pixels, n_class).total;
}

extern "C" int slcl_sample_balanced(const int64_t* perm, const int64_t* labels, int64_t n_pixels, int n_class, int64_t per_a,
                                    int64_t* out_a, int64_t* filled_a, int64_t per_b, int64_t* out_b, int64_t* filled_b,
                                    void* workspace, size_t workspace_bytes, slcl_stream_t stream_) {
  if (!perm || !labels || n_pixels <= 0 || n_class < 1 || n_class > KM || per_a < 1 || !out_a || !filled_a || !workspace)
    return SLCL_ERR_INVALID_ARGUMENT;
  if (per_b < 0 || (per_b > 0 && (!out_b || !filled_b))) return SLCL_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < slcl_sample_balanced_workspace_bytes(n_pixels, n_class) || !aligned16(workspace)) return SLCL_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  SampleWs w = carve_sample(workspace, n_pixels, n_class);
  const int nb = (int)ceil_div<int64_t>(n_pixels, kThreads);
  gather_labels_kernel<<<nb, kThreads, 0, stream>>>(perm, labels, n_pixels, w.lp);
  launch_compact(w.lp, n_pixels, n_class, w.counts, w.offsets, w.index, w.scratch, nullptr, stream);
  cudaMemsetAsync(w.flag[0], 0xFF, 2 * (size_t)n_pixels * sizeof(int64_t), stream);          // -1: not a phase-2 candidate
  PickArgs a{};
  a.perm = perm; a.counts = w.counts; a.offsets = w.offsets; a.index = w.index; a.K = n_class;
  a.per[0] = per_a; a.out[0] = out_a; a.flag[0] = w.flag[0]; a.need[0] = w.need;
  a.per[1] = per_b; a.out[1] = per_b > 0 ? out_b : nullptr; a.flag[1] = w.flag[1]; a.need[1] = w.need + 1;
  balanced_pick_kernel<<<nb, kThreads, 0, stream>>>(a);
  for (int q = 0; q < (per_b > 0 ? 2 : 1); ++q) {
    const int64_t per = q == 0 ? per_a : per_b;
    launch_compact(w.flag[q], n_pixels, 1, w.counts2[q], w.offsets2[q], w.index2[q], w.scratch, w.need + q, stream);
    const int fb = (int)std::min<int64_t>(ceil_div<int64_t>(n_class * per, kThreads), 148);
    balanced_fill_kernel<<<fb, kThreads, 0, stream>>>(perm, w.need + q, n_class * per, w.counts2[q], w.index2[q],
                                                     q == 0 ? out_a : out_b, q == 0 ? filled_a : filled_b);
  }
  return check_launch("slcl_sample_balanced");
}

extern "C" int slcl_self_maps(const int64_t* id_a, int64_t n_anchor, const int64_t* id_b, int64_t n_contrast, int64_t n_ids,
                              int32_t* a_selfcol, int32_t* b_selfrow, int32_t* row_of_id, slcl_stream_t stream_) {
  if (!id_a || !id_b || n_anchor <= 0 || n_contrast <= 0 || n_ids <= 0 || !a_selfcol || !b_selfrow || !row_of_id)
    return SLCL_ERR_INVALID_ARGUMENT;
  if (n_anchor > INT_MAX || n_contrast > INT_MAX) return SLCL_ERR_UNSUPPORTED;
  cudaStream_t stream = (cudaStream_t)stream_;
  int32_t* table_a = row_of_id;
  int32_t* table_b = table_a + n_ids;
  cudaMemsetAsync(table_a, 0xFF, 2 * (size_t)n_ids * sizeof(int32_t), stream);          // -1: id not present
  const int nb = (int)ceil_div<int64_t>(std::max(n_anchor, n_contrast), kThreads);
  self_table_kernel<<<nb, kThreads, 0, stream>>>(id_a, n_anchor, id_b, n_contrast, n_ids, table_a, table_b);
  self_lookup_kernel<<<nb, kThreads, 0, stream>>>(id_a, n_anchor, id_b, n_contrast, n_ids, table_a, table_b, a_selfcol, b_selfrow);
  return check_launch("slcl_self_maps");
}

extern "C" int slcl_scatter_rows_by_map(const float* feat, int64_t batch, int64_t channels, int64_t pixels, int normalize,
                                        const int32_t* row_of_pixel_a, const float* d_rows_a, const float* inv_norm_a,
                                        const int32_t* row_of_pixel_b, const float* d_rows_b, const float* inv_norm_b,
                                        float* dfeat, slcl_stream_t stream_) {
  if (!feat || batch <= 0 || channels <= 0 || pixels <= 0 || !row_of_pixel_a || !d_rows_a || !dfeat) return SLCL_ERR_INVALID_ARGUMENT;
  if ((row_of_pixel_b == nullptr) != (d_rows_b == nullptr)) return SLCL_ERR_INVALID_ARGUMENT;
  if (normalize && (!inv_norm_a || (row_of_pixel_b && !inv_norm_b))) return SLCL_ERR_INVALID_ARGUMENT;
  const int64_t n_pix = batch * pixels;
  scatter_by_map_kernel<<<(unsigned)ceil_div<int64_t>(n_pix, 32), kThreads, 0, (cudaStream_t)stream_>>>(
      feat, channels, pixels, n_pix, normalize, row_of_pixel_a, d_rows_a, inv_norm_a, row_of_pixel_b, d_rows_b, inv_norm_b, dfeat);
  return check_launch("slcl_scatter_rows_by_map");
}

extern "C" int slcl_tile_weights(const int32_t* meta, int64_t n_tiles, int64_t rows_per_tile, int zero_if_empty, float* tile_fg,
                                 float* weight, slcl_stream_t stream_) {
  if (!meta || n_tiles <= 0 || n_tiles > INT_MAX || rows_per_tile <= 0 || !tile_fg || !weight) return SLCL_ERR_INVALID_ARGUMENT;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int2* m = reinterpret_cast<const int2*>(meta);
  tile_fg_kernel<<<(unsigned)n_tiles, kThreads, 0, stream>>>(m, rows_per_tile, tile_fg);
  tile_weight_kernel<<<(unsigned)ceil_div<int64_t>(n_tiles * rows_per_tile, kThreads), kThreads, 0, stream>>>(
      m, n_tiles, rows_per_tile, tile_fg, zero_if_empty, weight);
  return check_launch("slcl_tile_weights");
}

extern "C" int slcl_rows_meta(const int64_t* labels, int64_t n_pixels, const int64_t* pixel_idx, int64_t n_rows, int32_t* meta,
                              slcl_stream_t stream_) {
  if (!labels || n_pixels <= 0 || n_pixels > INT_MAX || !pixel_idx || n_rows <= 0 || !meta) return SLCL_ERR_INVALID_ARGUMENT;
  const int64_t n_pad = (int64_t)align_up((size_t)n_rows, 64);
  rows_meta_kernel<<<(int)ceil_div<int64_t>(n_pad, kThreads), kThreads, 0, (cudaStream_t)stream_>>>(
      labels, n_pixels, pixel_idx, n_rows, n_pad, reinterpret_cast<int2*>(meta));
  return check_launch("slcl_rows_meta");
}

extern "C" int slcl_gather_unit_rows(const float* feat, int64_t batch, int64_t channels, int64_t pixels,
                                     const int64_t* pixel_idx, int64_t n_rows, int normalize, void* rows_bf16,
                                     int64_t bf16_row_stride, float* rows_f32, float* inv_norm, slcl_stream_t stream_) {
  if (!feat || batch <= 0 || channels <= 0 || pixels <= 0 || !pixel_idx || n_rows <= 0) return SLCL_ERR_INVALID_ARGUMENT;
  if (!rows_bf16 && !rows_f32) return SLCL_ERR_INVALID_ARGUMENT;
  if (rows_bf16 && bf16_row_stride < channels) return SLCL_ERR_INVALID_ARGUMENT;
  if (channels <= kDenseC && rows_bf16 && !rows_f32 && bf16_row_stride == kDenseC && aligned16(rows_bf16)) {
    gather_rows_dense_kernel<<<(int)ceil_div<int64_t>(n_rows, kDenseRows), kDenseRows, 0, (cudaStream_t)stream_>>>(
        feat, channels, pixels, pixel_idx, n_rows, normalize, reinterpret_cast<__nv_bfloat16*>(rows_bf16), inv_norm);
    return check_launch("slcl_gather_unit_rows");
  }
  const int blocks = (int)ceil_div<int64_t>(n_rows, kWarps);
  gather_rows_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream_>>>(feat, channels, pixels, pixel_idx, n_rows, normalize,
                                                                    reinterpret_cast<__nv_bfloat16*>(rows_bf16),
                                                                    bf16_row_stride, rows_f32, inv_norm);
  return check_launch("slcl_gather_unit_rows");
}

extern "C" int slcl_p2p_shift(const float* inv_norm_a, int64_t n_anchor, const float* inv_norm_b, int64_t n_contrast,
                              float temperature, float* shift, slcl_stream_t stream_) {
  if (!inv_norm_a || !inv_norm_b || !shift || n_anchor <= 0 || n_contrast <= 0 || !(temperature > 0.f))
    return SLCL_ERR_INVALID_ARGUMENT;
  if (n_contrast > (int64_t)1 << 20) return SLCL_ERR_UNSUPPORTED;          // every block reads all of inv_norm_b
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n_anchor, 1024), n_contrast > 65536 ? 148 : 592);
  p2p_shift_kernel<<<blocks, 1024, 0, (cudaStream_t)stream_>>>(inv_norm_a, n_anchor, inv_norm_b, n_contrast, 1.0f / temperature,
                                                                shift);
  return check_launch("slcl_p2p_shift");
}

extern "C" int slcl_scatter_rows_bwd(const float* feat, int64_t batch, int64_t channels, int64_t pixels,
                                     const int64_t* pixel_idx, int64_t n_rows, int normalize, const float* d_rows,
                                     const float* inv_norm, float* dfeat, slcl_stream_t stream_) {
  if (!feat || batch <= 0 || channels <= 0 || pixels <= 0 || !pixel_idx || n_rows <= 0 || !d_rows || !dfeat)
    return SLCL_ERR_INVALID_ARGUMENT;
  if (normalize && !inv_norm) return SLCL_ERR_INVALID_ARGUMENT;
  if (channels <= kDenseC) {
    const size_t smem = (size_t)kScatterRows * (channels + 1) * sizeof(float);          // <= 33 KB
    scatter_rows_dense_kernel<<<(int)ceil_div<int64_t>(n_rows, kScatterRows), kScatterRows, smem, (cudaStream_t)stream_>>>(
        feat, channels, pixels, pixel_idx, n_rows, normalize, d_rows, inv_norm, dfeat);
    return check_launch("slcl_scatter_rows_bwd");
  }
  const int blocks = (int)ceil_div<int64_t>(n_rows, kWarps);
  scatter_rows_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream_>>>(feat, channels, pixels, pixel_idx, n_rows,
                                                                     normalize, d_rows, inv_norm, dfeat);
  return check_launch("slcl_scatter_rows_bwd");
}

