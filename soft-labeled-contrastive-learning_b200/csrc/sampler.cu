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
#include <math.h>

namespace slcl {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 4096;          // pixels per block (16 rounds of 256)
constexpr int KM = SLCL_MAX_CLASSES;

// phase 1: per-block class histogram -> hist[block][K]
__global__ void __launch_bounds__(kThreads) compact_count_kernel(const int64_t* labels, int64_t n, int K, int* hist) {
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
                                                                int64_t* offsets, int64_t* block_base) {
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
                                                                 int64_t* index) {
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

}  // namespace
}  // namespace slcl

using namespace slcl;

extern "C" size_t slcl_compact_workspace_bytes(int64_t n_pixels, int n_class) {
  if (n_pixels <= 0 || n_class < 1 || n_class > KM) return 0;
  size_t blocks = (size_t)ceil_div<int64_t>(n_pixels, kChunk);
  return align_up(blocks * n_class * sizeof(int), 256) + align_up(blocks * n_class * sizeof(int64_t), 256);
}

extern "C" int slcl_compact_by_class(const int64_t* labels, int64_t n_pixels, int n_class, int64_t* counts,
                                     int64_t* offsets, int64_t* index, void* workspace, size_t workspace_bytes,
                                     slcl_stream_t stream_) {
  if (!labels || n_pixels <= 0 || !counts || !offsets || !index || !workspace) return SLCL_ERR_INVALID_ARGUMENT;
  if (n_class < 1 || n_class > KM) return SLCL_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < slcl_compact_workspace_bytes(n_pixels, n_class) || !aligned16(workspace))
    return SLCL_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int blocks = (int)ceil_div<int64_t>(n_pixels, kChunk);
  int* hist = reinterpret_cast<int*>(workspace);
  int64_t* block_base =
      reinterpret_cast<int64_t*>((char*)workspace + align_up((size_t)blocks * n_class * sizeof(int), 256));
  compact_count_kernel<<<blocks, kThreads, 0, stream>>>(labels, n_pixels, n_class, hist);
  compact_scan_kernel<<<1, kThreads, 0, stream>>>(hist, blocks, n_class, counts, offsets, block_base);
  compact_write_kernel<<<blocks, kThreads, 0, stream>>>(labels, n_pixels, n_class, offsets, block_base, index);
  return check_launch("slcl_compact_by_class");
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
