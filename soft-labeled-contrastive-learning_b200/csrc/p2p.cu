// p2p.cu -- pixel <-> pixel supervised contrastive loss on the 5th-gen tensor cores.
//
// Replaces (reference, file:line):
//   SupConLoss.forward   utils/loss.py:327-387 (= utils/losses.py:106-161): the M x M Gram matrix
//                        via conv2d (:342-349), >= 6 materialised M x M fp32 temporaries (:352-380)
//   and its autograd backward, generalised to anchors x contrast rows
//   (SURVEY.md 8(c)-3).
//
// Roofline: tensor core.  Flash-attention style: the similarity matrix never exists in memory.
//   One CTA owns a 128-row tile of the "row operand" R (resident in shared memory) and streams
//   64-row tiles of the "column operand" Cm through a 3-stage TMA ring.
//     MMA1 (tcgen05.mma, bf16 -> fp32 TMEM):  S[128 x 64] = R_tile . Cm_tile^T        (K = d)
//     epilogue warps (tcgen05.ld, one thread per row): temperature, self/positive masks from
//       labels + pixel ids, exp -> row statistics (forward) or the gradient tile G (backward),
//       G is written as a bf16 K-major swizzled smem operand
//     MMA2 (backward only):  dR[128 x d] += G[128 x 64] . Cm_tile[64 x d]   (Cm tile re-used from
//       shared memory as an MN-major B operand; accumulators stay in TMEM for the whole sweep)
//   S is double-buffered in TMEM so MMA1 of tile t+1 overlaps the epilogue of tile t.
//   The same kernel runs forward (rows = anchors), dA (rows = anchors) and dB (rows = contrast
//   rows, columns = anchors: per-column statistics).
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

namespace slcl {
namespace {

constexpr int BM = 128;            // rows per CTA (UMMA M)
constexpr int BN = 64;             // streamed rows (S tile columns) per step
constexpr int KCH = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kMaxD = 256;
constexpr int kEpiWarps = 8;       // warps 4..11; (warp % 4) selects the TMEM lane quarter
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr int kTmemCols = 512;
constexpr int kColS = 0;           // S double buffer: columns [0,64) and [64,128)
constexpr int kColAcc = 128;       // dR accumulators: columns [128, 128 + d)
constexpr float kLog2e = 1.4426950408889634f;

enum Mode { kFwd = 0, kBwdRows = 1 /* dA: stats per row */, kBwdCols = 2 /* dB: stats per column */ };

struct P2PArgs {
  int n_rows, n_cols, d;              // d padded to a multiple of 64
  int col_begin, cols_per_split;      // this kernel instance sweeps columns [col_begin + split*cols_per_split, ...)
  int mode;
  float scale_log2;                   // log2(e) / T
  const int2* row_meta;               // {label, id}
  const int2* col_meta;
  const float4* row_stat;             // kFwd: {shift*log2e,-,-,-}; kBwdRows: {shift*log2e, alpha, beta, -}
  const float4* col_stat;             // kBwdCols: {shift*log2e, alpha, beta, -}
  float* stat_partial;                // kFwd: [n_slots][n_rows][3]  (Zs, P_raw, n)
  float* grad_partial;                // bwd:  [n_splits][n_rows][d] fp32
};

// --------------------------- PTX wrappers ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SM100 UMMA), 128-byte swizzle.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1     bits [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, M = 128.
//   [4,6) c_format = 1 (F32)  [7,10) a_format = 1 (BF16)  [10,13) b_format = 1 (BF16)
//   [15] a_major (0 = K)  [16] b_major (0 = K, 1 = MN)  [17,23) N >> 3  [24,29) M >> 4
__device__ __forceinline__ uint32_t make_idesc(int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct __align__(8) Barriers {
  uint64_t r_full;
  uint64_t c_full[kStages], c_empty[kStages];
  uint64_t s_full[2], s_empty[2];
  uint64_t g_full[2], g_empty[2];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// dynamic smem carve-up (1024-byte aligned tiles)
//   R   : [d/64][128 rows][128 B]
//   Cm  : [kStages][d/64][64 rows][128 B]
//   G   : [2][128 rows][128 B]
__host__ __device__ inline size_t smem_bytes_for(int d) {
  const size_t kc = d / KCH;
  return 1024 /*align slack*/ + kc * BM * 128 + (size_t)kStages * kc * BN * 128 + 2 * BM * 128 + sizeof(Barriers) + 64;
}

__global__ void __launch_bounds__(kThreads, 1)
p2p_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_cols, const P2PArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int kc = a.d / KCH;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sR = base;
  uint8_t* sC = sR + (size_t)kc * BM * 128;
  uint8_t* sG = sC + (size_t)kStages * kc * BN * 128;
  Barriers* bars = reinterpret_cast<Barriers*>(sG + 2 * BM * 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BM;
  const int split = blockIdx.y;
  const int col0 = a.col_begin + split * a.cols_per_split;
  const int col_end = min(a.n_cols, col0 + a.cols_per_split);
  const int n_tiles = (col_end - col0 + BN - 1) / BN;
  const bool bwd = a.mode != kFwd;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_rows);
    tma_prefetch_desc(&map_cols);
    mbar_init(&bars->r_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->c_full[s], 1); mbar_init(&bars->c_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->s_empty[s], kEpiWarps);
      mbar_init(&bars->g_full[s], kEpiWarps);
      mbar_init(&bars->g_empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(&bars->r_full, (uint32_t)kc * BM * 128);
      for (int c = 0; c < kc; ++c) tma_load_2d(sR + (size_t)c * BM * 128, &map_rows, &bars->r_full, c * KCH, row0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&bars->c_empty[s], ((t / kStages) & 1) ^ 1);
        mbar_expect_tx(&bars->c_full[s], (uint32_t)kc * BN * 128);
        uint8_t* dst = sC + (size_t)s * kc * BN * 128;
        for (int c = 0; c < kc; ++c) tma_load_2d(dst + (size_t)c * BN * 128, &map_cols, &bars->c_full[s], c * KCH, col0 + t * BN);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(BN, 0);
      const uint32_t idesc2 = make_idesc(a.d, 1);
      const uint32_t r_addr = smem_u32(sR), c_addr = smem_u32(sC), g_addr = smem_u32(sG);
      mbar_wait(&bars->r_full, 0);
      for (int t = 0; t <= n_tiles; ++t) {
        if (t < n_tiles) {
          const int s = t % kStages, buf = t & 1;
          mbar_wait(&bars->c_full[s], (t / kStages) & 1);
          mbar_wait(&bars->s_empty[buf], ((t >> 1) & 1) ^ 1);
          tc_fence_after();
          // S[buf] = R_tile . Cm_tile^T : both operands K-major, 128-byte swizzle, 16 bf16 (32 B) per K step
          for (int c = 0; c < kc; ++c) {
            const uint32_t ra = r_addr + (uint32_t)c * BM * 128;
            const uint32_t ca = c_addr + (uint32_t)(s * kc + c) * BN * 128;
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k)
              umma_bf16(tmem + kColS + buf * BN, make_desc(ra + k * 32, 16, 1024), make_desc(ca + k * 32, 16, 1024), idesc1,
                        (c | k) != 0);
          }
          umma_commit(&bars->s_full[buf]);
          if (!bwd) umma_commit(&bars->c_empty[s]);
        }
        if (bwd && t > 0) {
          // dR += G(t-1)[128 x 64] . Cm_tile(t-1)[64 x d] : A = G K-major; B = Cm tile as MN-major operand
          const int tp = t - 1, sp = tp % kStages, bp = tp & 1;
          mbar_wait(&bars->g_full[bp], (tp >> 1) & 1);
          tc_fence_after();
          const uint32_t ga = g_addr + (uint32_t)bp * BM * 128;
          const uint32_t ca = c_addr + (uint32_t)(sp * kc) * BN * 128;
#pragma unroll
          for (int k = 0; k < BN / 16; ++k)
            umma_bf16(tmem + kColAcc, make_desc(ga + k * 32, 16, 1024), make_desc(ca + k * 2048, BN * 128, 1024), idesc2,
                      (tp | k) != 0);
          umma_commit(&bars->g_empty[bp]);
          umma_commit(&bars->c_empty[sp]);
        }
      }
      if (bwd) umma_commit(&bars->acc_full);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: one thread per row, 32 of the 64 tile columns per warp =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;             // column half of the S tile
    const int r_local = q * 32 + lane;
    const int row = row0 + r_local;
    const bool row_ok = row < a.n_rows;
    const int2 rm = row_ok ? a.row_meta[row] : make_int2(-1, -1);
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.mode != kBwdCols && row_ok) rs = a.row_stat[row];
    float zs = 0.f, praw = 0.f, npos = 0.f;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);

    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int jbase = col0 + t * BN + half * 32;
      mbar_wait(&bars->s_full[buf], (t >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + kColS + buf * BN + half * 32, v);
      tmem_ld_wait();
      // S buffer is free as soon as it sits in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty[buf]);

      if (!bwd) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int j = jbase + jj;
          const int2 cm = (j < col_end) ? __ldg(&a.col_meta[j]) : make_int2(-2, rm.y);      // out of range == self
          const float s = __uint_as_float(v[jj]);
          const bool valid = cm.y != rm.y;
          const bool pos = valid && (cm.x == rm.x);
          const float e = exp2f(fmaf(s, a.scale_log2, -rs.x));
          zs += valid ? e : 0.f;
          praw += pos ? s : 0.f;
          npos += pos ? 1.f : 0.f;
        }
      } else {
        uint32_t packed[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          float g2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = jbase + jj + u;
            const bool in = j < col_end;
            const int2 cm = in ? __ldg(&a.col_meta[j]) : make_int2(-2, rm.y);
            float4 st = rs;
            if (a.mode == kBwdCols) st = in ? __ldg(&a.col_stat[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float s = __uint_as_float(v[jj + u]);
            const bool valid = (cm.y != rm.y) && row_ok;
            const bool pos = cm.x == rm.x;
            float g = st.y * exp2f(fmaf(s, a.scale_log2, -st.x));
            g -= pos ? st.z : 0.f;
            g2[u] = valid ? g : 0.f;
          }
          __nv_bfloat162 h = __floats2bfloat162_rn(g2[0], g2[1]);
          packed[jj >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        // G tile -> smem as a K-major, 128-byte-swizzled A operand: row r_local, 16-byte chunk (half*4 + i) ^ (r_local & 7)
        mbar_wait(&bars->g_empty[buf], ((t >> 1) & 1) ^ 1);
        uint8_t* grow = sG + (size_t)buf * BM * 128 + (size_t)r_local * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int chunk = (half * 4 + i) ^ (r_local & 7);
          *reinterpret_cast<uint4*>(grow + chunk * 16) =
              make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->g_full[buf]);
      }
    }

    if (!bwd) {
      if (row_ok) {
        float* out = a.stat_partial + ((size_t)(split * 2 + half) * a.n_rows + row) * 3;
        out[0] = zs; out[1] = praw; out[2] = npos;
      }
    } else {
      // accumulators -> global partial, 32 columns at a time; warps of the two halves split the d columns
      mbar_wait(&bars->acc_full, 0);
      tc_fence_after();
      float* out = a.grad_partial + ((size_t)split * a.n_rows + row) * a.d;
      for (int c = half * 32; c < a.d; c += 64) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + kColAcc + c, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(out + c + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                  __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------
// stats[i] = sum over slots of partial (deterministic order)
__global__ void p2p_reduce_stats_kernel(const float* partial, int n_slots, int n_rows, float* stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * 3) return;
  float t = 0.f;
  for (int s = 0; s < n_slots; ++s) t += partial[(size_t)s * n_rows * 3 + i];
  stats[i] = t;
}

// loss = sum_i w_i * (shift_i + log(Zs_i) - P_i / n_i),  P_i = P_raw_i / T     (utils/loss.py:371-386)
__global__ void __launch_bounds__(256) p2p_loss_kernel(const float* stats, const float* shift, const float* weight, int n,
                                                        float inv_t, float* loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float w = weight[i];
    const float zs = stats[3 * i], p = stats[3 * i + 1] * inv_t, np = stats[3 * i + 2];
    const float li = shift[i] + logf(zs) - p / np;          // np == 0 -> NaN, as 0/0 in the reference (:376-380)
    acc += (double)(w * li);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[0] = (float)t;
  }
}

// per-anchor backward constants {shift*log2e, alpha, beta, 0}: alpha = g w /(T Zs), beta = g w /(T n)
__global__ void p2p_anchor_stat_kernel(const float* stats, const float* shift, const float* weight, const float* grad_out,
                                       int n, float inv_t, int with_grad, float4* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 o = make_float4(shift[i] * kLog2e, 0.f, 0.f, 0.f);
  if (with_grad) {
    const float gw = grad_out[0] * weight[i] * inv_t;
    o.y = gw / stats[3 * i];
    o.z = gw / stats[3 * i + 2];
  }
  out[i] = o;
}

__global__ void p2p_reduce_grad_kernel(const float* partial, int n_splits, int64_t n_elems, int d_pad, int d, float* out) {
  // partial: [splits][rows][d_pad] -> out [rows][d]
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t rows_d = n_elems;        // rows * d
  if (idx >= rows_d) return;
  const int64_t r = idx / d, c = idx % d;
  float t = 0.f;
  const int64_t stride = (rows_d / d) * d_pad;
  for (int s = 0; s < n_splits; ++s) t += partial[(size_t)s * stride + r * d_pad + c];
  out[idx] = t;
}

// ------------------------------ host side ----------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, d] bf16 row-major, box = 64 elements (128 B) x box_rows, 128-byte swizzle, OOB rows read as zero
int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t d, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return SLCL_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return SLCL_ERR_CUDA;
  }
  return SLCL_OK;
}

struct Sweep { int row_tiles, splits, cols_per_split; };

Sweep plan_sweep(int64_t n_rows, int64_t n_cols) {
  Sweep s;
  s.row_tiles = (int)ceil_div<int64_t>(n_rows, BM);
  int col_tiles = (int)ceil_div<int64_t>(n_cols, BN);
  int want = max(1, sm_count() / s.row_tiles);      // fill the SMs: row tiles x column splits ~ #SMs
  s.splits = min(want, col_tiles);
  int tiles_per_split = ceil_div(col_tiles, s.splits);
  s.splits = ceil_div(col_tiles, tiles_per_split);
  s.cols_per_split = tiles_per_split * BN;
  return s;
}

int launch_sweep(const void* rows, int64_t n_rows, const void* cols, int64_t n_cols, int d, int mode, float inv_t,
                 const int2* row_meta, const int2* col_meta, const float4* row_stat, const float4* col_stat,
                 float* stat_partial, float* grad_partial, const Sweep& sw, cudaStream_t stream) {
  CUtensorMap mr, mc;
  int st = make_map(&mr, rows, n_rows, d, BM);
  if (st != SLCL_OK) return st;
  st = make_map(&mc, cols, n_cols, d, BN);
  if (st != SLCL_OK) return st;
  P2PArgs a{};
  a.n_rows = (int)n_rows; a.n_cols = (int)n_cols; a.d = d;
  a.col_begin = 0; a.cols_per_split = sw.cols_per_split;
  a.mode = mode;
  a.scale_log2 = inv_t * kLog2e;
  a.row_meta = row_meta; a.col_meta = col_meta; a.row_stat = row_stat; a.col_stat = col_stat;
  a.stat_partial = stat_partial; a.grad_partial = grad_partial;
  const size_t smem = smem_bytes_for(d);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for(kMaxD));
    if (e != cudaSuccess) { set_cuda_error(e, "cudaFuncSetAttribute(p2p_kernel)"); return SLCL_ERR_CUDA; }
    attr_set = true;
  }
  p2p_kernel<<<dim3(sw.row_tiles, sw.splits), kThreads, smem, stream>>>(mr, mc, a);
  return check_launch("p2p_kernel");
}

// workspace layout helpers
struct P2PWs {
  float* stat_partial;     // [2*splits_a][Na][3]
  float4* anchor_stat;     // [Na]
  float* grad_partial_a;   // [splits_a][Na][d]
  float* grad_partial_b;   // [splits_b][M][d]
  size_t total;
};

P2PWs carve(void* ws, int64_t na, int64_t m, int d) {
  Sweep sa = plan_sweep(na, m), sb = plan_sweep(m, na);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  size_t o1 = take((size_t)2 * sa.splits * na * 3 * sizeof(float));
  size_t o2 = take((size_t)na * sizeof(float4));
  size_t o3 = take((size_t)sa.splits * na * d * sizeof(float));
  size_t o4 = take((size_t)sb.splits * m * d * sizeof(float));
  P2PWs w;
  char* b = reinterpret_cast<char*>(ws);
  w.stat_partial = reinterpret_cast<float*>(b + o1);
  w.anchor_stat = reinterpret_cast<float4*>(b + o2);
  w.grad_partial_a = reinterpret_cast<float*>(b + o3);
  w.grad_partial_b = reinterpret_cast<float*>(b + o4);
  w.total = off;
  return w;
}

bool p2p_args_ok(const void* a, const void* b, int64_t na, int64_t m, int64_t d_pad) {
  return a && b && na > 0 && m > 0 && d_pad >= KCH && d_pad <= kMaxD && d_pad % KCH == 0 && aligned16(a) && aligned16(b) &&
         na < (1ll << 31) && m < (1ll << 31);
}

}  // namespace
}  // namespace slcl

using namespace slcl;

extern "C" size_t slcl_p2p_workspace_bytes(int64_t n_anchor, int64_t n_contrast, int64_t dim_padded) {
  if (n_anchor <= 0 || n_contrast <= 0 || dim_padded < KCH || dim_padded > kMaxD || dim_padded % KCH) return 0;
  return carve(nullptr, n_anchor, n_contrast, (int)dim_padded).total;
}

extern "C" int slcl_p2p_fwd(const void* a_bf16, const void* b_bf16, int64_t n_anchor, int64_t n_contrast, int64_t dim_padded,
                            const int32_t* a_meta, const int32_t* b_meta, const float* shift, const float* weight,
                            float temperature, float* stats, float* loss, void* workspace, size_t workspace_bytes,
                            slcl_stream_t stream_) {
  if (!p2p_args_ok(a_bf16, b_bf16, n_anchor, n_contrast, dim_padded) || !a_meta || !b_meta || !shift || !weight || !stats ||
      !loss || !workspace || !(temperature > 0.f))
    return SLCL_ERR_INVALID_ARGUMENT;
  const int d = (int)dim_padded;
  if (workspace_bytes < slcl_p2p_workspace_bytes(n_anchor, n_contrast, dim_padded) || !aligned16(workspace))
    return SLCL_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  P2PWs w = carve(workspace, n_anchor, n_contrast, d);
  const float inv_t = 1.0f / temperature;
  const int na = (int)n_anchor;
  p2p_anchor_stat_kernel<<<ceil_div(na, 256), 256, 0, stream>>>(nullptr, shift, weight, nullptr, na, inv_t, 0, w.anchor_stat);
  Sweep sw = plan_sweep(n_anchor, n_contrast);
  int st = launch_sweep(a_bf16, n_anchor, b_bf16, n_contrast, d, kFwd, inv_t, reinterpret_cast<const int2*>(a_meta),
                        reinterpret_cast<const int2*>(b_meta), w.anchor_stat, nullptr, w.stat_partial, nullptr, sw, stream);
  if (st != SLCL_OK) return st;
  p2p_reduce_stats_kernel<<<ceil_div(na * 3, 256), 256, 0, stream>>>(w.stat_partial, 2 * sw.splits, na, stats);
  p2p_loss_kernel<<<1, 256, 0, stream>>>(stats, shift, weight, na, inv_t, loss);
  return check_launch("slcl_p2p_fwd");
}

extern "C" int slcl_p2p_bwd(const void* a_bf16, const void* b_bf16, int64_t n_anchor, int64_t n_contrast, int64_t dim_padded,
                            int64_t dim, const int32_t* a_meta, const int32_t* b_meta, const float* shift,
                            const float* weight, float temperature, const float* stats, const float* grad_out, float* d_a,
                            float* d_b, void* workspace, size_t workspace_bytes, slcl_stream_t stream_) {
  if (!p2p_args_ok(a_bf16, b_bf16, n_anchor, n_contrast, dim_padded) || !a_meta || !b_meta || !shift || !weight || !stats ||
      !grad_out || !workspace || !(temperature > 0.f) || dim <= 0 || dim > dim_padded || (!d_a && !d_b))
    return SLCL_ERR_INVALID_ARGUMENT;
  const int d = (int)dim_padded;
  if (workspace_bytes < slcl_p2p_workspace_bytes(n_anchor, n_contrast, dim_padded) || !aligned16(workspace))
    return SLCL_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  P2PWs w = carve(workspace, n_anchor, n_contrast, d);
  const float inv_t = 1.0f / temperature;
  const int na = (int)n_anchor;
  p2p_anchor_stat_kernel<<<ceil_div(na, 256), 256, 0, stream>>>(stats, shift, weight, grad_out, na, inv_t, 1, w.anchor_stat);
  const int2* am = reinterpret_cast<const int2*>(a_meta);
  const int2* bm = reinterpret_cast<const int2*>(b_meta);
  if (d_a) {
    Sweep sw = plan_sweep(n_anchor, n_contrast);
    int st = launch_sweep(a_bf16, n_anchor, b_bf16, n_contrast, d, kBwdRows, inv_t, am, bm, w.anchor_stat, nullptr, nullptr,
                          w.grad_partial_a, sw, stream);
    if (st != SLCL_OK) return st;
    const int64_t n = n_anchor * dim;
    p2p_reduce_grad_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(w.grad_partial_a, sw.splits, n, d, (int)dim, d_a);
  }
  if (d_b) {
    Sweep sw = plan_sweep(n_contrast, n_anchor);
    int st = launch_sweep(b_bf16, n_contrast, a_bf16, n_anchor, d, kBwdCols, inv_t, bm, am, nullptr, w.anchor_stat, nullptr,
                          w.grad_partial_b, sw, stream);
    if (st != SLCL_OK) return st;
    const int64_t n = n_contrast * dim;
    p2p_reduce_grad_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, stream>>>(w.grad_partial_b, sw.splits, n, d, (int)dim, d_b);
  }
  return check_launch("slcl_p2p_bwd");
}
