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
#include <limits.h>
#include <type_traits>
#include <stdlib.h>
#include <math.h>

namespace slcl {
namespace {

constexpr int BM = 128;            // rows per CTA (UMMA M)
constexpr int BN = 64;             // streamed rows (S tile columns) per step
constexpr int KCH = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kStages = 5;
constexpr int kMaxD = 256;
constexpr int kEpiWarps = 8;       // warps 4..11; (warp % 4) selects the TMEM lane quarter
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr int kTmemCols = 512;
constexpr int kColS = 0;           // S double buffer: columns [0,64) and [64,128)
constexpr int kColAcc = 128;       // dR accumulators: columns [128, 128 + d)
constexpr int kColR = 384;         // resident row operand R (bf16 pairs): columns [384, 384 + d/2)
constexpr float kLog2e = 1.4426950408889634f;

#ifdef SLCL_P2P_PROFILE
constexpr bool kProfile = true;
#else
constexpr bool kProfile = false;
#endif

enum Mode { kFwd = 0, kBwdRows = 1 /* dA: stats per row */, kBwdCols = 2 /* dB: stats per column */ };

struct P2PArgs {
  int n_rows, n_cols, d;              // d padded to a multiple of 64
  int col_begin, cols_per_split;      // this kernel instance sweeps columns [col_begin + split*cols_per_split, ...)
  int mode;
  float scale_log2;                   // log2(e) / T
  const uint32_t* rows_u32;           // resident operand, bf16 row-major [n_rows, d] viewed as 32-bit words
  const int2* row_meta;               // {label, id}
  const int2* col_meta;
  const float4* row_stat;             // kFwd: {shift*log2e,-,-,-}; kBwdRows: {shift*log2e, alpha, beta, -}
  const float4* col_stat;             // kBwdCols: {shift*log2e, alpha, beta, -}
  float* stat_partial;                // kFwd: [n_slots][n_rows][3]  (Zs, P_raw, n)
  float* grad_partial;                // bwd:  [n_splits][n_rows][d] fp32
  unsigned long long* prof;           // bring-up: per-role wait-cycle counters of CTA (0,0), or null
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
// Bring-up instrumentation (per-role barrier-wait cycle counters, read by tools/p2p_prof.py) is compiled
// in only with -DSLCL_P2P_PROFILE; the shipped library waits without touching the clock.
#ifdef SLCL_P2P_PROFILE
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long& acc) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}
#define SLCL_PROF_NOW() clock64()
#else
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long&) { mbar_wait(bar, parity); }
#define SLCL_PROF_NOW() 0ll
#endif
// one lane of a converged warp (elect.sync): the idiom the compiler recognises for single-thread
// issue of uniform-datapath instructions
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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
// same, multicast: the box lands at the same CTA-relative offset (and signals the same CTA-relative
// mbarrier) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
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
// same with the A operand read from tensor memory (rows = lanes, bf16 pairs packed along K in 32-bit columns):
// no per-instruction shared-memory fetch of the 128 A rows, so a narrow-N MMA runs at its N/2-cycle floor
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same, arriving on the barrier at this CTA-relative offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// Column metadata of one 64-column tile.  The TMA producer brings it in with 1-D bulk copies into a
// kMetaSlots-deep ring of its own (the epilogue reads it after the tile's smem stage has already been
// handed back), so the epilogue warps never issue a global load.  The caller pads the metadata arrays
// to a multiple of 64 entries; pad entries carry id INT_MIN (masked).
constexpr int kMetaSlots = 4;
struct __align__(128) ColMeta {
  float4 stat[BN];      // kBwdCols: {shift*log2e, alpha, beta, -}
  int2 meta[BN];        // {label, id}
};

struct __align__(8) Barriers {
  uint64_t r_full;
  uint64_t c_full[kStages], c_empty[kStages];
  uint64_t s_full[2], s_empty[2];
  uint64_t g_full[2], g_empty[2];
  uint64_t acc_full;
  uint64_t m_full[kMetaSlots], m_empty[kMetaSlots];
  uint32_t tmem_base;
};

// dynamic smem carve-up (1024-byte aligned tiles)
//   (the resident row operand R lives in tensor memory, not here)
//   Cm  : [kStages][d/64][64 rows][128 B]
//   G   : [2][128 rows][128 B]
__host__ __device__ inline size_t smem_bytes_for(int d) {
  const size_t kc = d / KCH;
  return 1024 /*align slack*/ + (size_t)kStages * kc * BN * 128 + 2 * BM * 128 + kMetaSlots * sizeof(ColMeta) +
         sizeof(Barriers) + 64;
}

// CS = thread-block-cluster size along the row-tile axis.  The CS CTAs of a cluster sweep the same
// column tiles in lock step: each one fetches 1/CS of every column tile and TMA-multicasts it to all,
// so the L2 -> SM traffic of the streamed operand drops CS-fold.
template <int CS>
__global__ void __launch_bounds__(kThreads, 1)
p2p_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_cols, const P2PArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int kc = a.d / KCH;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sC = base;
  uint8_t* sG = sC + (size_t)kStages * kc * BN * 128;
  ColMeta* sMeta = reinterpret_cast<ColMeta*>(sG + 2 * BM * 128);
  Barriers* bars = reinterpret_cast<Barriers*>(sMeta + kMetaSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BM;
  const int split = blockIdx.y;
  const int col0 = a.col_begin + split * a.cols_per_split;
  const int col_end = min(a.n_cols, col0 + a.cols_per_split);
  const int n_tiles = (col_end - col0 + BN - 1) / BN;
  const bool bwd = a.mode != kFwd;
  // Every CTA sweeps the same column tiles; start each one at a different tile so the CTAs do not
  // all hit the same L2 lines at the same moment (the sums do not depend on the sweep order).
  const int rot = (int)(((blockIdx.x / CS) * 37u + blockIdx.y * 11u) % (unsigned)n_tiles);
  const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CS) - 1u);
  auto tile_col = [&](int t) { int tt = t + rot; if (tt >= n_tiles) tt -= n_tiles; return col0 + tt * BN; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_cols);
    mbar_init(&bars->r_full, 4);                 // the four warps that fill R's lane quarters
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->c_full[s], 1); mbar_init(&bars->c_empty[s], CS); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->s_empty[s], kEpiWarps);
      mbar_init(&bars->g_full[s], kEpiWarps);
      mbar_init(&bars->g_empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    for (int s = 0; s < kMetaSlots; ++s) { mbar_init(&bars->m_full[s], 1); mbar_init(&bars->m_empty[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();          // every CTA's barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // Producer and MMA roles run with the WHOLE warp in the loop (converged, warp-uniform control flow) and
  // elect one lane only around the asynchronous instructions.  Running them under `if (lane == 0)` makes
  // the compiler wrap every UTCHMMA / UTMALDG in an ELECT + R2UR waterfall loop (~100 cycles per MMA).
  if (warp == 0) {
    // ===================== TMA producer =====================
    unsigned long long w0 = 0, w1 = 0;
    const long long tstart = SLCL_PROF_NOW();
    const bool with_stat = a.mode == kBwdCols;
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kStages;
      const int ms = t % kMetaSlots;
      mbar_wait_t(&bars->c_empty[s], ((t / kStages) & 1) ^ 1, w0);
      mbar_wait_t(&bars->m_empty[ms], ((t / kMetaSlots) & 1) ^ 1, w1);
      if (elect_one()) {
        mbar_expect_tx(&bars->c_full[s], (uint32_t)kc * BN * 128);
        uint8_t* dst = sC + (size_t)s * kc * BN * 128;
        if (CS == 1) {
          for (int c = 0; c < kc; ++c) tma_load_2d(dst + (size_t)c * BN * 128, &map_cols, &bars->c_full[s], c * KCH, tile_col(t));
        } else {
          constexpr int kSlice = BN / CS;        // rows of the tile this CTA fetches for the whole cluster
          for (int c = 0; c < kc; ++c)
            tma_load_2d_mc(dst + (size_t)c * BN * 128 + (size_t)crank * kSlice * 128, &map_cols, &bars->c_full[s], c * KCH,
                           tile_col(t) + (int)crank * kSlice, kAllCtas);
        }
        mbar_expect_tx(&bars->m_full[ms], (uint32_t)(BN * sizeof(int2) + (with_stat ? BN * sizeof(float4) : 0)));
        bulk_load_1d(sMeta[ms].meta, a.col_meta + tile_col(t), BN * sizeof(int2), &bars->m_full[ms]);
        if (with_stat) bulk_load_1d(sMeta[ms].stat, a.col_stat + tile_col(t), BN * sizeof(float4), &bars->m_full[ms]);
      }
      __syncwarp();
    }
    if (kProfile && a.prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
      a.prof[0] = (unsigned long long)(SLCL_PROF_NOW() - tstart); a.prof[1] = w0; a.prof[2] = w1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc1 = make_idesc(BN, 0);
    const uint32_t idesc2 = make_idesc(a.d, 1);
    const uint32_t c_addr = smem_u32(sC), g_addr = smem_u32(sG);
    unsigned long long w0 = 0, w1 = 0, w2 = 0;
    const long long tstart = SLCL_PROF_NOW();
    mbar_wait(&bars->r_full, 0);
    tc_fence_after();
    // Descriptors are built once; inside the loops only the 14-bit start-address field moves
    // (all operand addresses are < 256 KB, so adding (bytes >> 4) to the low word never carries out).
    const uint64_t descC = make_desc(c_addr, 16, 1024);              // K-major view of a column tile (MMA1)
    const uint64_t descCmn = make_desc(c_addr, BN * 128, 1024);      // MN-major view of the same bytes (MMA2)
    const uint64_t descG = make_desc(g_addr, 16, 1024);
    const uint32_t stage_units = (uint32_t)(kc * BN * 128) >> 4;
    for (int t = 0; t <= n_tiles; ++t) {
      if (t < n_tiles) {
        const int s = t % kStages, buf = t & 1;
        mbar_wait_t(&bars->c_full[s], (t / kStages) & 1, w0);
        mbar_wait_t(&bars->s_empty[buf], ((t >> 1) & 1) ^ 1, w1);
        tc_fence_after();
        if (elect_one()) {
          // S[buf] = R . Cm_tile^T : A = R from tensor memory, B = column tile K-major, 16 bf16 of K per MMA
          const uint32_t d_tmem = tmem + kColS + buf * BN;
          uint32_t ta = tmem + kColR;                          // 16 bf16 of K = 8 tensor-memory columns per step
          uint64_t db = descC + (uint64_t)(s * stage_units);
          for (int c = 0; c < kc; ++c) {
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k)
              umma_bf16_ts(d_tmem, ta + 8 * k, db + 2 * k, idesc1, (c | k) != 0);
            ta += KCH / 2;
            db += (BN * 128) >> 4;
          }
          umma_commit(&bars->s_full[buf]);
          if (!bwd) { if (CS == 1) umma_commit(&bars->c_empty[s]); else umma_commit_mc(&bars->c_empty[s], kAllCtas); }
        }
        __syncwarp();
      }
      if (bwd && t > 0) {
        // dR += G(t-1)[128 x 64] . Cm_tile(t-1)[64 x d] : A = G K-major (smem); B = Cm tile as MN-major operand
        const int tp = t - 1, sp = tp % kStages, bp = tp & 1;
        mbar_wait_t(&bars->g_full[bp], (tp >> 1) & 1, w2);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dg = descG + (uint64_t)(bp * ((BM * 128) >> 4));
          const uint64_t dc = descCmn + (uint64_t)(sp * stage_units);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) umma_bf16(tmem + kColAcc, dg + 2 * k, dc + k * (2048 >> 4), idesc2, (tp | k) != 0);
          umma_commit(&bars->g_empty[bp]);
          if (CS == 1) umma_commit(&bars->c_empty[sp]); else umma_commit_mc(&bars->c_empty[sp], kAllCtas);
        }
        __syncwarp();
      }
    }
    if (bwd && elect_one()) umma_commit(&bars->acc_full);
    __syncwarp();
    if (kProfile && a.prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
      a.prof[4] = (unsigned long long)(SLCL_PROF_NOW() - tstart); a.prof[5] = w0; a.prof[6] = w1; a.prof[7] = w2;
    }
  } else if (warp >= 4) {
    // ===================== epilogue: one thread per row, 32 of the 64 tile columns per warp =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;             // column half of the S tile
    const int r_local = q * 32 + lane;
    const int row = row0 + r_local;
    const bool row_ok = row < a.n_rows;
    const int2 rm = row_ok ? a.row_meta[row] : make_int2(INT_MIN + 1, INT_MIN + 1);
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.mode != kBwdCols && row_ok) rs = a.row_stat[row];
    float zs = 0.f, praw = 0.f, npos = 0.f;
    unsigned long long w0 = 0, w1 = 0, w2 = 0;
    const uint32_t lane_addr0 = ((uint32_t)(q * 32) << 16);
    if (half == 0) {
      // resident operand: this thread's row (bf16 pairs, 32-bit words) -> tensor memory lane r_local
      const uint4* src = reinterpret_cast<const uint4*>(a.rows_u32 + (size_t)row * (a.d / 2));
      for (int c = 0; c < a.d / 2; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 t = row_ok ? __ldg(src + (c >> 2) + i) : make_uint4(0u, 0u, 0u, 0u);
          v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
        tmem_st32(tmem + lane_addr0 + kColR + c, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->r_full);
    }
    const long long tstart = SLCL_PROF_NOW();
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int ms = t % kMetaSlots;
      const ColMeta& cmeta = sMeta[ms];
      mbar_wait_t(&bars->m_full[ms], (t / kMetaSlots) & 1, w0);
      mbar_wait_t(&bars->s_full[buf], (t >> 1) & 1, w1);
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
          const int2 cm = cmeta.meta[half * 32 + jj];
          const float s = __uint_as_float(v[jj]);
          const float e = ex2_approx(fmaf(s, a.scale_log2, -rs.x));
          // predicated adds spelled out in PTX: 2 compares + 3 predicated FADDs, no selects, no branches
          asm("{\n\t"
              ".reg .pred pv, pp;\n\t"
              "setp.ne.s32 pv, %5, %6;\n\t"
              "setp.eq.and.s32 pp, %7, %8, pv;\n\t"
              "@pv add.f32 %0, %0, %3;\n\t"
              "@pp add.f32 %1, %1, %4;\n\t"
              "@pp add.f32 %2, %2, 0f3F800000;\n\t"
              "}\n"
              : "+f"(zs), "+f"(praw), "+f"(npos)
              : "f"(e), "f"(s), "r"(cm.y), "r"(rm.y), "r"(cm.x), "r"(rm.x));
        }
        // Padding columns (beyond col_end; only the last tile of a sweep has them) were read as zero rows
        // by TMA, carry the sentinel id/label and therefore entered zs as exp(0 - shift): take them out
        // analytically instead of testing every element.
        {
          const int jb = tile_col(t) + half * 32;
          const int n_pad = max(0, min(32, jb + 32 - col_end));
          if (n_pad > 0) zs -= (float)n_pad * ex2_approx(-rs.x);
        }
      } else {
        uint32_t packed[16];
        // the statistics {shift, alpha, beta} belong to the row (dA sweep) or to the column (dB sweep):
        // decide once per tile, not per element
        auto make_g = [&](auto col_stat) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            float g2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int2 cm = cmeta.meta[half * 32 + jj + u];
              float sh = rs.x, al = rs.y, be = rs.z;
              if constexpr (decltype(col_stat)::value) {
                const float4 st = cmeta.stat[half * 32 + jj + u];
                sh = st.x; al = st.y; be = st.z;
              }
              const float s = __uint_as_float(v[jj + u]);
              const float e = ex2_approx(fmaf(s, a.scale_log2, -sh));
              float g;      // alpha*e - [same label] beta, zero for the self pair
              asm("{\n\t"
                  ".reg .pred pv, pp;\n\t"
                  "setp.ne.s32 pv, %4, %5;\n\t"
                  "setp.eq.s32 pp, %6, %7;\n\t"
                  "mul.f32 %0, %1, %2;\n\t"
                  "@pp sub.f32 %0, %0, %3;\n\t"
                  "@!pv mov.f32 %0, 0f00000000;\n\t"
                  "}\n"
                  : "=&f"(g)
                  : "f"(al), "f"(e), "f"(be), "r"(cm.y), "r"(rm.y), "r"(cm.x), "r"(rm.x));
              g2[u] = g;
            }
            __nv_bfloat162 h = __floats2bfloat162_rn(g2[0], g2[1]);
            packed[jj >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
        };
        if (a.mode == kBwdCols) make_g(std::true_type{}); else make_g(std::false_type{});
        // G tile -> smem as a K-major, 128-byte-swizzled A operand: row r_local, 16-byte chunk (half*4 + i) ^ (r_local & 7)
        mbar_wait_t(&bars->g_empty[buf], ((t >> 1) & 1) ^ 1, w2);
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->m_empty[ms]);
    }

    if (kProfile && a.prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 4 || warp == 11)) {
      unsigned long long* pp = a.prof + (warp == 4 ? 8 : 12);
      pp[0] = (unsigned long long)(SLCL_PROF_NOW() - tstart); pp[1] = w0; pp[2] = w1; pp[3] = w2;
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
  if (CS > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------
// stats[i] = sum over slots of partial (deterministic order); also the per-block partial of
// sum_i w_i * (shift_i + log(Zs_i) - P_i / n_i),  P_i = P_raw_i / T     (utils/loss.py:371-386)
__global__ void __launch_bounds__(256) p2p_reduce_stats_kernel(const float* partial, int n_slots, int n_rows, const float* shift,
                                                               const float* weight, float inv_t, float* stats,
                                                               double* loss_partial) {
  __shared__ double red[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double acc = 0.0;
  if (i < n_rows) {
    float t[3] = {0.f, 0.f, 0.f};
    for (int s = 0; s < n_slots; ++s) {
      const float* p = partial + ((size_t)s * n_rows + i) * 3;
      t[0] += p[0]; t[1] += p[1]; t[2] += p[2];
    }
    stats[3 * i] = t[0]; stats[3 * i + 1] = t[1]; stats[3 * i + 2] = t[2];
    const float li = shift[i] + logf(t[0]) - (t[1] * inv_t) / t[2];     // n == 0 -> NaN, as 0/0 in the reference (:376-380)
    acc = (double)(weight[i] * li);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss_partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) p2p_loss_kernel(const double* loss_partial, int n_blocks, float* loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_blocks; i += 256) acc += loss_partial[i];
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
                                       int n, int n_padded, float inv_t, int with_grad, float4* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  if (i >= n) { out[i] = make_float4(0.f, 0.f, 0.f, 0.f); return; }      // pad entries (bulk-copied as column stats)
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

struct Sweep { int row_tiles, splits, cols_per_split, cluster; };

Sweep plan_sweep(int64_t n_rows, int64_t n_cols) {
  Sweep s;
  s.row_tiles = (int)ceil_div<int64_t>(n_rows, BM);
  int col_tiles = (int)ceil_div<int64_t>(n_cols, BN);
  int want = max(1, sm_count() / s.row_tiles);      // fill the SMs: row tiles x column splits ~ #SMs
  s.splits = min(want, col_tiles);
  int tiles_per_split = ceil_div(col_tiles, s.splits);
  s.splits = ceil_div(col_tiles, tiles_per_split);
  s.cols_per_split = tiles_per_split * BN;
  // Cluster size of the multicast variant.  Measured on B200 (cfg3 and 16384^2): clusters of 2 and 4 bring no
  // gain -- the sweep is MMA-issue/epilogue-bound, not L2-bound -- so the default is 1; SLCL_P2P_CLUSTER selects
  // 2 or 4 for experiments on other shapes.
  s.cluster = 1;
  { const char* e = getenv("SLCL_P2P_CLUSTER"); if (e) s.cluster = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 1); }
  return s;
}

int launch_sweep(const void* rows, int64_t n_rows, const void* cols, int64_t n_cols, int d, int mode, float inv_t,
                 const int2* row_meta, const int2* col_meta, const float4* row_stat, const float4* col_stat,
                 float* stat_partial, float* grad_partial, const Sweep& sw, cudaStream_t stream) {
  CUtensorMap mr, mc;
  int st = make_map(&mr, rows, n_rows, d, BM);
  if (st != SLCL_OK) return st;
  st = make_map(&mc, cols, n_cols, d, BN / sw.cluster);
  if (st != SLCL_OK) return st;
  P2PArgs a{};
  a.n_rows = (int)n_rows; a.n_cols = (int)n_cols; a.d = d;
  a.col_begin = 0; a.cols_per_split = sw.cols_per_split;
  a.mode = mode;
#ifdef SLCL_P2P_PROFILE
  { const char* e = getenv("SLCL_P2P_PROF"); a.prof = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  a.scale_log2 = inv_t * kLog2e;
  a.rows_u32 = reinterpret_cast<const uint32_t*>(rows);
  a.row_meta = row_meta; a.col_meta = col_meta; a.row_stat = row_stat; a.col_stat = col_stat;
  a.stat_partial = stat_partial; a.grad_partial = grad_partial;
  const size_t smem = smem_bytes_for(d);
  static bool attr_set = false;
  if (!attr_set) {
    const int mx = (int)smem_bytes_for(kMaxD);
    cudaError_t e = cudaFuncSetAttribute(p2p_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(p2p_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(p2p_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaFuncSetAttribute(p2p_kernel)"); return SLCL_ERR_CUDA; }
    attr_set = true;
  }
  const int cs = sw.cluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ceil_div(sw.row_tiles, cs) * cs), (unsigned)sw.splits, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le;
  if (cs == 4) le = cudaLaunchKernelEx(&cfg, p2p_kernel<4>, mr, mc, a);
  else if (cs == 2) le = cudaLaunchKernelEx(&cfg, p2p_kernel<2>, mr, mc, a);
  else le = cudaLaunchKernelEx(&cfg, p2p_kernel<1>, mr, mc, a);
  if (le != cudaSuccess) { set_cuda_error(le, "cudaLaunchKernelEx(p2p_kernel)"); return SLCL_ERR_CUDA; }
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
  size_t o2 = take((size_t)align_up((size_t)na, BN) * sizeof(float4));
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
  p2p_anchor_stat_kernel<<<ceil_div(na + BN, 256), 256, 0, stream>>>(nullptr, shift, weight, nullptr, na, (int)align_up((size_t)na, BN), inv_t, 0,
                                                                     w.anchor_stat);
  Sweep sw = plan_sweep(n_anchor, n_contrast);
  int st = launch_sweep(a_bf16, n_anchor, b_bf16, n_contrast, d, kFwd, inv_t, reinterpret_cast<const int2*>(a_meta),
                        reinterpret_cast<const int2*>(b_meta), w.anchor_stat, nullptr, w.stat_partial, nullptr, sw, stream);
  if (st != SLCL_OK) return st;
  // the forward no longer needs grad_partial_a: its head doubles as the per-block loss partials
  double* loss_partial = reinterpret_cast<double*>(w.grad_partial_a);
  const int nb = ceil_div(na, 256);
  p2p_reduce_stats_kernel<<<nb, 256, 0, stream>>>(w.stat_partial, 2 * sw.splits, na, shift, weight, inv_t, stats, loss_partial);
  p2p_loss_kernel<<<1, 256, 0, stream>>>(loss_partial, nb, loss);
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
  p2p_anchor_stat_kernel<<<ceil_div(na + BN, 256), 256, 0, stream>>>(stats, shift, weight, grad_out, na, (int)align_up((size_t)na, BN), inv_t, 1,
                                                                     w.anchor_stat);
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
