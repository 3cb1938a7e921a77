// common.cuh -- shared device/host helpers for libslcl (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "slcl.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libslcl is written for sm_100a (B200) only"
#endif

namespace slcl {

constexpr int kMaxK = SLCL_MAX_CLASSES;
constexpr int kNumSMsDefault = 148;

// thread-local text of the last CUDA launch error (slcl_last_cuda_error()).
void set_cuda_error(cudaError_t e, const char* where);
int  check_launch(const char* where);     // returns SLCL_OK or SLCL_ERR_CUDA
int  sm_count();                          // cached cudaDevAttrMultiProcessorCount of the current device
int  current_device_slot();               // current CUDA device index clamped to [0, 64): index of per-device caches
// Driver-API entry points (cuTensorMapEncodeTiled) need a context that is CURRENT ON THE CALLING THREAD; a thread whose
// first CUDA activity is one of our calls (PyTorch's autograd worker running a backward) has none until a runtime call
// binds the primary context.  Call this before any driver-API call.
void ensure_context_on_this_thread();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming (read-once / write-once) global accesses: evict-first so a 1 GB
// feature map does not wash the small reused state (centres, stash) out of L2.
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float  ld_stream1(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }

// L2 residency hints.  A feature map that fits in the 126 MB L2 is read twice by consecutive kernels of one step
// (class sums -> loss forward -> loss backward): the first reader loads it with evict_last so the later readers hit L2
// instead of HBM; maps larger than L2 keep the evict_first streaming behaviour.  One code path: the policy is a
// run-time 64-bit operand of ld.global.L2::cache_hint.
constexpr size_t kL2KeepBytes = 104u << 20;       // largest map we try to keep resident (cfg4 per-GPU map: 98 MiB)
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
  uint64_t pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_policy4(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_policy1(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while the kernel in front of it in the
// stream is still running; it signals its own dependents right away (pdl_trigger) and waits for the kernel in front to
// complete (pdl_wait) BEFORE it touches global memory.  Launch latency and block scheduling overlap the predecessor's tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);        // errors surface through check_launch()
}

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace slcl
