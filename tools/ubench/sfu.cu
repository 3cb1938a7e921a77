// Micro-benchmark: per-SM throughput of the epilogue instruction mix candidates (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu sfu.cu && ./sfu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kUnroll = 8;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2b2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int KIND>
__global__ void k(float* out, float seed, float scale, float shift) {
  float acc[kUnroll];
  unsigned uacc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) { acc[u] = seed + u + threadIdx.x * 1e-3f; uacc[u] = __float_as_uint(acc[u]); }
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (KIND == 0) acc[u] = ex2f(acc[u]);                                   // MUFU f32 only
      if (KIND == 1) uacc[u] = ex2h2(uacc[u]);                                // MUFU f16x2 only
      if (KIND == 2) uacc[u] = ex2b2(uacc[u]);                                // MUFU bf16x2 only
      if (KIND == 3) acc[u] = fmaf(acc[u], scale, shift);                     // FFMA 3-reg only
      if (KIND == 4) acc[u] = acc[u] + ex2f(fmaf(acc[u], scale, -shift));     // fwd epilogue: FFMA + MUFU + FADD
      if (KIND == 5) {                                                         // polynomial exp2 on the FMA/ALU pipes
        float x = fmaf(acc[u], scale, -shift);
        float t = x + 12582912.f;
        float f = x - (t - 12582912.f);
        float p = fmaf(fmaf(fmaf(0.0555f, f, 0.2402f), f, 0.6931f), f, 1.0f);
        acc[u] = __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
      }
      if (KIND == 6) {                                                         // packed: 2 FFMA + cvt + MUFU f16x2 (+ keep dependency)
        float x0 = fmaf(acc[u], scale, -shift), x1 = fmaf(acc[u], shift, -scale);
        __half2 h = __floats2half2_rn(x0, x1);
        unsigned e = ex2h2(*reinterpret_cast<unsigned*>(&h));
        acc[u] = __uint_as_float(e);
      }
      if (KIND == 7) acc[u] = acc[u] + scale;                                  // FADD
    }
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) s += acc[u] + __uint_as_float(uacc[u]);
  if (s == 123.456f) out[0] = s;
}

template <int KIND>
void run(const char* name, int warps_per_sm, float ops_per_iter) {
  int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms, warps_per_sm * 32>>>(out, 0.5f, 0.999f, 0.001f);
  cudaEventRecord(e0);
  k<KIND><<<sms, warps_per_sm * 32>>>(out, 0.5f, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double cycles = ms * 1e-3 * clk * 1e3;
  double lane_ops = (double)kIters * kUnroll * warps_per_sm * 32 * ops_per_iter;
  printf("%-34s warps/SM %2d: %.3f ms  %.1f lane-results/clk/SM (at %d MHz nominal)\n", name, warps_per_sm, ms, lane_ops / cycles, clk / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("ex2.f32", w, 1);
    run<1>("ex2.f16x2 (2 results/op)", w, 2);
    run<2>("ex2.bf16x2 (2 results/op)", w, 2);
    run<3>("ffma 3-reg", w, 1);
    run<7>("fadd", w, 1);
    run<4>("ffma+ex2.f32+fadd (per element)", w, 1);
    run<5>("poly exp2 (per element)", w, 1);
    run<6>("2ffma+cvt+ex2.f16x2 (2 elements)", w, 2);
  }
  return 0;
}
