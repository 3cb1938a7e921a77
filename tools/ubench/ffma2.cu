// Micro-benchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput per SM (B200).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096, kU = 8;
template <int KIND>
__global__ void k(float* out, float s0, float s1) {
  float a[kU * 2];
#pragma unroll
  for (int u = 0; u < kU * 2; ++u) a[u] = s0 + u + threadIdx.x * 1e-3f;
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (KIND == 0) { a[2 * u] = fmaf(a[2 * u], s1, s0); a[2 * u + 1] = fmaf(a[2 * u + 1], s1, s0); }
      if (KIND == 1) {
        unsigned long long d, x, y, z;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[2 * u]), "f"(a[2 * u + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(s1));
        asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(s0));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * u]), "=f"(a[2 * u + 1]) : "l"(d));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kU * 2; ++u) s += a[u];
  if (s == 123.456f) out[0] = s;
}
template <int KIND> void run(const char* name, int warps) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms, warps * 32>>>(out, 0.5f, 0.999f);
  cudaEventRecord(e0); k<KIND><<<sms, warps * 32>>>(out, 0.5f, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double fmas = (double)kIters * kU * 2 * warps * 32;
  printf("%-10s warps/SM %2d: %.3f ms  %.1f FMA/clk/SM\n", name, warps, ms, fmas / (ms * 1e-3 * clk * 1e3));
}
int main() { for (int w : {8, 16, 32}) { run<0>("ffma", w); run<1>("ffma2", w); } return 0; }
