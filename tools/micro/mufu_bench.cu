// Microbenchmark: per-SM throughput of MUFU ex2 variants on sm_100a (f32, f16x2, bf16x2) and of the FMA-pipe polynomial.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  // 8 independent chains per thread
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed - 0.01f * i - threadIdx.x * 1e-4f; h[i] = 0xb800b900u + i * 3 + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.5f), "f"(-0.25f)); }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  int iters = 20000;
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps_per_sm * 32>>>(out, 100, -0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148, warps_per_sm * 32>>>(out, iters, -0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * warps_per_sm * 32 * iters * 8;   // thread-instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-10s warps/SM %2d  %.3f ms  %.2f Gthread-instr/s  => %.2f lanes/clk/SM at %.0f MHz nominal (err %s)\n", name, warps_per_sm, ms,
         ops / ms / 1e6, ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("ex2.f32", w);
    run<1>("ex2.f16x2", w);
    run<2>("ex2.bf16x2", w);
    run<3>("ffma", w);
  }
  return 0;
}
