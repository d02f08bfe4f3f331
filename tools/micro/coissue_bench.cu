// Microbenchmark: does the MUFU (XU) pipe overlap with FMA / ALU issue on sm_100a? Each iteration issues NM independent
// ex2 and NF independent FFMA (or NA integer adds) per thread; compare the mixed loop against the two pure loops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NM, int NF, int NA>
__global__ void k(float* out, int iters, float seed) {
  float a[8], f[16];
  uint32_t u[16];
  for (int i = 0; i < 8; ++i) a[i] = seed - 0.01f * i - threadIdx.x * 1e-4f;
  for (int i = 0; i < 16; ++i) { f[i] = seed + i; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[r]));
#pragma unroll
      for (int j = 0; j < NF / 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(r * (NF / 8) + j) & 15]) : "f"(0.5f), "f"(-0.25f));
#pragma unroll
      for (int j = 0; j < NA / 8; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[(r * (NA / 8) + j) & 15]) : "r"(it));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 16; ++i) s += f[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF, int NA>
void run(int warps_per_sm) {
  const int iters = 20000;
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NM, NF, NA><<<148, warps_per_sm * 32>>>(out, 100, -0.5f);
  cudaEventRecord(e0);
  k<NM, NF, NA><<<148, warps_per_sm * 32>>>(out, iters, -0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3 / iters;  // cycles per iteration (all warps of an SM in parallel), nominal clock
  printf("ex2 %d ffma %2d iadd %2d  warps/SM %2d (%d per scheduler): %.1f cycles per iteration at %.0f MHz nominal (%s)\n", NM, NF, NA,
         warps_per_sm, warps_per_sm / 4, cyc, clk / 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<8, 0, 0>(w);
    run<0, 32, 0>(w);
    run<8, 32, 0>(w);
    run<0, 64, 0>(w);
    run<8, 64, 0>(w);
    run<0, 0, 32>(w);
    run<8, 0, 32>(w);
    run<8, 32, 32>(w);
    run<0, 32, 32>(w);
  }
  return 0;
}
