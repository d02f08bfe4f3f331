// Microbenchmark: issue cost (cycles per warp instruction per scheduler) of the instructions the softmax loop is made of,
// on sm_100a: packed f32x2 add / fma, scalar add / fma, 3-input max, bf16x2 pack, integer shift-add, MUFU ex2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
  uint64_t p[8];
  uint32_t u[16];
  unsigned short h[16];
  for (int i = 0; i < 16; ++i) { a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; u[i] = threadIdx.x * 3 + i; h[i] = (unsigned short)(threadIdx.x + i); }
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  const float c = seed * 0.5f;
  uint64_t c2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i & 7]) : "l"(c2));
        if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i & 7]) : "l"(c2));
        if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(c));
        if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == 6) {  // result fed back as the next input so that ptxas cannot hoist it out of the loop
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
          a[i] = __uint_as_float(u[i]);
        }
        if (OP == 7) asm volatile("mad.lo.u32 %0, %1, 8388608, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
        if (OP == 8) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (OP == 9) asm volatile("shf.l.wrap.b32 %0, %0, %1, 23;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
        if (OP == 10) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
        if (OP == 11) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
        if (OP == 12) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
        if (OP == 13) {
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
          a[i] = __uint_as_float(u[i]);
        }
        if (OP == 14) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
        if (OP == 15) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
        if (OP == 16) asm volatile("ex2.approx.ftz.bf16 %0, %0;" : "+h"(h[i]));
        if (OP == 17) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
        if (OP == 18) asm volatile("sub.rn.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i] + (float)u[i] + (float)h[i];
  for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name) {
  const int iters = 20000;
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  for (int w : {4, 8, 16}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, w * 32>>>(out, 100, 0.25f);
    cudaEventRecord(e0);
    k<OP><<<148, w * 32>>>(out, iters, 0.25f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3 / iters / 32.0 / (w / 4);  // cycles per warp instruction per scheduler
    printf("%-22s %d warp(s)/scheduler: %.2f cycles per warp instruction (nominal clock; %s)\n", name, w / 4, cyc, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
}

int main() {
  run<0>("add.f32");
  run<1>("fma.f32");
  run<2>("add.f32x2");
  run<3>("fma.f32x2");
  run<4>("max.f32 (3-input)");
  run<5>("max.f32");
  run<6>("cvt.rn.bf16x2.f32");
  run<7>("mad.lo.u32 (<<23 add)");
  run<8>("ex2.approx.f32");
  run<9>("shf.l.wrap.b32");
  run<10>("max.bf16x2");
  run<11>("ex2.approx.ftz.bf16x2");
  run<12>("ex2.approx.f16x2");
  run<13>("cvt.rn.f16x2.f32");
  run<14>("add.rn.f16x2");
  run<15>("fma.rn.bf16x2");
  run<16>("ex2.approx.ftz.bf16");
  run<17>("tanh.approx.f32");
  run<18>("sub.rn.bf16x2");
  return 0;
}
