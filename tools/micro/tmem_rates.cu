// Microbenchmark: tcgen05.ld / tcgen05.st rates on sm_100a. How many cycles does a warp need to pull its 32 lanes x N
// columns out of TMEM, alone, with one warp per lane quarter, and with two warps per lane quarter (the attention kernel's
// softmax arrangement)? Answers whether the softmax chain is bounded by TMEM read bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rates tmem_rates.cu && ./tmem_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: ld x32 + wait each; 1: three loads (x32, x32, x16 = 80 columns) then one wait; 2: st x16 x2 + x8 (40 columns) + wait::st
__global__ void k(long long* cyc_out, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 160;
  uint32_t r[80];
#pragma unroll
  for (int i = 0; i < 80; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(base + (it & 1) * 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[17] ^ r[31];
    } else if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[c * 32 + 0]), "=r"(r[c * 32 + 1]), "=r"(r[c * 32 + 2]), "=r"(r[c * 32 + 3]), "=r"(r[c * 32 + 4]), "=r"(r[c * 32 + 5]),
              "=r"(r[c * 32 + 6]), "=r"(r[c * 32 + 7]), "=r"(r[c * 32 + 8]), "=r"(r[c * 32 + 9]), "=r"(r[c * 32 + 10]), "=r"(r[c * 32 + 11]),
              "=r"(r[c * 32 + 12]), "=r"(r[c * 32 + 13]), "=r"(r[c * 32 + 14]), "=r"(r[c * 32 + 15]), "=r"(r[c * 32 + 16]), "=r"(r[c * 32 + 17]),
              "=r"(r[c * 32 + 18]), "=r"(r[c * 32 + 19]), "=r"(r[c * 32 + 20]), "=r"(r[c * 32 + 21]), "=r"(r[c * 32 + 22]), "=r"(r[c * 32 + 23]),
              "=r"(r[c * 32 + 24]), "=r"(r[c * 32 + 25]), "=r"(r[c * 32 + 26]), "=r"(r[c * 32 + 27]), "=r"(r[c * 32 + 28]), "=r"(r[c * 32 + 29]),
              "=r"(r[c * 32 + 30]), "=r"(r[c * 32 + 31])
            : "r"(base + c * 32));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71]), "=r"(r[72]), "=r"(r[73]),
            "=r"(r[74]), "=r"(r[75]), "=r"(r[76]), "=r"(r[77]), "=r"(r[78]), "=r"(r[79])
          : "r"(base + 64));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[40] ^ r[79];
    } else {
#pragma unroll
      for (int c = 0; c < 2; ++c)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + c * 16),
            "r"(r[c * 16 + 0]), "r"(r[c * 16 + 1]), "r"(r[c * 16 + 2]), "r"(r[c * 16 + 3]), "r"(r[c * 16 + 4]), "r"(r[c * 16 + 5]), "r"(r[c * 16 + 6]),
            "r"(r[c * 16 + 7]), "r"(r[c * 16 + 8]), "r"(r[c * 16 + 9]), "r"(r[c * 16 + 10]), "r"(r[c * 16 + 11]), "r"(r[c * 16 + 12]),
            "r"(r[c * 16 + 13]), "r"(r[c * 16 + 14]), "r"(r[c * 16 + 15])
            : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base + 32), "r"(r[32]), "r"(r[33]),
                   "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39])
                   : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      r[0] += 1; r[20] ^= r[0]; r[39] += r[20];
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc_out[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[5];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int cols) {
  const int iters = 20000;
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 32 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  for (int w : {1, 4, 8, 12}) {  // 12 warps: three per lane quarter
    k<MODE><<<148, w * 32>>>(cyc, sink, 100);
    k<MODE><<<148, w * 32>>>(cyc, sink, iters);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / iters;
    printf("%-44s %2d warp(s)/CTA: %7.1f cycles per iteration per warp, %6.1f B/clk per warp, %7.1f B/clk per SM (%s)\n", name, w, per,
           cols * 128.0 / per, cols * 128.0 * w / per, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  run<0>("tcgen05.ld x32 + wait::ld", 32);
  run<1>("tcgen05.ld x32,x32,x16 (80 cols) + wait::ld", 80);
  run<2>("tcgen05.st x16,x16,x8 (40 cols) + wait::st", 40);
  return 0;
}
