// Persistent warp-specialised bf16 GEMM on tcgen05 / TMEM / TMA with fused epilogues (sm_100a).
//
//   C[M,N] = epilogue(A[M,K] . B[N,K]^T)      A, B bf16 K-major (nn.Linear layout), f32 accumulate in TMEM.
//
// Stands in for every nn.Linear / Conv3d the tower runs (HF models/qwen2_vl/modeling_qwen2_vl.py:304-310 PatchEmbed,
// :401-403 qkv, :457 proj, :336-337 fc1/fc2, :324-326 merger) plus the elementwise ops HF runs after them
// (bias, QuickGELU, GELU, residual add, RoPE :257-268), which are fused into the TMEM->register epilogue.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM allocator), warps 2-9 epilogue.
// Tile 128 x BN x 64; UMMA 128xBNx16, cta_group::1; kStages-deep smem ring (SWIZZLE_128B); two TMEM accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1. Grid = #SMs, static round-robin tile order with
// N fastest so the CTAs running together share A row-blocks through L2.
#include <algorithm>

#include "kocr_common.cuh"
#include "kocr_kernels.h"

namespace kocr {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kEpiWarps = 8;
static constexpr int kGemmThreads = 64 + kEpiWarps * 32;

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN > 128) ? 4 : 6;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual 1 KB alignment
  static constexpr int kAccStride = (BN > 128) ? 256 : 128;               // TMEM columns between accumulator stages
  static constexpr int kTmemCols = 2 * kAccStride;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(1.702 x) and x * sigmoid(x) with MUFU ex2 + rcp (no IEEE-division slow path: the epilogue must stay
// branch-free so the 32 elements of a chunk overlap their MUFU latencies). Error ~1e-6 relative, far below bf16.
__device__ __forceinline__ float quick_gelu(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * x)); }
__device__ __forceinline__ float silu(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread's 64-byte share of a row is two full 32-byte
// sectors, so each warp-level access touches 32 whole sectors instead of 32 half sectors (half the LSU transactions).
__device__ __forceinline__ void stg_256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ldg_256(const void* p, uint32_t* w) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float* v) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  stg_256(dst, w);
  stg_256(dst + 16, w + 8);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmEpilogue ep,
            int M, int N, int K) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::kStages * S::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full = empty_bar + S::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = (int)uniform_u32(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<S::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
      for (int kb = 0; kb < num_k; ++kb, ++it) {
        const int s = it % S::kStages;
        const uint32_t ph = (it / S::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], S::kStageBytes);
          tma_load_2d(smem_a + s * S::kABytes, &tm_a, &full_bar[s], kb * BK, m0);
          tma_load_2d(smem_b + s * S::kBBytes, &tm_b, &full_bar[s], kb * BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t desc_hi = smem_desc_hi(1024, 2);  // SWIZZLE_128B, 8-row groups 1024 B apart
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a), 16);
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b), 16);
    const uint32_t tmem_u = uniform_u32(tmem_base);
    uint32_t it = 0, tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int acc = tl & 1;
      const uint32_t acc_ph = (tl >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + acc * S::kAccStride;
      for (int kb = 0; kb < num_k; ++kb, ++it) {
        const int s = it % S::kStages;
        const uint32_t ph = (it / S::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (S::kABytes >> 4);
        const uint32_t b_lo = b_lo0 + s * (S::kBBytes >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_ss_lo(d_tmem, a_lo + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
        tc_commit_elect(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      tc_commit_elect(&tmem_full[acc]);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps, TMEM -> regs -> global.
    // Warp w may touch TMEM lanes 32*(w%4)..+31; the two warps that share a lane quarter split the tile's columns.
    // Everything that does not depend on the accumulator (residual, RoPE angles) is fetched BEFORE the tmem_full wait,
    // and the TMEM load of chunk c+1 is in flight while chunk c is processed, so the 1-2 warps per scheduler are not
    // serialised on load latencies.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int acc = tl & 1;
      const uint32_t acc_ph = (tl >> 1) & 1;
      const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_row = tmem_base + acc * S::kAccStride + ((uint32_t)(q * 32) << 16);
      // folded LayerNorm / RMSNorm: mean and 1/std of this thread's row of A from the producer's partial sums
      float ln_mean = 0.f, ln_rstd = 1.f;
      const bool ln_on = ep.ln_part != nullptr;
      if (ln_on && row_ok) {
        float s1 = 0.f, s2 = 0.f;
        const float2* pp = ep.ln_part + (size_t)row * ep.ln_slots;
        for (int i = 0; i < ep.ln_slots; ++i) {
          const float2 v = pp[i];
          s1 += v.x;
          s2 += v.y;
        }
        if (ep.ln_rms) {
          ln_rstd = rsqrtf(s2 * ep.ln_inv_dim + ep.ln_eps);
        } else {
          ln_mean = s1 * ep.ln_inv_dim;
          ln_rstd = rsqrtf(fmaxf(s2 * ep.ln_inv_dim - ln_mean * ln_mean, 0.f) + ep.ln_eps);
        }
      }

      if constexpr (EPI == kEpiQkvRope) {
        // BN == 240 = [q_h | k_h | v_h] of one head (weights prepacked in that order).
        // half 0: q (RoPE, pre-scale) + v[0:40);  half 1: k (RoPE) + v[40:80).
        const int2 pos = row_ok ? ep.pos_hw[row] : make_int2(0, 0);
        float2 cs[40];  // (cos, sin) of the 40 distinct angles: j < 20 from the row position, else the column position
        {
          const float4* pr = reinterpret_cast<const float4*>(ep.rope_cs + (size_t)pos.x * 20);
          const float4* pc = reinterpret_cast<const float4*>(ep.rope_cs + (size_t)pos.y * 20);
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const float4 a = __ldg(pr + i), b = __ldg(pc + i);
            cs[2 * i] = make_float2(a.x, a.y);
            cs[2 * i + 1] = make_float2(a.z, a.w);
            cs[20 + 2 * i] = make_float2(b.x, b.y);
            cs[20 + 2 * i + 1] = make_float2(b.z, b.w);
          }
        }
        const float mul = (half == 0) ? ep.q_scale : 1.0f;
        const int c_qk = half * 80;         // accumulator column of this half's rotated slot
        const int c_v = 160 + half * 40;    // and of its share of v
        __nv_bfloat16* orow = ep.out + (size_t)row * ep.ldc + n0;
        const float* brow = ep.bias + n0;
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
        uint32_t a[2][8], b[2][8];
        tmem_ld_x8(t_row + c_qk, a[0]);
        tmem_ld_x8(t_row + c_qk + 40, b[0]);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          tc_wait_ld();
          if (c + 1 < 5) {
            tmem_ld_x8(t_row + c_qk + (c + 1) * 8, a[(c + 1) & 1]);
            tmem_ld_x8(t_row + c_qk + 40 + (c + 1) * 8, b[(c + 1) & 1]);
          } else {
            tmem_ld_x8(t_row + c_v, a[(c + 1) & 1]);
            tmem_ld_x8(t_row + c_v + 8, b[(c + 1) & 1]);
          }
          const float4 b1a = __ldg(reinterpret_cast<const float4*>(brow + c_qk + c * 8));
          const float4 b1b = __ldg(reinterpret_cast<const float4*>(brow + c_qk + c * 8 + 4));
          const float4 b2a = __ldg(reinterpret_cast<const float4*>(brow + c_qk + 40 + c * 8));
          const float4 b2b = __ldg(reinterpret_cast<const float4*>(brow + c_qk + 40 + c * 8 + 4));
          const float bb1[8] = {b1a.x, b1a.y, b1a.z, b1a.w, b1b.x, b1b.y, b1b.z, b1b.w};
          const float bb2[8] = {b2a.x, b2a.y, b2a.z, b2a.w, b2b.x, b2b.y, b2b.z, b2b.w};
          float cc1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cc2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          if (ln_on && !ep.ln_rms) {
            const float* crow = ep.ln_c1 + n0 + c_qk + c * 8;
            const float4 c1a = __ldg(reinterpret_cast<const float4*>(crow)), c1b = __ldg(reinterpret_cast<const float4*>(crow + 4));
            const float4 c2a = __ldg(reinterpret_cast<const float4*>(crow + 40)), c2b = __ldg(reinterpret_cast<const float4*>(crow + 44));
            cc1[0] = c1a.x; cc1[1] = c1a.y; cc1[2] = c1a.z; cc1[3] = c1a.w; cc1[4] = c1b.x; cc1[5] = c1b.y; cc1[6] = c1b.z; cc1[7] = c1b.w;
            cc2[0] = c2a.x; cc2[1] = c2a.y; cc2[2] = c2a.z; cc2[3] = c2a.w; cc2[4] = c2b.x; cc2[5] = c2b.y; cc2[6] = c2b.z; cc2[7] = c2b.w;
          }
          float lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x1 = ln_rstd * (__uint_as_float(a[c & 1][i]) - ln_mean * cc1[i]) + bb1[i];
            const float x2 = ln_rstd * (__uint_as_float(b[c & 1][i]) - ln_mean * cc2[i]) + bb2[i];
            const float2 t = cs[c * 8 + i];
            lo[i] = (x1 * t.x - x2 * t.y) * mul;
            hi[i] = (x2 * t.x + x1 * t.y) * mul;
          }
          if (row_ok) {
            // q and k are stored with their head dims permuted to [j0..7, j40..47, j8..15, j48..55, ...]: q.k is invariant
            // under a permutation applied to both, and each iteration becomes one aligned 256-bit store
            const uint32_t w8[8] = {pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]),
                                    pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7])};
            stg_256(orow + c_qk + c * 16, w8);
          }
        }
        // v share: 40 columns = 5 x8 loads; the first two are already in flight in a[1], b[1] (5 & 1 == 1)
        uint32_t v2[3][8];
        tc_wait_ld();
        tmem_ld_x8(t_row + c_v + 16, v2[0]);
        tmem_ld_x8(t_row + c_v + 24, v2[1]);
        tmem_ld_x8(t_row + c_v + 32, v2[2]);
        float vb[40], vc[40];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(brow + c_v) + i);
          vb[4 * i] = t.x; vb[4 * i + 1] = t.y; vb[4 * i + 2] = t.z; vb[4 * i + 3] = t.w;
          float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ln_on && !ep.ln_rms) u = __ldg(reinterpret_cast<const float4*>(ep.ln_c1 + n0 + c_v) + i);
          vc[4 * i] = u.x; vc[4 * i + 1] = u.y; vc[4 * i + 2] = u.z; vc[4 * i + 3] = u.w;
        }
        tc_wait_ld();
        if (row_ok) {
          const uint32_t* src[5] = {a[1], b[1], v2[0], v2[1], v2[2]};
          uint32_t w[20];
#pragma unroll
          for (int g = 0; g < 5; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              w[g * 4 + i] = pack_bf16(ln_rstd * (__uint_as_float(src[g][2 * i]) - ln_mean * vc[g * 8 + 2 * i]) + vb[g * 8 + 2 * i],
                                       ln_rstd * (__uint_as_float(src[g][2 * i + 1]) - ln_mean * vc[g * 8 + 2 * i + 1]) + vb[g * 8 + 2 * i + 1]);
          // v keeps its natural order. Its 80-byte share starts 32-byte aligned for half 0 and 16 bytes off for half 1.
          __nv_bfloat16* dst = orow + c_v;
          if (half == 0) {
            stg_256(dst, w);
            stg_256(dst + 16, w + 8);
            *reinterpret_cast<uint4*>(dst + 32) = make_uint4(w[16], w[17], w[18], w[19]);
          } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            stg_256(dst + 8, w + 4);
            stg_256(dst + 24, w + 12);
          }
        }
      } else {
        constexpr bool kSwiglu = EPI == KOCR_EPI_BIAS_SWIGLU;
        constexpr int kChunks = BN / 32 / 2;  // chunks of 32 accumulator columns per half
        const int col0 = n0 + half * (BN / 2);
        int nch = (N - col0 + 31) / 32;
        nch = nch < 0 ? 0 : (nch > kChunks ? kChunks : nch);
        __nv_bfloat16* orow = ep.out + (size_t)row * ep.ldc + (kSwiglu ? col0 / 2 : col0);
        const __nv_bfloat16* rrow = ep.residual + (size_t)row * ep.ld_res + col0;
        uint32_t res[2][16];
        if constexpr (EPI == KOCR_EPI_BIAS_RESIDUAL) {
          if (row_ok && nch > 0) {
            ldg_256(rrow, res[0]);
            ldg_256(rrow + 16, res[0] + 8);
          }
        }
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
        uint32_t r[2][32];
        float st_sum = 0.f, st_sq = 0.f;
        tmem_ld_x32(t_row + half * (BN / 2), r[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          // bias of this chunk is fetched before the TMEM wait so both latencies overlap
          float4 bv[8], cv[8];
          if constexpr (EPI != KOCR_EPI_NONE) {
            if (c < nch) {
#pragma unroll
              for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + c * 32) + i);
              if (ln_on && !ep.ln_rms) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cv[i] = __ldg(reinterpret_cast<const float4*>(ep.ln_c1 + col0 + c * 32) + i);
              }
            }
          }
          tc_wait_ld();
          if (c + 1 < kChunks) tmem_ld_x32(t_row + half * (BN / 2) + (c + 1) * 32, r[(c + 1) & 1]);
          if (c < nch) {
            if constexpr (EPI == KOCR_EPI_BIAS_RESIDUAL) {
              if (row_ok && c + 1 < nch) {
                ldg_256(rrow + (c + 1) * 32, res[(c + 1) & 1]);
                ldg_256(rrow + (c + 1) * 32 + 16, res[(c + 1) & 1] + 8);
              }
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if constexpr (EPI == KOCR_EPI_NONE) bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (!(ln_on && !ep.ln_rms)) cv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * i] = ln_rstd * (__uint_as_float(r[c & 1][4 * i]) - ln_mean * cv[i].x) + bv[i].x;
              v[4 * i + 1] = ln_rstd * (__uint_as_float(r[c & 1][4 * i + 1]) - ln_mean * cv[i].y) + bv[i].y;
              v[4 * i + 2] = ln_rstd * (__uint_as_float(r[c & 1][4 * i + 2]) - ln_mean * cv[i].z) + bv[i].z;
              v[4 * i + 3] = ln_rstd * (__uint_as_float(r[c & 1][4 * i + 3]) - ln_mean * cv[i].w) + bv[i].w;
            }
            if constexpr (EPI == KOCR_EPI_BIAS_QUICKGELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
            }
            if constexpr (EPI == KOCR_EPI_BIAS_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            }
            if constexpr (EPI == KOCR_EPI_BIAS_RESIDUAL) {
              const uint32_t* rw = res[c & 1];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[2 * i] += bf16_lo(rw[i]);
                v[2 * i + 1] += bf16_hi(rw[i]);
              }
            }
            if constexpr (kSwiglu) {
              // accumulator columns alternate gate_j, up_j; output column j = silu(gate) * up (output width N/2)
              float o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = silu(v[2 * i]) * v[2 * i + 1];
              if (row_ok) {
                uint32_t w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = pack_bf16(o[2 * i], o[2 * i + 1]);
                stg_256(orow + c * 16, w8);
              }
            } else {
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) w[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
              if (ep.stat_part) {  // statistics of the values as the next norm will read them (bf16-rounded)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float x0 = bf16_lo(w[i]), x1 = bf16_hi(w[i]);
                  st_sum += x0 + x1;
                  st_sq += x0 * x0 + x1 * x1;
                }
              }
              if (row_ok) {
                stg_256(orow + c * 32, w);
                stg_256(orow + c * 32 + 16, w + 8);
              }
            }
          }
        }
        if constexpr (EPI == KOCR_EPI_BIAS_RESIDUAL || EPI == KOCR_EPI_NONE) {
          if (ep.stat_part && row_ok) ep.stat_part[(size_t)row * ep.stat_slots + (tile % num_n) * 2 + half] = make_float2(st_sum, st_sq);
        }
      }
      // all of this warp's TMEM reads are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI>
static int launch_one(Ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N,
                      int K, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  auto kern = gemm_kernel<BN, EPI>;
  if (int rc = ctx->opt_in_smem(reinterpret_cast<const void*>(kern), S::kTotal)) return rc;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = std::min(tiles, ctx->num_sms - ctx->reserved_sms);  // one persistent CTA per SM that is ours
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(ta, tb, ep, M, N, K);
  KOCR_LAUNCH_CHECK("gemm_kernel");
  return KOCR_OK;
}

int launch_gemm(Ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                int epi, const GemmEpilogue& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return fail(KOCR_ERR_INVALID, "gemm: non-positive dimension");
  if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail(KOCR_ERR_UNSUPPORTED, "gemm: dimension over 2^31");
  if (lda % 8 || ldb % 8 || ep.ldc % 16 || ep.ld_res % 16)
    return fail(KOCR_ERR_UNSUPPORTED, "gemm: A/B row pitches must be multiples of 8 elements, C/residual pitches of 16");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15)
    return fail(KOCR_ERR_UNSUPPORTED, "gemm: A and B must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(ep.out) | reinterpret_cast<uintptr_t>(ep.residual)) & 31)
    return fail(KOCR_ERR_UNSUPPORTED, "gemm: C and residual must be 32-byte aligned (256-bit epilogue accesses)");
  const int bn = (epi == kEpiQkvRope) ? 240 : 256;
  if (epi == kEpiQkvRope) {
    if (N % 240 || ep.ldc % 16) return fail(KOCR_ERR_UNSUPPORTED, "gemm(qkv_rope): N must be a multiple of 240, ldc of 16");
  } else if (epi == KOCR_EPI_BIAS_SWIGLU) {
    if (N % 64) return fail(KOCR_ERR_UNSUPPORTED, "gemm(swiglu): N must be a multiple of 64");
  } else if (N % 32) {
    return fail(KOCR_ERR_UNSUPPORTED, "gemm: N must be a multiple of 32");
  }
  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tensor_map(&ta, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_tensor_map(&tb, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  const int m = (int)M, n = (int)N, k = (int)K;
  switch (epi) {
    case KOCR_EPI_NONE: return launch_one<256, KOCR_EPI_NONE>(ctx, ta, tb, ep, m, n, k, stream);
    case KOCR_EPI_BIAS: return launch_one<256, KOCR_EPI_BIAS>(ctx, ta, tb, ep, m, n, k, stream);
    case KOCR_EPI_BIAS_QUICKGELU: return launch_one<256, KOCR_EPI_BIAS_QUICKGELU>(ctx, ta, tb, ep, m, n, k, stream);
    case KOCR_EPI_BIAS_GELU: return launch_one<256, KOCR_EPI_BIAS_GELU>(ctx, ta, tb, ep, m, n, k, stream);
    case KOCR_EPI_BIAS_RESIDUAL: return launch_one<256, KOCR_EPI_BIAS_RESIDUAL>(ctx, ta, tb, ep, m, n, k, stream);
    case KOCR_EPI_BIAS_SWIGLU: return launch_one<256, KOCR_EPI_BIAS_SWIGLU>(ctx, ta, tb, ep, m, n, k, stream);
    case kEpiQkvRope: return launch_one<240, kEpiQkvRope>(ctx, ta, tb, ep, m, n, k, stream);
    default: return fail(KOCR_ERR_INVALID, "gemm: unknown epilogue");
  }
}

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_op_gemm(KocrCtx* ctx_, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                            const void* residual, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue,
                            void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !A || !B || !C) return fail(KOCR_ERR_INVALID, "kocr_op_gemm: null argument");
  if (epilogue < KOCR_EPI_NONE || epilogue > KOCR_EPI_BIAS_SWIGLU) return fail(KOCR_ERR_INVALID, "kocr_op_gemm: bad epilogue");
  if (epilogue != KOCR_EPI_NONE && !bias) return fail(KOCR_ERR_INVALID, "kocr_op_gemm: epilogue needs a bias");
  if (epilogue == KOCR_EPI_BIAS_RESIDUAL && !residual) return fail(KOCR_ERR_INVALID, "kocr_op_gemm: residual is null");
  GemmEpilogue ep{};
  ep.bias = bias;
  ep.residual = static_cast<const __nv_bfloat16*>(residual);
  ep.ld_res = ldc;
  ep.out = static_cast<__nv_bfloat16*>(C);
  ep.ldc = ldc;
  reset_launch_count();
  return launch_gemm(ctx, A, lda, B, ldb, M, N, K, epilogue, ep, reinterpret_cast<cudaStream_t>(stream));
}
