// Varlen non-causal flash attention for head_dim 80 on tcgen05 / TMEM / TMA (sm_100a).
//
// Stands in for VisionAttention's attention call (HF models/qwen2_vl/modeling_qwen2_vl.py:411-454: one
// softmax(q k^T / sqrt(80)) v per cu_seqlens segment, non-causal), which HF runs as a Python loop of SDPA calls or as
// flash_attn_varlen_func.
//
// Input is the tower's packed QKV buffer [S, heads*240] with, per head, 80 q | 80 k | 80 v columns; q is already
// rotated and multiplied by head_dim^-0.5 * log2(e) (fused QKV GEMM epilogue), so scores are in the log2 domain.
//
// Two kernels share the operand layouts, the online softmax and the MMA forms below.
//
// attention3_kernel (full attention, the one the tower launches): CTA = one 384-row query block (three 128-row tiles) of one
// sequence and one head, 512 threads: warp 0 TMA, warps 1-3 one MMA issuer per tile, warps 4-15 softmax (three per TMEM lane
// quarter = three per scheduler). Scores are SINGLE-buffered per tile (TMEM: S_t 80 columns with P written over the first 40,
// O_t 80 columns, 3 x 160 = 480): S_t(i+1) is issued right behind P_t(i).V. Tiles without a valid row (a sequence's last
// block) are skipped. See the comment at the kernel.
//
// attention_kernel<kWin> (windowed layers of Qwen2.5-VL; <false> is the two-tile full-attention kernel that attention3 replaced,
// kept for A/B runs): CTA = one 256-row query block of one sequence and one head; 384 threads:
//   warp 0      TMA producer: Q once, then a ring of K tiles and a ring of V tiles (80 keys each = one score sub-step)
//   warps 1,2   MMA issuers (one per query tile): S_t = Q_t K^T (SS) and O_t += P_t V (P from TMEM, V MN-major)
//   warps 4-7   softmax of query tile 0 (rows 0..127), one row per thread
//   warps 8-11  softmax of query tile 1 (rows 128..255)
// Operand tiles are stored as five [rows x 16 columns] SWIZZLE_32B chunks, which is a canonical K-major layout for
// Q/K (K = head_dim) and, unchanged, a canonical MN-major layout for V (N = head_dim): no transpose, no padding of 80.
// Scores are produced in 80-key sub-steps and double-buffered in TMEM per query tile (S[t][b], b = sub-step parity), so
// S_t(i+2) is computed while the softmax warps still work on sub-step i: they never wait for the tensor pipe. 80 is
// the largest sub-step for which two tiles x two score buffers and the two O accumulators fit the 512 TMEM columns
// (4*80 + 2*80 = 480). What bounds the kernel is the softmax warps' own serial chain per sub-step (TMEM load, row max,
// 80 exponentials, pack, TMEM store, barrier round trips); the two tiles' warps take turns on the MUFU unit (kPingPong).
// TMEM: S[t][b] at t*160 + b*80 (80 columns); P[t][b] (bf16) overwrites the first 40 columns of S[t][b]; O_t at 320 + t*80
// (windowed shape: 64-key sub-steps, S[b] at b*64, O at 128).
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "kocr_common.cuh"
#include "kocr_kernels.h"

namespace kocr {

static constexpr int kHd = 80;
static constexpr int kChunks = kHd / 16;           // 5 chunks of 16 columns
static constexpr int kTileRows = 128;
static constexpr int kChunkBytes = kTileRows * 32;  // 4096 (Q)
static constexpr int kTileBytes = kChunks * kChunkBytes;  // 20480 (Q)
// Two kernel shapes. Full attention (long key ranges): two query tiles per CTA, 384 threads, all 512 TMEM columns, six
// K/V stages, one CTA per SM. Windowed layers (a 128-row block needs at most ~190 keys, three sub-steps): one query tile
// per CTA, 256 threads, 256 TMEM columns, three K/V stages (97 KB), so that two CTAs share an SM and the prologue /
// epilogue of one overlaps the few sub-steps of the other; each tile also gets its own, tight key range.
template <bool kWin> struct AttnShape {
  // keys per score sub-step = keys per K/V tile. Full attention: 80, the largest for which two tiles x two score buffers
  // and two O accumulators fit the 512 TMEM columns. Windowed: 64 (a block's ~190 keys are three sub-steps either way).
  static constexpr int kSub = kWin ? 64 : 80;
  static constexpr int kKvChunkBytes = kSub * 32;               // 2560 / 2048
  static constexpr int kKvTileBytes = kChunks * kKvChunkBytes;  // 12800 / 10240
  static constexpr int kTiles = kWin ? 1 : 2;
  static constexpr int kStages = kWin ? 3 : 6;
  static constexpr int kThreads = 128 + 128 * kTiles;
  static constexpr int kTmemCols = kWin ? 256 : 512;
  static constexpr int kSmem = kTiles * kTileBytes + 2 * kStages * kKvTileBytes + 512 + 1024;
  static constexpr int kProdRegs = kWin ? 56 : 96;  // TMA / MMA warps; softmax warps take 200: 128*56 + 128*200 = 256*128, 128*96 + 256*200 <= 384*168
  __host__ __device__ static constexpr int s_col(int t, int b) { return t * 2 * kSub + b * kSub; }     // S[t][b] (P[t][b] in its first half)
  __host__ __device__ static constexpr int o_col(int t) { return (kWin ? 2 : 4) * kSub + t * kHd; }   // O_t behind the score buffers
};
#ifndef KOCR_PINGPONG
#define KOCR_PINGPONG 1
#endif
static constexpr bool kPingPong = KOCR_PINGPONG;
#ifndef KOCR_PP_AT
#define KOCR_PP_AT 28
#endif
static constexpr int kPpAt = KOCR_PP_AT;  // the exponent phase is handed over after this pair of columns (of 40; must be a MUFU pair)
#ifndef KOCR_POLY_EVERY
#define KOCR_POLY_EVERY 4
#endif
static constexpr int kPolyEvery = KOCR_POLY_EVERY;  // every 4th pair of exponentials is evaluated on the FMA pipe instead of MUFU (0 = never)
#ifndef KOCR_PROBE
#define KOCR_PROBE 0   // timing probes (wrong results): bit 0 no exponentials, bit 1 no row max, bit 2 no TMEM score load, bit 3 no P store, bit 4 no row sum, bit 5 no subtraction, bit 6 no bf16 conversion
#endif
static constexpr int kProbe = KOCR_PROBE;
static constexpr float kRescaleThreshold = 8.0f;

// Timeline trace of one CTA (variant builds with -DKOCR_TRACE only): lane 0 of every warp stamps clock64 at fixed points of each
// sub-step; tools/attn_trace.py reads the stamps back and prints where each warp's time goes.
#ifdef KOCR_TRACE
static constexpr int kTraceSubs = 96, kTracePts = 12, kTraceWarps = 16;
__device__ long long g_attn_trace[kTraceWarps][kTraceSubs][kTracePts];
#define KOCR_STAMP(i, pt)                                                                                     \
  do {                                                                                                        \
    if (trace_on && lane == 0 && (i) < kTraceSubs) g_attn_trace[warp][(i)][(pt)] = clock64();                \
  } while (0)
#else
#define KOCR_STAMP(i, pt) do {} while (0)
#endif  // log2 units: rescale O only when the row max grows by more than 2^8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Maximum of N values held in registers as a 3-ary tree of FMNMX3 (depth 4 for 80 values): every index is a compile-time constant
template <int N, typename T>
__device__ __forceinline__ float max_tree(const T* v) {
  auto f = [](T x) -> float {
    if constexpr (sizeof(T) == 4 && !std::is_same<T, float>::value) return __uint_as_float(x);
    else return x;
  };
  if constexpr (N == 1) {
    return f(v[0]);
  } else if constexpr (N == 2) {
    return fmaxf(f(v[0]), f(v[1]));
  } else if constexpr (N == 3) {
    return max3(f(v[0]), f(v[1]), f(v[2]));
  } else {
    constexpr int M = (N + 2) / 3;
    float u[M];
#pragma unroll
    for (int g = 0; g < N / 3; ++g) u[g] = max3(f(v[3 * g]), f(v[3 * g + 1]), f(v[3 * g + 2]));
    if constexpr (N % 3 == 1) u[M - 1] = f(v[N - 1]);
    if constexpr (N % 3 == 2) u[M - 1] = fmaxf(f(v[N - 2]), f(v[N - 1]));
    return max_tree<M, float>(u);
  }
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
  static_assert(N == 80 || N == 40, "unsupported column count");
  if constexpr (N == 80) {
#pragma unroll
    for (int c = 0; c < 5; ++c) tmem_ld_x16(taddr + c * 16, r + c * 16);
  } else {
    tmem_ld_x16(taddr, r);
    tmem_ld_x16(taddr + 16, r + 16);
    tmem_ld_x8(taddr + 32, r + 32);
  }
}
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* r) {
  if constexpr (N == 80) {
#pragma unroll
    for (int c = 0; c < 5; ++c) tmem_st_x16(taddr + c * 16, r + c * 16);
  } else {
    tmem_st_x16(taddr, r);
    tmem_st_x16(taddr + 16, r + 16);
    tmem_st_x8(taddr + 32, r + 32);
  }
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 2^x for a pair of x <= 0 without the MUFU unit: x = n + r, n = round(x) via the 1.5*2^23 trick, 2^r by a degree-3
// minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), then n is added
// into the exponent field. x is clamped at -126 so the result never leaves the normal range (masked scores are -inf).
__device__ __forceinline__ uint64_t ex2_poly_f32x2(uint64_t x2) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);
  const uint64_t t2 = add_f32x2(xc, magic);                                    // low mantissa bits = round(x)
  const uint64_t n2 = add_f32x2(t2, pack_f32x2(-12582912.0f, -12582912.0f));   // round(x) as a float
  const uint64_t r2 = fma_f32x2(n2, pack_f32x2(-1.0f, -1.0f), xc);             // r in [-0.5, 0.5]
  uint64_t p2 = fma_f32x2(r2, pack_f32x2(0.055171649903059006f, 0.055171649903059006f), pack_f32x2(0.2426111251115799f, 0.2426111251115799f));
  p2 = fma_f32x2(p2, r2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  p2 = fma_f32x2(p2, r2, pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
  float p0, p1, t0, t1;
  unpack_f32x2(p2, p0, p1);
  unpack_f32x2(t2, t0, t1);
  const uint32_t b0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  const uint32_t b1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
  return pack_u32x2(b0, b1);
}

// kWin: block-diagonal attention inside a query block (Qwen2.5-VL windows, HF modeling_qwen2_5_vl.py:498-502): row r may
// attend only the keys [win[r].x, win[r].y) of its own window; a 128-row block packs two or more windows.
template <bool kWin>
__global__ void __launch_bounds__(AttnShape<kWin>::kThreads, kWin ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, __nv_bfloat16* __restrict__ out,
                 const AttnWork* __restrict__ work, int num_heads, const int2* __restrict__ win) {
  using Shape = AttnShape<kWin>;
  constexpr int kTiles = Shape::kTiles;
  constexpr int kKvStages = Shape::kStages;
  constexpr int kSub = Shape::kSub, kKvChunkBytes = Shape::kKvChunkBytes, kKvTileBytes = Shape::kKvTileBytes;
  constexpr bool kPP = kPingPong && kTiles == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // kTiles query tiles
  uint8_t* smem_k = smem_q + kTiles * kTileBytes;           // kKvStages tiles of kSub keys
  uint8_t* smem_v = smem_k + kKvStages * kKvTileBytes;      // kKvStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKvStages * kKvTileBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // kKvStages
  uint64_t* k_empty = k_full + kKvStages;  // kKvStages
  uint64_t* v_full = k_empty + kKvStages;
  uint64_t* v_empty = v_full + kKvStages;
  uint64_t* s_full = v_empty + kKvStages;  // [tile][buffer] = 4
  uint64_t* p_full = s_full + 4;           // [tile][buffer] = 4
  uint64_t* o_done = p_full + 4;           // [tile][sub-step parity] = 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 4);

  const int warp = (int)uniform_u32(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const AttnWork w = work[blockIdx.x];
  const int head = blockIdx.y;
#ifdef KOCR_TRACE
  const bool trace_on = !kWin && blockIdx.x == 13 && blockIdx.y == 5;
#endif
  const int n_sub = (w.kv_len + kSub - 1) / kSub;           // score sub-steps = K/V tiles
  const int col_q = head * 3 * kHd, col_k = col_q + kHd, col_v = col_q + 2 * kHd;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], kTiles);  // one tcgen05.commit from each query tile's MMA warp
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], kTiles);
    }
    for (int t = 0; t < 4; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Shape::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S[t][b] (80 f32 columns, b = sub-step parity) at t*160 + b*80; P[t][b] (bf16 pairs) overwrites the
  // first 40 columns of S[t][b]; O[t] (80 columns) behind the score buffers.

  if (warp < 4) {
    setmaxnreg_dec<Shape::kProdRegs>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (warp-uniform, elected lane issues)
      const int kv_begin = (int)uniform_u32(w.kv_begin), q_begin = (int)uniform_u32(w.q_begin);
      const int n_sub_u = (int)uniform_u32(n_sub);
      if (elect_one()) {
        mbar_expect_tx(q_full, kTiles * kTileBytes);
        for (int t = 0; t < kTiles; ++t)
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_q + t * kTileBytes + c * kChunkBytes, &tm_q, q_full, col_q + c * 16, q_begin + t * kTileRows);
      }
      __syncwarp();
      for (int j = 0; j < n_sub_u; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        const int row = kv_begin + j * kSub;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[s], kKvTileBytes);
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_k + s * kKvTileBytes + c * kKvChunkBytes, &tm_kv, &k_full[s], col_k + c * 16, row);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&v_full[s], kKvTileBytes);
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_v + s * kKvTileBytes + c * kKvChunkBytes, &tm_kv, &v_full[s], col_v + c * 16, row);
        }
        __syncwarp();
      }
    } else if (warp == 1 || (warp == 2 && kTiles == 2)) {
      // ---------------------------------------------------------------- MMA issuers: warp 1 -> query tile 0, warp 2 -> tile 1
      // (warp-uniform loops, elected lane issues). Scores are double-buffered per query tile: S_t(i+2) is issued right
      // after P_t(i).V, two sub-steps ahead of the softmax that will read it, so the softmax warps do not wait for the
      // tensor pipe. One issuing warp per tile: a single warp could not keep up with 20 MMAs per 80-key sub-step.
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kSub, 0, 0);  // Q (K-major) x K (K-major), 80 keys
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd, 0, 1);   // P (TMEM) x V (MN-major)
      constexpr uint32_t hi32 = smem_desc_hi(256, 6);                 // SWIZZLE_32B, 8-row groups 256 B apart
      const int t = warp - 1;
      const uint32_t tmem_u = uniform_u32(tmem_base);
      const int n_sub_u = (int)uniform_u32(n_sub);
      const uint32_t q_lo = smem_desc_lo(smem_u32(smem_q), 16) + t * (kTileBytes >> 4);
      const uint32_t k_lo = smem_desc_lo(smem_u32(smem_k), 16);
      const uint32_t v_lo = smem_desc_lo(smem_u32(smem_v), kKvChunkBytes);  // LBO = distance between 16-column groups
      const uint32_t d_o = tmem_u + Shape::o_col(t);
      auto issue_s = [&](int i) {
        const int s = i % kKvStages;
        mbar_wait(&k_full[s], (i / kKvStages) & 1);
        tc_fence_after();
        const uint32_t ka = k_lo + s * (kKvTileBytes >> 4);
        const uint32_t d = tmem_u + Shape::s_col(t, i & 1);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          umma_ss_lo(d, q_lo + c * (kChunkBytes >> 4), ka + c * (kKvChunkBytes >> 4), hi32, idesc_s, c != 0);
        tc_commit_elect(&s_full[t * 2 + (i & 1)]);
        tc_commit_elect(&k_empty[s]);
      };
      auto issue_pv = [&](int i) {
        const int s = i % kKvStages;
        mbar_wait(&v_full[s], (i / kKvStages) & 1);
        KOCR_STAMP(i, 3);
        mbar_wait(&p_full[t * 2 + (i & 1)], (i >> 1) & 1);
        KOCR_STAMP(i, 4);
        tc_fence_after();
        const uint32_t va = v_lo + s * (kKvTileBytes >> 4);
        const uint32_t pa = tmem_u + Shape::s_col(t, i & 1);
#pragma unroll
        for (int ks = 0; ks < kSub / 16; ++ks)  // 16 keys per step: rows ks*16.. of every chunk, 512 B further
          umma_ts_lo(d_o, pa + ks * 8, va + ks * (512 >> 4), hi32, idesc_o, (i > 0 || ks != 0));
        tc_commit_elect(&o_done[t * 2 + (i & 1)]);
        tc_commit_elect(&v_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n_sub_u > 1) issue_s(1);
      for (int i = 0; i < n_sub_u; ++i) {
        KOCR_STAMP(i, 0);
        issue_pv(i);
        KOCR_STAMP(i, 1);
        if (i + 2 < n_sub_u) issue_s(i + 2);
        KOCR_STAMP(i, 2);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps, one query row per thread
    setmaxnreg_inc<200>();                   // see AttnShape::kProdRegs for the budget
    const int sw = warp - 4;
    const int t = sw >> 2;                   // query tile
    const int qtr = warp & 3;                // TMEM lane quarter
    const int r = qtr * 32 + lane;           // row within the tile
    const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
    const uint32_t t_o = tmem_base + Shape::o_col(t) + lane_off;
    // Exponent phase hand-over (kPingPong): the two tiles' warps of one lane quarter sit on the same scheduler and share
    // its MUFU unit. Left alone they run in lockstep - both in the exponent phase, then both in the TMEM / max / pack
    // phase - and the unit idles half the time. A token passed through two 64-thread named barriers makes them alternate.
    const int pp_mine = 1 + t * 4 + qtr, pp_other = 1 + (1 - t) * 4 + qtr;
    if (kPP && t == 1) named_bar_arrive(pp_other, 64);  // tile 0 goes first
    float m_ref = -INFINITY, l = 0.f;
    int w_lo = 0, w_hi = 0;  // this row's window, as key indices relative to kv_begin
    if (kWin) {
      const int qr = t * kTileRows + r;
      if (qr < w.q_rows) {
        const int2 wb = win[w.q_begin + qr];
        w_lo = wb.x - w.kv_begin;
        w_hi = wb.y - w.kv_begin;
      }
    }
    for (int i = 0; i < n_sub; ++i) {
      const int b = i & 1;                    // S/P buffer and barrier slot of this sub-step
      const uint32_t ph = (i >> 1) & 1;      // phase of s_full / p_full / o_done[b] for this sub-step
      const uint32_t t_s = tmem_base + Shape::s_col(t, b) + lane_off;
      KOCR_STAMP(i, 0);
      mbar_wait(&s_full[t * 2 + b], ph);
      KOCR_STAMP(i, 1);
      tc_fence_after();
      uint32_t sr[kSub];
      if (kProbe & 4) {
#pragma unroll
        for (int c = 0; c < kSub; ++c) sr[c] = __float_as_uint(-0.01f * (float)((c * 7 + i + lane) & 63));
      } else {
        tmem_ld_x32(t_s, sr);
        tmem_ld_x32(t_s + 32, sr + 32);
        if constexpr (kSub == 80) tmem_ld_x16(t_s + 64, sr + 64);
        tc_wait_ld();
      }
      KOCR_STAMP(i, 2);
      const int c0 = i * kSub;  // key index (relative to kv_begin) of the first column
      if (kWin) {
        const int c_lo = w_lo - c0, c_hi = w_hi - c0;  // valid columns: [c_lo, c_hi)
        if (c_lo > 0 || c_hi < kSub) {
#pragma unroll
          for (int c = 0; c < kSub; ++c)
            if (c < c_lo || c >= c_hi) sr[c] = 0xff800000u;  // -inf
        }
      } else {
        const int valid = w.kv_len - c0;
        if (valid < kSub) {
#pragma unroll
          for (int c = 0; c < kSub; ++c)
            if (c >= valid) sr[c] = 0xff800000u;  // -inf
        }
      }
      // row max: independent FMNMX3 chains of 16 columns
      float mxa[kSub / 16];
#pragma unroll
      for (int g = 0; g < kSub / 16; ++g) {
        mxa[g] = max3(__uint_as_float(sr[16 * g]), __uint_as_float(sr[16 * g + 1]), __uint_as_float(sr[16 * g + 2]));
#pragma unroll
        for (int c = 3; c < 15; c += 2) mxa[g] = max3(mxa[g], __uint_as_float(sr[16 * g + c]), __uint_as_float(sr[16 * g + c + 1]));
        mxa[g] = fmaxf(mxa[g], __uint_as_float(sr[16 * g + 15]));
      }
      float mx = mxa[0];
#pragma unroll
      for (int g = 1; g < kSub / 16; ++g) mx = fmaxf(mx, mxa[g]);
      if (kProbe & 2) mx = 0.f;
      float alpha = 1.0f;
      const bool grow = mx > m_ref + kRescaleThreshold;  // true on the first sub-step with an unmasked key (m_ref = -inf)
      if (grow) {
        alpha = ex2(m_ref - mx);  // 0 when m_ref = -inf
        m_ref = mx;
      }
      // p = 2^(s - m): packed f32x2 subtract and independent packed row-sum accumulators
      // (a row whose keys so far are all masked still has m_ref = -inf: subtract 0 so its p are 2^-inf = 0, not NaN)
      float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      KOCR_STAMP(i, 3);
      if (kPP) asm volatile("bar.sync %1, 64;" : "+f"(neg_m) : "r"(pp_mine) : "memory");  // the exponents depend on neg_m
      KOCR_STAMP(i, 4);
      const uint64_t neg_m2 = pack_f32x2(neg_m, neg_m);
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[kSub / 2];
#pragma unroll
      for (int c = 0; c < kSub / 2; ++c) {
        const uint64_t x2 = (kProbe & 32) ? pack_u32x2(sr[2 * c], sr[2 * c + 1]) : add_f32x2(pack_u32x2(sr[2 * c], sr[2 * c + 1]), neg_m2);
        uint64_t p2;
        if (kProbe & 1) {
          p2 = x2;
        } else if (kPolyEvery > 0 && (c % kPolyEvery) == kPolyEvery - 1) {
          p2 = ex2_poly_f32x2(x2);  // FMA/ALU pipes: relieves the MUFU unit
        } else {
          float x0, x1;
          unpack_f32x2(x2, x0, x1);
          p2 = pack_f32x2(ex2(x0), ex2(x1));
        }
        if (!(kProbe & 16) || c < 4) acc2[c & 3] = add_f32x2(acc2[c & 3], p2);
        float p0, p1;
        unpack_f32x2(p2, p0, p1);
        pk[c] = (kProbe & 64) ? (__float_as_uint(p0) ^ __float_as_uint(p1)) : pack_bf16(p0, p1);
        // hand the exponent phase to the other tile's warp once pair kPpAt is through the MUFU unit: the remaining
        // exponents cover the hand-over latency. The barrier id is made to depend on that pair's result (p >= 0, so the
        // sign bit adds nothing) because ptxas otherwise hoists the arrive to the top of the phase.
        if (kPP && c == kPpAt && (t == 0 || i + 1 < n_sub)) named_bar_arrive(pp_other + (int)(pk[c] >> 31), 64);
      }
      float sum;
      {
        float a0, a1, b0, b1;
        unpack_f32x2(add_f32x2(acc2[0], acc2[1]), a0, a1);
        unpack_f32x2(add_f32x2(acc2[2], acc2[3]), b0, b1);
        sum = (a0 + a1) + (b0 + b1);
      }
      l = l * alpha + sum;
      KOCR_STAMP(i, 5);
      if (!(kProbe & 8)) {
        tmem_st_x16(t_s, pk);
        tmem_st_x16(t_s + 16, pk + 16);
        if constexpr (kSub == 80) tmem_st_x8(t_s + 32, pk + 32);
      } else {
        uint32_t x = 0;  // keeps the conversions alive
#pragma unroll
        for (int c = 0; c < kSub / 2; ++c) x ^= pk[c];
        asm volatile("" ::"r"(x));
      }
      // P.V completions are signalled on one barrier per S/P buffer, o_done[t][b]; every completion is observed, in
      // order - P_{i-2} V here, normally long done - so a parity wait stays unambiguous.
      KOCR_STAMP(i, 6);
      if (i > 1) mbar_wait(&o_done[t * 2 + b], ph ^ 1);
      KOCR_STAMP(i, 7);
      if (i > 0) {
        // O_t holds P V of sub-steps < i relative to the old reference; bring it to the new one before P_i V is added. P_{i-1} V is waited for only when O really has to be rescaled.
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&o_done[t * 2 + (b ^ 1)], ((i - 1) >> 1) & 1);
          tc_fence_after();
          uint32_t o[kHd];
          tmem_ld_cols<kHd>(t_o, o);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < kHd; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_cols<kHd>(t_o, o);
        }
      }
      tc_wait_st();
      KOCR_STAMP(i, 8);
      tc_fence_before();
      mbar_arrive(&p_full[t * 2 + b]);
      KOCR_STAMP(i, 9);
    }
    // ---- epilogue: O / l -> bf16 -> out[row, head*80 ...]
    mbar_wait(&o_done[t * 2 + ((n_sub - 1) & 1)], ((n_sub - 1) >> 1) & 1);  // the last P.V (the pipe completes in order)
    tc_fence_after();
    uint32_t o[kHd];
    tmem_ld_cols<kHd>(t_o, o);
    tc_wait_ld();
    const float inv = 1.0f / l;
    const int qrow = t * kTileRows + r;
    if (qrow < w.q_rows) {
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(w.q_begin + qrow) * (num_heads * kHd) + head * kHd);
#pragma unroll
      for (int c = 0; c < kHd / 8; ++c) {
        const uint32_t* x = o + c * 8;
        dst[c] = make_uint4(pack_bf16(__uint_as_float(x[0]) * inv, __uint_as_float(x[1]) * inv),
                            pack_bf16(__uint_as_float(x[2]) * inv, __uint_as_float(x[3]) * inv),
                            pack_bf16(__uint_as_float(x[4]) * inv, __uint_as_float(x[5]) * inv),
                            pack_bf16(__uint_as_float(x[6]) * inv, __uint_as_float(x[7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Shape::kTmemCols>(tmem_base);
  }
}

int launch_attention(Ctx* ctx, const void* qkv, void* out, const AttnWork* d_work, int n_work, int num_heads,
                     int64_t total_rows, cudaStream_t stream, const int2* d_win) {
  if (n_work <= 0) return KOCR_OK;
  if (num_heads <= 0 || num_heads > 65535) return fail(KOCR_ERR_UNSUPPORTED, "attention: bad head count");
  const int sub = d_win ? AttnShape<true>::kSub : AttnShape<false>::kSub;
  CUtensorMap tm_q, tm_kv;
  uint64_t dims[2] = {(uint64_t)num_heads * 3 * kHd, (uint64_t)total_rows};
  uint64_t str[1] = {(uint64_t)num_heads * 3 * kHd * 2};
  uint32_t box_q[2] = {16, kTileRows}, box_kv[2] = {16, (uint32_t)sub};
  int rc = make_tensor_map(&tm_q, qkv, 2, dims, str, box_q, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  rc = make_tensor_map(&tm_kv, qkv, 2, dims, str, box_kv, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&attention_kernel<false>), AttnShape<false>::kSmem))) return rc;
  if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&attention_kernel<true>), AttnShape<true>::kSmem))) return rc;
  dim3 grid((unsigned)n_work, (unsigned)num_heads);
  if (d_win)
    attention_kernel<true><<<grid, AttnShape<true>::kThreads, AttnShape<true>::kSmem, stream>>>(tm_q, tm_kv, static_cast<__nv_bfloat16*>(out), d_work, num_heads, d_win);
  else
    attention_kernel<false><<<grid, AttnShape<false>::kThreads, AttnShape<false>::kSmem, stream>>>(tm_q, tm_kv, static_cast<__nv_bfloat16*>(out), d_work, num_heads, nullptr);
  KOCR_LAUNCH_CHECK("attention_kernel");
  return KOCR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Experimental shape (KOCR_ATTN3=1, full attention only): THREE query tiles per CTA with single-buffered scores. TMEM: S_t (80
// columns, P written over its first 40) and O_t (80) per tile = 480 columns. Three softmax warps per scheduler instead of two;
// in exchange a tile's next scores can only be issued behind P_t.V of the current sub-step, so each tile's chain contains the
// tensor pipe's latency and the other two tiles have to cover it. 512 threads: warp 0 TMA, warps 1-3 MMA (one per tile),
// warps 4-15 softmax (tile = (warp - 4) / 4, lane quarter = warp % 4).
struct Attn3 {
#ifndef KOCR_A3_STAGES
#define KOCR_A3_STAGES 4
#endif
  static constexpr int kSub = 80, kTiles = 3, kStages = KOCR_A3_STAGES;
  static constexpr int kKvChunkBytes = kSub * 32, kKvTileBytes = kChunks * kKvChunkBytes;
  static constexpr int kThreads = 128 + 128 * kTiles;
  static constexpr int kSmem = kTiles * kTileBytes + 2 * kStages * kKvTileBytes + 512 + 1024;
  __host__ __device__ static constexpr int s_col(int t) { return t * 160; }
  __host__ __device__ static constexpr int o_col(int t) { return t * 160 + 80; }
};

__global__ void __launch_bounds__(Attn3::kThreads, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, __nv_bfloat16* __restrict__ out,
                  const AttnWork* __restrict__ work, int num_heads) {
  constexpr int kTiles = Attn3::kTiles, kKvStages = Attn3::kStages, kSub = Attn3::kSub;
  constexpr int kKvChunkBytes = Attn3::kKvChunkBytes, kKvTileBytes = Attn3::kKvTileBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem_q + kTiles * kTileBytes;
  uint8_t* smem_v = smem_k + kKvStages * kKvTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKvStages * kKvTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kKvStages;
  uint64_t* v_full = k_empty + kKvStages;
  uint64_t* v_empty = v_full + kKvStages;
  uint64_t* s_full = v_empty + kKvStages;  // [tile]
  uint64_t* p_full = s_full + kTiles;      // [tile]
  uint64_t* o_last = p_full + kTiles;      // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_last + kTiles);

  const int warp = (int)uniform_u32(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const AttnWork w = work[blockIdx.x];
  const int head = blockIdx.y;
  const int n_sub = (w.kv_len + kSub - 1) / kSub;
  const int col_q = head * 3 * kHd, col_k = col_q + kHd, col_v = col_q + 2 * kHd;
  // a sequence's last block may hold fewer than 384 rows: tiles without a valid row do nothing at all (their MMA and softmax
  // warps go straight to the final barrier), so the block costs what its active tiles cost and stays next to its sequence's
  // other blocks in time - they share K/V through L2
  const int n_act = (w.q_rows + kTileRows - 1) / kTileRows;
#ifdef KOCR_TRACE
  const bool trace_on = blockIdx.x == 13 && blockIdx.y == 5;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], n_act);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], n_act);
    }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_last[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<64>();
    if (warp == 0) {
      const int kv_begin = (int)uniform_u32(w.kv_begin), q_begin = (int)uniform_u32(w.q_begin);
      const int n_sub_u = (int)uniform_u32(n_sub);
      const int n_act_u = (int)uniform_u32(n_act);
      if (elect_one()) {
        mbar_expect_tx(q_full, n_act_u * kTileBytes);
        for (int t = 0; t < n_act_u; ++t)
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_q + t * kTileBytes + c * kChunkBytes, &tm_q, q_full, col_q + c * 16, q_begin + t * kTileRows);
      }
      __syncwarp();
      for (int j = 0; j < n_sub_u; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        const int row = kv_begin + j * kSub;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[s], kKvTileBytes);
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_k + s * kKvTileBytes + c * kKvChunkBytes, &tm_kv, &k_full[s], col_k + c * 16, row);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&v_full[s], kKvTileBytes);
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_v + s * kKvTileBytes + c * kKvChunkBytes, &tm_kv, &v_full[s], col_v + c * 16, row);
        }
        __syncwarp();
      }
    } else if (warp - 1 < (int)uniform_u32(n_act)) {
      // MMA issuer of query tile t = warp - 1: S_t(0); then per sub-step P_t(i).V followed at once by S_t(i+1) - the tensor pipe
      // runs them in order, so the scores may overwrite P the moment P.V has read it
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kSub, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd, 0, 1);
      constexpr uint32_t hi32 = smem_desc_hi(256, 6);
      const int t = warp - 1;
      const uint32_t tmem_u = uniform_u32(tmem_base);
      const int n_sub_u = (int)uniform_u32(n_sub);
      const uint32_t q_lo = smem_desc_lo(smem_u32(smem_q), 16) + t * (kTileBytes >> 4);
      const uint32_t k_lo = smem_desc_lo(smem_u32(smem_k), 16);
      const uint32_t v_lo = smem_desc_lo(smem_u32(smem_v), kKvChunkBytes);
      const uint32_t d_s = tmem_u + Attn3::s_col(t), d_o = tmem_u + Attn3::o_col(t);
      auto issue_s = [&](int i) {  // the caller has seen k_full of this tile
        const int s = i % kKvStages;
        const uint32_t ka = k_lo + s * (kKvTileBytes >> 4);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          umma_ss_lo(d_s, q_lo + c * (kChunkBytes >> 4), ka + c * (kKvChunkBytes >> 4), hi32, idesc_s, c != 0);
        tc_commit_elect(&s_full[t]);
        tc_commit_elect(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int i = 0; i < n_sub_u; ++i) {
        const int s = i % kKvStages;
        KOCR_STAMP(i, 0);
        mbar_wait(&v_full[s], (i / kKvStages) & 1);
        if (i + 1 < n_sub_u) mbar_wait(&k_full[(i + 1) % kKvStages], ((i + 1) / kKvStages) & 1);  // off the critical path: before P arrives
        KOCR_STAMP(i, 3);
        mbar_wait(&p_full[t], i & 1);
        KOCR_STAMP(i, 4);
        tc_fence_after();
        const uint32_t va = v_lo + s * (kKvTileBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kSub / 16; ++ks)
          umma_ts_lo(d_o, d_s + ks * 8, va + ks * (512 >> 4), hi32, idesc_o, (i > 0 || ks != 0));
        tc_commit_elect(&v_empty[s]);
        KOCR_STAMP(i, 1);
        if (i + 1 < n_sub_u) issue_s(i + 1);
        else tc_commit_elect(&o_last[t]);
        KOCR_STAMP(i, 2);
      }
    }
  } else {
    setmaxnreg_inc<144>();
    const int t = (warp - 4) >> 2;
    if (t < n_act) {
    const int qtr = warp & 3;
    const int r = qtr * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
    const uint32_t t_s = tmem_base + Attn3::s_col(t) + lane_off;
    const uint32_t t_o = tmem_base + Attn3::o_col(t) + lane_off;
    float m_ref = -INFINITY, l = 0.f;
#ifndef KOCR_PP3
#define KOCR_PP3 0
#endif
    // exponent-phase token around the three tiles' warps of one lane quarter (same scheduler, same MUFU unit): 0 -> 1 -> 2 -> 0
    constexpr bool kPP3 = KOCR_PP3;
#ifndef KOCR_A3_POLY
#define KOCR_A3_POLY 4
#endif
    constexpr int kPoly3 = KOCR_A3_POLY;
    const int pp_mine = 1 + t * 4 + qtr, pp_next = 1 + ((t + 1) % kTiles) * 4 + qtr;
    if (kPP3 && t == kTiles - 1) named_bar_arrive(pp_next, 64);  // tile 0 goes first
    for (int i = 0; i < n_sub; ++i) {
      KOCR_STAMP(i, 0);
      mbar_wait(&s_full[t], i & 1);  // S_t(i) complete; it was issued behind P_t(i-1).V, which is therefore complete as well
      tc_fence_after();
      uint32_t sr[kSub];
      tmem_ld_x32(t_s, sr);
      tmem_ld_x32(t_s + 32, sr + 32);
      tmem_ld_x16(t_s + 64, sr + 64);
      KOCR_STAMP(i, 1);
      tc_wait_ld();
      KOCR_STAMP(i, 2);
      const int valid = w.kv_len - i * kSub;
      if (valid < kSub) {
#pragma unroll
        for (int c = 0; c < kSub; ++c)
          if (c >= valid) sr[c] = 0xff800000u;
      }
#ifndef KOCR_A3_TREE
#define KOCR_A3_TREE 0
#endif
      float mx;
      if constexpr (KOCR_A3_TREE) {
        mx = max_tree<kSub, uint32_t>(sr);
      } else {
        float mxa[kSub / 16];
#pragma unroll
        for (int g = 0; g < kSub / 16; ++g) {
          mxa[g] = max3(__uint_as_float(sr[16 * g]), __uint_as_float(sr[16 * g + 1]), __uint_as_float(sr[16 * g + 2]));
#pragma unroll
          for (int c = 3; c < 15; c += 2) mxa[g] = max3(mxa[g], __uint_as_float(sr[16 * g + c]), __uint_as_float(sr[16 * g + c + 1]));
          mxa[g] = fmaxf(mxa[g], __uint_as_float(sr[16 * g + 15]));
        }
        mx = mxa[0];
#pragma unroll
        for (int g = 1; g < kSub / 16; ++g) mx = fmaxf(mx, mxa[g]);
      }
      float alpha = 1.0f;
      const bool grow = mx > m_ref + kRescaleThreshold;
      if (grow) {
        alpha = ex2(m_ref - mx);
        m_ref = mx;
      }
      float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      KOCR_STAMP(i, 3);
      KOCR_STAMP(i, 4);
      if (kPP3) asm volatile("bar.sync %1, 64;" : "+f"(neg_m) : "r"(pp_mine) : "memory");
      const uint64_t neg_m2 = pack_f32x2(neg_m, neg_m);
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[kSub / 2];
#pragma unroll
      for (int c = 0; c < kSub / 2; ++c) {
        const uint64_t x2 = add_f32x2(pack_u32x2(sr[2 * c], sr[2 * c + 1]), neg_m2);
        uint64_t p2;
        if (kPoly3 > 0 && (c % kPoly3) == kPoly3 - 1) {
          p2 = ex2_poly_f32x2(x2);
        } else {
          float x0, x1;
          unpack_f32x2(x2, x0, x1);
          p2 = pack_f32x2(ex2(x0), ex2(x1));
        }
        acc2[c & 3] = add_f32x2(acc2[c & 3], p2);
        float p0, p1;
        unpack_f32x2(p2, p0, p1);
        pk[c] = pack_bf16(p0, p1);
        if (kPP3 && c == kPpAt && (t != kTiles - 1 || i + 1 < n_sub)) named_bar_arrive(pp_next + (int)(pk[c] >> 31), 64);
#ifndef KOCR_A3_EARLYST
#define KOCR_A3_EARLYST 0
#endif
        if (KOCR_A3_EARLYST && c == 15) tmem_st_x16(t_s, pk);       // P leaves in pieces: the stores' latency overlaps the remaining exponents
        if (KOCR_A3_EARLYST && c == 31) tmem_st_x16(t_s + 16, pk + 16);
      }
      float sum;
      {
        float a0, a1, b0, b1;
        unpack_f32x2(add_f32x2(acc2[0], acc2[1]), a0, a1);
        unpack_f32x2(add_f32x2(acc2[2], acc2[3]), b0, b1);
        sum = (a0 + a1) + (b0 + b1);
      }
      l = l * alpha + sum;
      KOCR_STAMP(i, 5);
      if (!KOCR_A3_EARLYST) {
        tmem_st_x16(t_s, pk);
        tmem_st_x16(t_s + 16, pk + 16);
      }
      tmem_st_x8(t_s + 32, pk + 32);
      if (i > 0 && __any_sync(0xffffffffu, grow)) {  // P_t(i-1).V is complete (see the wait above): O_t may be rescaled right away
        uint32_t o[kHd];
        tmem_ld_cols<kHd>(t_o, o);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < kHd; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
        tmem_st_cols<kHd>(t_o, o);
      }
      KOCR_STAMP(i, 6);
      KOCR_STAMP(i, 7);
      tc_wait_st();
      KOCR_STAMP(i, 8);
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      KOCR_STAMP(i, 9);
    }
    mbar_wait(&o_last[t], 0);
    tc_fence_after();
    uint32_t o[kHd];
    tmem_ld_cols<kHd>(t_o, o);
    tc_wait_ld();
    const float inv = 1.0f / l;
    const int qrow = t * kTileRows + r;
    if (qrow < w.q_rows) {
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(w.q_begin + qrow) * (num_heads * kHd) + head * kHd);
#pragma unroll
      for (int c = 0; c < kHd / 8; ++c) {
        const uint32_t* x = o + c * 8;
        dst[c] = make_uint4(pack_bf16(__uint_as_float(x[0]) * inv, __uint_as_float(x[1]) * inv),
                            pack_bf16(__uint_as_float(x[2]) * inv, __uint_as_float(x[3]) * inv),
                            pack_bf16(__uint_as_float(x[4]) * inv, __uint_as_float(x[5]) * inv),
                            pack_bf16(__uint_as_float(x[6]) * inv, __uint_as_float(x[7]) * inv));
      }
    }
    }  // active tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_attention3(Ctx* ctx, const void* qkv, void* out, const AttnWork* d_work, int n_work, int num_heads, int64_t total_rows,
                             cudaStream_t stream) {
  CUtensorMap tm_q, tm_kv;
  uint64_t dims[2] = {(uint64_t)num_heads * 3 * kHd, (uint64_t)total_rows};
  uint64_t str[1] = {(uint64_t)num_heads * 3 * kHd * 2};
  uint32_t box_q[2] = {16, kTileRows}, box_kv[2] = {16, (uint32_t)Attn3::kSub};
  int rc = make_tensor_map(&tm_q, qkv, 2, dims, str, box_q, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  rc = make_tensor_map(&tm_kv, qkv, 2, dims, str, box_kv, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&attention3_kernel), Attn3::kSmem))) return rc;
  dim3 grid((unsigned)n_work, (unsigned)num_heads);
  attention3_kernel<<<grid, Attn3::kThreads, Attn3::kSmem, stream>>>(tm_q, tm_kv, static_cast<__nv_bfloat16*>(out), d_work, num_heads);
  KOCR_LAUNCH_CHECK("attention3_kernel");
  return KOCR_OK;
}

// Host: full attention = 384-row blocks on the three-tile kernel; a sequence's last block holds what is left (tiles without rows
// are skipped by the kernel). build_attn_work_mixed (three-tile blocks plus up to two 256-row blocks for the two-tile kernel) pads
// less but makes a second launch that re-reads every sequence's K/V from HBM (ncu: +2.9 GB per layer): kept for A/B runs only.
int build_attn_work3(const int32_t* cu, int n_seqs, std::vector<AttnWork>* out) {
  out->clear();
  for (int i = 0; i < n_seqs; ++i) {
    const int b = cu[i], len = cu[i + 1] - cu[i];
    if (len <= 0) return fail(KOCR_ERR_INVALID, "attention: empty or negative sequence");
    for (int q = 0; q < len; q += 3 * kTileRows) out->push_back(AttnWork{b + q, std::min(3 * kTileRows, len - q), b, len});
  }
  std::stable_sort(out->begin(), out->end(), [](const AttnWork& x, const AttnWork& y) { return x.kv_len > y.kv_len; });
  return KOCR_OK;
}

int build_attn_work_mixed(const int32_t* cu, int n_seqs, std::vector<AttnWork>* w3, std::vector<AttnWork>* w2) {
  w3->clear();
  w2->clear();
  constexpr int kB3 = 3 * kTileRows, kB2 = 2 * kTileRows;
  for (int i = 0; i < n_seqs; ++i) {
    const int b = cu[i], len = cu[i + 1] - cu[i];
    if (len <= 0) return fail(KOCR_ERR_INVALID, "attention: empty or negative sequence");
    int na = 0, nb = 0, best_pad = INT32_MAX;
    for (int tb = 0; tb <= 2; ++tb) {
      const int rest = len - tb * kB2;
      const int ta = rest > 0 ? (rest + kB3 - 1) / kB3 : 0;
      const int pad = ta * kB3 + tb * kB2 - len;
      if (pad >= 0 && pad < best_pad) { best_pad = pad; na = ta; nb = tb; }
    }
    int q = 0;
    for (int k = 0; k < na; ++k) {  // full blocks when two-tile blocks follow (they take the remainder), else the last one is short
      const int rows = std::min(kB3, len - q);
      w3->push_back(AttnWork{b + q, rows, b, len});
      q += rows;
    }
    for (int k = 0; k < nb; ++k) {  // the remainder, spread evenly
      const int rows = (len - q + (nb - k) - 1) / (nb - k);
      w2->push_back(AttnWork{b + q, rows, b, len});
      q += rows;
    }
    if (q != len) return fail(KOCR_ERR_INVALID, "attention: work split does not cover the sequence");
  }
  // longest sequences first: their blocks are the longest-running CTAs, the short ones fill the tail of the launch
  auto by_len = [](const AttnWork& x, const AttnWork& y) { return x.kv_len > y.kv_len; };
  std::stable_sort(w3->begin(), w3->end(), by_len);
  std::stable_sort(w2->begin(), w2->end(), by_len);
  return KOCR_OK;
}

// Host: split sequences into 256-row query blocks (one CTA each per head)
int build_attn_work(const int32_t* cu, int n_seqs, std::vector<AttnWork>* out) {
  out->clear();
  for (int i = 0; i < n_seqs; ++i) {
    const int b = cu[i], len = cu[i + 1] - cu[i];
    if (len <= 0) return fail(KOCR_ERR_INVALID, "attention: empty or negative sequence");
    for (int q = 0; q < len; q += 2 * kTileRows) out->push_back(AttnWork{b + q, std::min(2 * kTileRows, len - q), b, len});
  }
  return KOCR_OK;
}

// Host: windowed layers. Sequences are whole images (cu), windows are the contiguous segments cu_win inside them;
// query blocks of 256 rows per image, each with the key range spanned by its rows' windows, plus the per-row table.
int build_attn_work_windowed(const int32_t* cu, int n_seqs, const int32_t* cu_win, int n_win, std::vector<AttnWork>* out,
                             std::vector<int32_t>* row_win) {
  out->clear();
  const int total = cu[n_seqs];
  row_win->assign((size_t)total * 2, 0);
  for (int k = 0; k < n_win; ++k) {
    if (cu_win[k + 1] <= cu_win[k] || cu_win[k + 1] > total) return fail(KOCR_ERR_INVALID, "attention: bad window table");
    for (int r = cu_win[k]; r < cu_win[k + 1]; ++r) {
      (*row_win)[2 * (size_t)r] = cu_win[k];
      (*row_win)[2 * (size_t)r + 1] = cu_win[k + 1];
    }
  }
  for (int i = 0; i < n_seqs; ++i) {
    const int b = cu[i], e = cu[i + 1];
    for (int q = b; q < e; q += kTileRows) {  // one query tile per CTA on windowed layers (AttnShape<true>)
      const int rows = std::min(kTileRows, e - q);
      const int kv0 = (*row_win)[2 * (size_t)q], kv1 = (*row_win)[2 * (size_t)(q + rows - 1) + 1];
      out->push_back(AttnWork{q, rows, kv0, kv1 - kv0});
    }
  }
  return KOCR_OK;
}

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_op_attention(KocrCtx* ctx_, const void* qkv, void* out, const int32_t* cu_seqlens_host, int n_seqs,
                                 int num_heads, int head_dim, void* stream_) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!ctx || !qkv || !out || !cu_seqlens_host || n_seqs <= 0) return fail(KOCR_ERR_INVALID, "kocr_op_attention: bad argument");
  if (head_dim != kHd) return fail(KOCR_ERR_UNSUPPORTED, "kocr_op_attention: kernels are built for head_dim 80");
  reset_launch_count();
  std::vector<AttnWork> work, work3;
  static const bool two_tile_only = getenv("KOCR_ATTN2") != nullptr;   // A/B switches: the two-tile kernel alone,
  static const bool mixed = getenv("KOCR_ATTN_MIXED") != nullptr;       // three-tile blocks + two-tile remainder blocks
  int rc = two_tile_only ? build_attn_work(cu_seqlens_host, n_seqs, &work)
           : mixed       ? build_attn_work_mixed(cu_seqlens_host, n_seqs, &work3, &work)
                         : build_attn_work3(cu_seqlens_host, n_seqs, &work3);
  if (rc) return rc;
  std::vector<AttnWork> all(work3);
  all.insert(all.end(), work.begin(), work.end());
  void* d_work;
  int slot = -1;
  rc = ctx->stage(all.data(), all.size() * sizeof(AttnWork), stream, &d_work, &slot);
  if (rc) return rc;
  StageGuard guard(ctx, slot, stream);  // released after the kernels that read the work list are enqueued
  const AttnWork* dw = static_cast<const AttnWork*>(d_work);
  if (!work3.empty() && (rc = launch_attention3(ctx, qkv, out, dw, (int)work3.size(), num_heads, cu_seqlens_host[n_seqs], stream))) return rc;
  if (work.empty()) return KOCR_OK;
  return launch_attention(ctx, qkv, out, dw + work3.size(), (int)work.size(), num_heads, cu_seqlens_host[n_seqs], stream, nullptr);
}

#ifdef KOCR_TRACE
extern "C" __attribute__((visibility("default"))) int kocr_debug_attn_trace(void* dst, int64_t bytes) {
  if (bytes < (int64_t)sizeof(long long) * kTraceWarps * kTraceSubs * kTracePts) return KOCR_ERR_INVALID;
  return cudaMemcpyFromSymbol(dst, g_attn_trace, sizeof(long long) * kTraceWarps * kTraceSubs * kTracePts) == cudaSuccess ? KOCR_OK : KOCR_ERR_CUDA;
}
#endif
