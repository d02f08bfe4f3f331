// Varlen non-causal flash attention for head_dim 80 on tcgen05 / TMEM / TMA (sm_100a).
//
// Stands in for VisionAttention's attention call (HF models/qwen2_vl/modeling_qwen2_vl.py:411-454: one
// softmax(q k^T / sqrt(80)) v per cu_seqlens segment, non-causal), which HF runs as a Python loop of SDPA calls or as
// flash_attn_varlen_func.
//
// Input is the tower's packed QKV buffer [S, heads*240] with, per head, 80 q | 80 k | 80 v columns; q is already
// rotated and multiplied by head_dim^-0.5 * log2(e) (fused QKV GEMM epilogue), so scores are in the log2 domain.
//
// CTA = one 256-row query block of one sequence and one head; 384 threads:
//   warp 0      TMA producer: Q once, then a ring of K tiles and a ring of V tiles (128 keys each)
//   warp 1      MMA issuer:   S_t = Q_t K^T (128x128x80, SS) and O_t += P_t V (128x80x128, P from TMEM, V MN-major)
//   warps 4-7   softmax of query tile 0 (rows 0..127), one row per thread
//   warps 8-11  softmax of query tile 1 (rows 128..255)
// Operand tiles are stored as five [128 rows x 16 columns] SWIZZLE_32B chunks, which is a canonical K-major layout for
// Q/K (K = head_dim) and, unchanged, a canonical MN-major layout for V (N = head_dim): no transpose, no padding of 80.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,336) O1 [384,464); P_t (bf16) overwrites the first 64 columns of S_t.
// The two query tiles ping-pong on the tensor pipe: while softmax works on S_0 the pipe runs P_1 V and the next S_1.
#include <algorithm>
#include <vector>

#include "kocr_common.cuh"
#include "kocr_kernels.h"

namespace kocr {

static constexpr int kHd = 80;
static constexpr int kChunks = kHd / 16;           // 5 chunks of 16 columns
static constexpr int kTileRows = 128;
static constexpr int kChunkBytes = kTileRows * 32;  // 4096
static constexpr int kTileBytes = kChunks * kChunkBytes;  // 20480
static constexpr int kKvStages = 3;
static constexpr int kAttnThreads = 384;
static constexpr int kAttnSmem = 2 * kTileBytes + 2 * kKvStages * kTileBytes + 512 + 1024;
static constexpr float kRescaleThreshold = 8.0f;  // log2 units: rescale O only when the row max grows by more than 2^8

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

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                 const AttnWork* __restrict__ work, int num_heads) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 tiles
  uint8_t* smem_k = smem + 2 * kTileBytes;                  // kKvStages tiles
  uint8_t* smem_v = smem_k + kKvStages * kTileBytes;        // kKvStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKvStages * kTileBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // kKvStages
  uint64_t* k_empty = k_full + kKvStages;  // kKvStages
  uint64_t* v_full = k_empty + kKvStages;
  uint64_t* v_empty = v_full + kKvStages;
  uint64_t* s_full = v_empty + kKvStages;  // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* o_done = p_full + 2;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const AttnWork w = work[blockIdx.x];
  const int head = blockIdx.y;
  const int n_kv = (w.kv_len + kTileRows - 1) / kTileRows;
  const int col_q = head * 3 * kHd, col_k = col_q + kHd, col_v = col_q + 2 * kHd;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<96>();
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, 2 * kTileBytes);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < kChunks; ++c)
          tma_load_2d(smem_q + t * kTileBytes + c * kChunkBytes, &tm_qkv, q_full, col_q + c * 16,
                      w.q_begin + t * kTileRows);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        const int row = w.kv_begin + j * kTileRows;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kTileBytes);
        for (int c = 0; c < kChunks; ++c)
          tma_load_2d(smem_k + s * kTileBytes + c * kChunkBytes, &tm_qkv, &k_full[s], col_k + c * 16, row);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kTileBytes);
        for (int c = 0; c < kChunks; ++c)
          tma_load_2d(smem_v + s * kTileBytes + c * kChunkBytes, &tm_qkv, &v_full[s], col_v + c * 16, row);
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd, 0, 1);   // P (TMEM) x V (MN-major)
      auto issue_s = [&](int t, int s) {
        const uint32_t qa = smem_u32(smem_q + t * kTileBytes), ka = smem_u32(smem_k + s * kTileBytes);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const uint64_t da = make_smem_desc(qa + c * kChunkBytes, 16, 256, 6);
          const uint64_t db = make_smem_desc(ka + c * kChunkBytes, 16, 256, 6);
          umma_ss(tmem_base + t * 128, da, db, idesc_s, c != 0);
        }
      };
      auto issue_pv = [&](int t, int s, bool accumulate) {
        const uint32_t va = smem_u32(smem_v + s * kTileBytes);
#pragma unroll
        for (int ks = 0; ks < kTileRows / 16; ++ks) {
          // 16 keys per step: rows ks*16.. of every chunk (512 B further); N groups of 16 columns are 4096 B apart
          const uint64_t db = make_smem_desc(va + ks * 512, kChunkBytes, 256, 6);
          umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + ks * 8, db, idesc_o, (accumulate || ks != 0));
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      tc_commit(&s_full[0]);
      issue_s(1, 0);
      tc_commit(&s_full[1]);
      tc_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        const int s1 = (j + 1) % kKvStages;
        const uint32_t ph1 = ((j + 1) / kKvStages) & 1;
        mbar_wait(&v_full[s], ph);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          issue_pv(t, s, j > 0);
          tc_commit(&o_done[t]);
          if (j + 1 < n_kv) {
            if (t == 0) {
              mbar_wait(&k_full[s1], ph1);
              tc_fence_after();
            }
            issue_s(t, s1);
            tc_commit(&s_full[t]);
            if (t == 1) tc_commit(&k_empty[s1]);
          }
        }
        tc_commit(&v_empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<200>();  // 128*96 + 256*200 <= 384*168 (the launch allocation): inc can never starve
    const int t = (warp - 4) >> 2;   // query tile
    const int qtr = warp & 3;        // TMEM lane quarter
    const int r = qtr * 32 + lane;   // row within the tile
    const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
    const uint32_t t_s = tmem_base + t * 128 + lane_off;
    const uint32_t t_o = tmem_base + 256 + t * 128 + lane_off;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t sr[128];
      tmem_ld_x32(t_s, sr);
      tmem_ld_x32(t_s + 32, sr + 32);
      tmem_ld_x32(t_s + 64, sr + 64);
      tmem_ld_x32(t_s + 96, sr + 96);
      tc_wait_ld();
      const int valid = w.kv_len - j * kTileRows;
      if (valid < kTileRows) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) sr[c] = 0xff800000u;  // -inf
      }
      // row max: 4 independent FMNMX3 chains (a single chain of 127 dependent max ops would cost ~500 cycles)
      float mxa[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        mxa[g] = max3(__uint_as_float(sr[32 * g]), __uint_as_float(sr[32 * g + 1]), __uint_as_float(sr[32 * g + 2]));
#pragma unroll
        for (int c = 3; c < 31; c += 2) mxa[g] = max3(mxa[g], __uint_as_float(sr[32 * g + c]), __uint_as_float(sr[32 * g + c + 1]));
        mxa[g] = fmaxf(mxa[g], __uint_as_float(sr[32 * g + 31]));
      }
      const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      float alpha = 1.0f;
      const bool grow = mx > m_ref + kRescaleThreshold;  // always true on the first tile (m_ref = -inf)
      if (grow) {
        alpha = ex2(m_ref - mx);  // 0 on the first tile
        m_ref = mx;
      }
      // p = 2^(s - m): packed f32x2 subtract and 4 independent packed row-sum accumulators
      const uint64_t neg_m2 = pack_f32x2(-m_ref, -m_ref);
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float x0, x1;
        unpack_f32x2(add_f32x2(pack_u32x2(sr[2 * c], sr[2 * c + 1]), neg_m2), x0, x1);
        const float p0 = ex2(x0), p1 = ex2(x1);
        acc2[c & 3] = add_f32x2(acc2[c & 3], pack_f32x2(p0, p1));
        pk[c] = pack_bf16(p0, p1);
      }
      float sum;
      {
        float a0, a1, b0, b1;
        unpack_f32x2(add_f32x2(acc2[0], acc2[1]), a0, a1);
        unpack_f32x2(add_f32x2(acc2[2], acc2[3]), b0, b1);
        sum = (a0 + a1) + (b0 + b1);
      }
      l = l * alpha + sum;
      tmem_st_x16(t_s, pk);
      tmem_st_x16(t_s + 16, pk + 16);
      tmem_st_x16(t_s + 32, pk + 32);
      tmem_st_x16(t_s + 48, pk + 48);
      if (j > 0) {
        // O_t holds P V of tiles < j relative to the old reference; bring it to the new one before P_j V is added
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&o_done[t], (j - 1) & 1);
          tc_fence_after();
          uint32_t o[80];
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_ld_x16(t_o + c * 16, o + c * 16);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 80; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
#pragma unroll
          for (int c = 0; c < 5; ++c) tmem_st_x16(t_o + c * 16, o + c * 16);
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> bf16 -> out[row, head*80 ..]
    mbar_wait(&o_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    uint32_t o[80];
#pragma unroll
    for (int c = 0; c < 5; ++c) tmem_ld_x16(t_o + c * 16, o + c * 16);
    tc_wait_ld();
    const float inv = 1.0f / l;
    const int qrow = t * kTileRows + r;
    if (qrow < w.q_rows) {
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(w.q_begin + qrow) * (num_heads * kHd) + head * kHd);
#pragma unroll
      for (int c = 0; c < 10; ++c) {
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
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_attention(Ctx* ctx, const void* qkv, void* out, const AttnWork* d_work, int n_work, int num_heads,
                     int64_t total_rows, cudaStream_t stream) {
  if (n_work <= 0) return KOCR_OK;
  if (num_heads <= 0 || num_heads > 65535) return fail(KOCR_ERR_UNSUPPORTED, "attention: bad head count");
  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)num_heads * 3 * kHd, (uint64_t)total_rows};
  uint64_t str[1] = {(uint64_t)num_heads * 3 * kHd * 2};
  uint32_t box[2] = {16, kTileRows};
  int rc = make_tensor_map(&tm, qkv, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    KOCR_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr_set = true;
  }
  dim3 grid((unsigned)n_work, (unsigned)num_heads);
  attention_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tm, static_cast<__nv_bfloat16*>(out), d_work, num_heads);
  KOCR_LAUNCH_CHECK("attention_kernel");
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

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_op_attention(KocrCtx* ctx_, const void* qkv, void* out, const int32_t* cu_seqlens_host, int n_seqs,
                                 int num_heads, int head_dim, void* stream_) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!ctx || !qkv || !out || !cu_seqlens_host || n_seqs <= 0) return fail(KOCR_ERR_INVALID, "kocr_op_attention: bad argument");
  if (head_dim != kHd) return fail(KOCR_ERR_UNSUPPORTED, "kocr_op_attention: kernels are built for head_dim 80");
  reset_launch_count();
  std::vector<AttnWork> work;
  int rc = build_attn_work(cu_seqlens_host, n_seqs, &work);
  if (rc) return rc;
  void* d_work;
  rc = ctx->stage(work.data(), work.size() * sizeof(AttnWork), stream, &d_work);
  if (rc) return rc;
  return launch_attention(ctx, qkv, out, static_cast<const AttnWork*>(d_work), (int)work.size(), num_heads,
                          cu_seqlens_host[n_seqs], stream);
}
