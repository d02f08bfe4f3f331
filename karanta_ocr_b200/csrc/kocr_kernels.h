// Internal launcher interface shared by the kernel files and the tower orchestration.
#pragma once
#include "kocr_common.cuh"

namespace kocr {

static constexpr int kEpiQkvRope = 100;  // internal: bias + 2D RoPE + q pre-scale, BN = 240 = [q_h|k_h|v_h]

struct GemmEpilogue {
  const float* bias = nullptr;             // [N] f32 (in the prepacked column order)
  const __nv_bfloat16* residual = nullptr; // [M, ld_res]
  int64_t ld_res = 0;
  __nv_bfloat16* out = nullptr;            // [M, ldc]
  int64_t ldc = 0;
  // kEpiQkvRope only
  const int2* pos_hw = nullptr;            // [M] (row, col) of each patch
  const float2* rope_cs = nullptr;         // [max_pos][head_dim/4] (cos, sin) of pos * inv_freq[j]
  float q_scale = 1.0f;                    // head_dim^-0.5 * log2(e), folded into q
  // Norm folded into this GEMM (A = raw residual stream, B = W*diag(gamma)):
  //   out = rstd[m] * (acc - mean[m] * ln_c1[n]) + bias[n],  bias already holding b + W.beta; mean/rstd from ln_part.
  const float2* ln_part = nullptr;         // [M][ln_slots] partial (sum, sum of squares) of each row of A
  const float* ln_c1 = nullptr;            // [N] row sums of the gamma-scaled, bf16-rounded weight
  int ln_slots = 0;
  int ln_rms = 0;                          // RMSNorm: no mean subtraction
  float ln_inv_dim = 0.f, ln_eps = 0.f;
  // Producer side: the rows this GEMM stores are the next norm's input; each (n-tile, column half) writes the partial
  // (sum, sum of squares) of the bf16 values it stored to its own slot (no atomics: deterministic).
  float2* stat_part = nullptr;             // [M][stat_slots], slot = n_tile * 2 + half
  int stat_slots = 0;
};

int launch_gemm(Ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                int epi, const GemmEpilogue& ep, cudaStream_t stream);

// y = norm(x) * w (+ b); one warp per row; dim % 8 == 0, dim <= 8192. rows of bf16, row pitches in elements.
int launch_norm(const void* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int64_t rows, int dim,
                float eps, bool rms, cudaStream_t stream);

// f32 -> bf16 cast (pixel_values drop-in path), n elements, n % 4 == 0
int launch_cast_f32_bf16(const float* x, void* y, int64_t n, cudaStream_t stream);
// any supported dtype -> f32 / bf16 (weight prepack)
int launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, cudaStream_t stream);
// dst[r, :] = src[perm[r], :] for rows of `cols` elements of `elt` bytes (weight row permutation), perm on device
int launch_permute_rows(const void* src, void* dst, const int32_t* perm, int64_t rows, int64_t cols, int64_t src_ld,
                        int64_t dst_ld, int elt, cudaStream_t stream);
// LayerNorm folding (weight prepack): per row n of W [N,K] (bf16, in place): bias[n] += sum_k beta[k]*W[n,k] (beta may
// be null), W[n,k] = bf16(W[n,k]*gamma[k]), c1[n] = sum_k W'[n,k]. gamma/beta f32 [K].
int launch_fold_norm(void* W, int64_t ldw, float* bias, float* c1, const float* gamma, const float* beta, int64_t N, int64_t K,
                     cudaStream_t stream);
// rope table: cs[p][j] = (cos, sin)(p * inv_freq[j]), p < max_pos, j < n_freq
int launch_rope_table(float2* cs, int max_pos, int n_freq, float theta, cudaStream_t stream);
// gather groups of `group` rows: dst[g] = src[index[g]] (window permutation and its inverse), rows of `cols` bf16
int launch_gather_groups(const void* src, void* dst, const int32_t* index, int64_t n_groups, int group, int cols,
                         bool inverse, cudaStream_t stream);

// varlen attention over the tower's QKV layout [S, heads, 3, 80]; seq table on device
struct AttnWork { int q_begin; int q_rows; int kv_begin; int kv_len; };  // one CTA work item (a 256-row q block of one sequence)
// d_win != nullptr: windowed mode, d_win[row] = (first key row, one past the last key row) of the row's window
int launch_attention(Ctx* ctx, const void* qkv, void* out, const AttnWork* d_work, int n_work, int num_heads,
                     int64_t total_rows, cudaStream_t stream, const int2* d_win = nullptr);

// full attention, three query tiles per CTA (384-row blocks); launch_attention (two tiles) takes what build_attn_work_mixed leaves it
int launch_attention3(Ctx* ctx, const void* qkv, void* out, const AttnWork* d_work, int n_work, int num_heads, int64_t total_rows,
                      cudaStream_t stream);

}  // namespace kocr
