// Memory-bound helper kernels of the tower: LayerNorm / RMSNorm (HF modeling_qwen2_vl.py:464-465,
// modeling_qwen2_5_vl.py:57-71), dtype casts (PatchEmbed.forward :306-309 `.to(dtype)`), the RoPE angle table
// (VisionRotaryEmbedding :271-284), window gather (modeling_qwen2_5_vl.py:478-484,:512-513) and weight prepack.
// All use 16-byte vector accesses with consecutive lanes on consecutive addresses.
#include <cuda_fp16.h>

#include <algorithm>

#include "kocr_common.cuh"
#include "kocr_kernels.h"

namespace kocr {

// ---------------------------------------------------------------- LayerNorm / RMSNorm: one warp per row
template <int kVecPerLane, bool kRms>
__global__ void __launch_bounds__(256) norm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                   __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t rows, int dim,
                                                   float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  const int nvec = dim >> 3;
  float v[kVecPerLane][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = xr[vi];
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[i][2 * k] = bf16_lo(uu[k]);
        v[i][2 * k + 1] = bf16_hi(uu[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += kRms ? v[i][k] * v[i][k] : v[i][k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float mean = 0.f, rstd;
  if (kRms) {
    rstd = rsqrtf(sum / dim + eps);
  } else {
    mean = sum / dim;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = v[i][k] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd = rsqrtf(sq / dim + eps);
  }
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi);
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * vi + 1);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float o[8];
      if (kRms) {
        // HF: weight * (x * rsqrt(var + eps)).to(input_dtype)
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ww[k] * __bfloat162float(__float2bfloat16_rn(v[i][k] * rstd));
      } else {
        float bb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (b) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * vi);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * vi + 1);
          bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
          bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * ww[k] + bb[k];
      }
      yr[vi] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

int launch_norm(const void* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int64_t rows, int dim,
                float eps, bool rms, cudaStream_t stream) {
  if (dim % 8 || dim <= 0 || dim > 8192 || ldx % 8 || ldy % 8) return fail(KOCR_ERR_UNSUPPORTED, "norm: dim must be a multiple of 8 and <= 8192");
  if (rows <= 0) return KOCR_OK;
  const int nvec = dim / 8;
  const int vpl = (nvec + 31) / 32;
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  auto xp = static_cast<const __nv_bfloat16*>(x);
  auto yp = static_cast<__nv_bfloat16*>(y);
#define KOCR_NORM_CASE(V)                                                                                     \
  if (vpl <= V) {                                                                                             \
    if (rms) norm_kernel<V, true><<<grid, warps * 32, 0, stream>>>(xp, ldx, w, b, yp, ldy, rows, dim, eps);   \
    else norm_kernel<V, false><<<grid, warps * 32, 0, stream>>>(xp, ldx, w, b, yp, ldy, rows, dim, eps);      \
    KOCR_LAUNCH_CHECK("norm_kernel");                                                                         \
    return KOCR_OK;                                                                                           \
  }
  KOCR_NORM_CASE(1)
  KOCR_NORM_CASE(2)
  KOCR_NORM_CASE(5)
  KOCR_NORM_CASE(8)
  KOCR_NORM_CASE(16)
  KOCR_NORM_CASE(32)
#undef KOCR_NORM_CASE
  return fail(KOCR_ERR_UNSUPPORTED, "norm: dim too large");
}

// ---------------------------------------------------------------- casts
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

int launch_cast_f32_bf16(const float* x, void* y, int64_t n, cudaStream_t stream) {
  if (n % 4) return fail(KOCR_ERR_UNSUPPORTED, "cast: element count must be a multiple of 4");
  if (n == 0) return KOCR_OK;
  const int64_t n4 = n / 4;
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, 148 * 16);
  cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), static_cast<uint2*>(y), n4);
  KOCR_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return KOCR_OK;
}

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
  if (dtype == KOCR_DTYPE_F32) return static_cast<const float*>(p)[i];
  if (dtype == KOCR_DTYPE_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(static_cast<const __half*>(p)[i]);
}

__global__ void __launch_bounds__(256) convert_kernel(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt,
                                                      int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = load_as_f32(src, sdt, i);
    if (ddt == KOCR_DTYPE_F32) static_cast<float*>(dst)[i] = v;
    else static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
  }
}

int launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, cudaStream_t stream) {
  if (n <= 0) return KOCR_OK;
  if (src_dtype < 0 || src_dtype > KOCR_DTYPE_F16 || (dst_dtype != KOCR_DTYPE_F32 && dst_dtype != KOCR_DTYPE_BF16))
    return fail(KOCR_ERR_INVALID, "convert: unsupported dtype");
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  convert_kernel<<<grid, 256, 0, stream>>>(src, src_dtype, dst, dst_dtype, n);
  KOCR_LAUNCH_CHECK("convert_kernel");
  return KOCR_OK;
}

// ---------------------------------------------------------------- weight row permutation (prepack; not on the hot path)
__global__ void __launch_bounds__(256) permute_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           const int32_t* __restrict__ perm, int64_t rows,
                                                           int64_t row_bytes, int64_t src_ld_bytes, int64_t dst_ld_bytes) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t s = perm[r];
    uint8_t* d = dst + r * dst_ld_bytes;
    if (s < 0) {
      for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = 0;
    } else {
      const uint8_t* p = src + s * src_ld_bytes;
      for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = p[i];
    }
  }
}

int launch_permute_rows(const void* src, void* dst, const int32_t* perm, int64_t rows, int64_t cols, int64_t src_ld,
                        int64_t dst_ld, int elt, cudaStream_t stream) {
  if (rows <= 0) return KOCR_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(rows, 148 * 8);
  permute_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), perm, rows,
                                                cols * elt, src_ld * elt, dst_ld * elt);
  KOCR_LAUNCH_CHECK("permute_rows_kernel");
  return KOCR_OK;
}

// ---------------------------------------------------------------- LayerNorm folding into the following Linear (prepack)
__global__ void __launch_bounds__(256) fold_norm_kernel(__nv_bfloat16* __restrict__ W, int64_t ldw, float* __restrict__ bias,
                                                        float* __restrict__ c1, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int64_t N, int64_t K) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  __nv_bfloat16* w = W + n * ldw;
  float acc_b = 0.f, acc_c = 0.f;
  for (int64_t k = lane; k < K; k += 32) {
    const float v = __bfloat162float(w[k]);
    if (beta) acc_b += beta[k] * v;
    const __nv_bfloat16 s = __float2bfloat16_rn(v * gamma[k]);
    acc_c += __bfloat162float(s);
    w[k] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
  }
  if (lane == 0) {
    bias[n] += acc_b;
    c1[n] = acc_c;
  }
}

int launch_fold_norm(void* W, int64_t ldw, float* bias, float* c1, const float* gamma, const float* beta, int64_t N, int64_t K,
                     cudaStream_t stream) {
  if (N <= 0) return KOCR_OK;
  fold_norm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(W), ldw, bias, c1, gamma, beta, N, K);
  KOCR_LAUNCH_CHECK("fold_norm_kernel");
  return KOCR_OK;
}

// ---------------------------------------------------------------- RoPE table
__global__ void rope_table_kernel(float2* cs, int max_pos, int n_freq, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * n_freq) return;
  const int p = i / n_freq, j = i % n_freq;
  // inv_freq[j] = 1 / theta^(2j / (2*n_freq)) in f32 (HF :278), angle = f32(p) * inv_freq (torch.outer, f32)
  const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)(2 * n_freq));
  const float ang = (float)p * inv_freq;
  float s, c;
  sincosf(ang, &s, &c);
  cs[i] = make_float2(c, s);
}

int launch_rope_table(float2* cs, int max_pos, int n_freq, float theta, cudaStream_t stream) {
  const int n = max_pos * n_freq;
  if (n <= 0) return KOCR_OK;
  rope_table_kernel<<<(n + 127) / 128, 128, 0, stream>>>(cs, max_pos, n_freq, theta);
  KOCR_LAUNCH_CHECK("rope_table_kernel");
  return KOCR_OK;
}

// ---------------------------------------------------------------- window gather of row groups
__global__ void __launch_bounds__(256) gather_groups_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                            const int32_t* __restrict__ index, int64_t n_groups,
                                                            int vec_per_group, bool inverse) {
  const int64_t total = n_groups * vec_per_group;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / vec_per_group;
    const int v = (int)(i % vec_per_group);
    const int64_t o = index[g];
    if (inverse) dst[o * vec_per_group + v] = src[i];   // dst[index[g]] = src[g]  (== src[argsort(index)])
    else dst[i] = src[o * vec_per_group + v];           // dst[g] = src[index[g]]
  }
}

int launch_gather_groups(const void* src, void* dst, const int32_t* index, int64_t n_groups, int group, int cols,
                         bool inverse, cudaStream_t stream) {
  if ((int64_t)group * cols % 8) return fail(KOCR_ERR_UNSUPPORTED, "gather: group bytes must be a multiple of 16");
  if (n_groups <= 0) return KOCR_OK;
  const int vpg = group * cols / 8;
  const unsigned grid = (unsigned)std::min<int64_t>((n_groups * vpg + 255) / 256, 148 * 16);
  gather_groups_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), index, n_groups,
                                                 vpg, inverse);
  KOCR_LAUNCH_CHECK("gather_groups_kernel");
  return KOCR_OK;
}

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_op_norm(KocrCtx* ctx, const void* x, const float* weight, const float* bias, void* y, int64_t rows,
                            int dim, float eps, int rms, void* stream) {
  if (!ctx || !x || !weight || !y) return fail(KOCR_ERR_INVALID, "kocr_op_norm: null argument");
  reset_launch_count();
  return launch_norm(x, dim, weight, bias, y, dim, rows, dim, eps, rms != 0, reinterpret_cast<cudaStream_t>(stream));
}
