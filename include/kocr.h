/*
 * kocr.h - C ABI of libkocr.so, the B200 (sm_100a) implementation of karanta-ocr's page-image hot path:
 *   page image -> smart_resize -> bicubic-AA resize -> normalise -> 14px patchify (pixel_values, image_grid_thw)
 *   -> Qwen2-VL / Qwen2.5-VL vision tower -> merged embeddings.
 *
 * The reference (The-African-Research-Collective/karanta-ocr) has no FFI of its own for this path: it calls
 * the third-party `transformers` Python API.  Each entry point below names the Python
 * call it stands in for, as `reference call site` -> `third-party function it reaches` (HF = transformers).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - every function returns KOCR_OK (0) or a negative KOCR_ERR_* code; kocr_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - device buffers (images, pixel_values, workspace, outputs, weight sources) are allocated and owned
 *     by the caller; the library owns only its prepacked weight copies and small planning tables.
 *   - all device work is enqueued on the caller's stream; no call synchronises the device except
 *     kocr_create / kocr_tower_set_weight / kocr_tower_finalize (one-time setup).
 *   - functions in the "host planning" group touch no GPU and work on a machine without one.
 */
#ifndef KOCR_H_
#define KOCR_H_

#include <stdint.h>

#if defined(__GNUC__)
#define KOCR_API __attribute__((visibility("default")))
#else
#define KOCR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define KOCR_OK 0
#define KOCR_ERR_INVALID (-1)     /* bad argument                                   -> Python ValueError  */
#define KOCR_ERR_ASPECT (-2)      /* smart_resize: aspect ratio > 200               -> Python ValueError  */
#define KOCR_ERR_CUDA (-3)        /* CUDA runtime / driver failure                  -> Python RuntimeError */
#define KOCR_ERR_UNSUPPORTED (-4) /* shape outside what the kernels were built for  -> Python RuntimeError */
#define KOCR_ERR_STATE (-5)       /* e.g. forward before all weights were set       -> Python RuntimeError */

/* resize arithmetic: which third-party fixed-point path to reproduce bit-for-bit */
#define KOCR_RESIZE_PIL 0  /* Pillow ImagingResample 8bpc (HF "pil" backend; transformers 4.53.3 slow processor) */
#define KOCR_RESIZE_ATEN 1 /* ATen uint8 _upsample_bicubic2d_aa (HF "torchvision" backend, transformers 5.x)     */

#define KOCR_LAYOUT_CHW 0  /* uint8 [3,H,W] planar      */
#define KOCR_LAYOUT_HWC 1  /* uint8 [H,W,3] interleaved */
#define KOCR_LAYOUT_GRAY 2 /* uint8 [H,W]; do_convert_rgb replicates it to 3 channels */

#define KOCR_DTYPE_F32 0
#define KOCR_DTYPE_BF16 1
#define KOCR_DTYPE_F16 2

#define KOCR_ARCH_QWEN2_VL 0
#define KOCR_ARCH_QWEN2_5_VL 1

typedef struct KocrCtx KocrCtx;
typedef struct KocrTower KocrTower;

KOCR_API const char* kocr_last_error(void);
KOCR_API const char* kocr_version(void);

/* ------------------------------------------------------------------ host planning (no GPU needed) */

/* HF models/qwen2_vl/image_processing_qwen2_vl.py:62-88 smart_resize (reached from
 * karanta/training/pipeline_steps.py:289-294).  KOCR_ERR_ASPECT when max/min > 200. */
KOCR_API int kocr_smart_resize(int height, int width, int factor, int64_t min_pixels, int64_t max_pixels,
                      int* out_height, int* out_width);

/* Taps per output sample of the antialiased bicubic filter: ceil(2*max(in/out,1))*2+1
 * (Pillow precompute_coeffs / ATen _compute_index_ranges_weights). */
KOCR_API int kocr_resample_ksize(int in_size, int out_size);

/* Fixed-point filter bank for one axis. bounds[out_size*2] = (first tap, tap count); coeffs[out_size*ksize];
 * *precision = right shift applied after accumulation.  mode = KOCR_RESIZE_*. */
KOCR_API int kocr_resample_coeffs(int in_size, int out_size, int mode, int32_t* bounds, int32_t* coeffs, int* precision);

/* lut[3*256]: the float32 a uint8 level of channel c maps to (HF image_processing_backends.py:291-331 for
 * KOCR_RESIZE_ATEN; HF image_transforms.py rescale+normalize for KOCR_RESIZE_PIL). */
KOCR_API int kocr_normalize_lut(int mode, float* lut);

/* HF Qwen2VLImageProcessor.get_number_of_image_patches (image_processing_qwen2_vl.py:234-261). */
KOCR_API int64_t kocr_num_patches(int height, int width, int patch, int merge, int64_t min_pixels, int64_t max_pixels);

/* HF modeling_qwen2_vl.py:725-748 rot_pos_emb: pos_hw[sumN*2] = (row, col) of each patch, merge-major order. */
KOCR_API int kocr_pos_ids(const int64_t* grid_thw, int n_images, int merge, int32_t* pos_hw);

/* HF modeling_qwen2_vl.py:772-780: cu[sum(t)+1], int32; *n_cu receives the entry count. */
KOCR_API int kocr_cu_seqlens(const int64_t* grid_thw, int n_images, int32_t* cu, int* n_cu);

/* HF modeling_qwen2_5_vl.py:411-451 get_window_index (+ unique_consecutive :476).
 * window_index[sumN/merge^2]; cu_window[<= sum windows + 1]; *n_cu_window receives the entry count. */
KOCR_API int kocr_window_index(const int64_t* grid_thw, int n_images, int window_size, int merge, int patch,
                      int32_t* window_index, int32_t* cu_window, int* n_cu_window);

/* ------------------------------------------------------------------ context */

/* One context per (device, caller thread).  Fails loudly (KOCR_ERR_CUDA / KOCR_ERR_UNSUPPORTED) when the
 * device is absent or is not compute capability 10.x: there is no CPU fallback. */
KOCR_API int kocr_create(int device, KocrCtx** out);
KOCR_API void kocr_destroy(KocrCtx* ctx);
/* Keep `n` SMs (0 <= n <= half the device) out of the tower's persistent GEMM grids and let kocr_png_decode use at most
 * that many: page decode for the NEXT batch then runs beside the tower of the current one (a side stream of higher
 * priority) instead of stealing SMs from one-CTA-per-SM kernels that were sized for the whole device.  0 (default) = off. */
KOCR_API int kocr_set_reserved_sms(KocrCtx* ctx, int n);

/* ------------------------------------------------------------------ image processor (device) */

typedef struct KocrImage {
  const uint8_t* data; /* DEVICE pointer */
  int32_t height;
  int32_t width;
  int32_t layout; /* KOCR_LAYOUT_* */
  int32_t reserved;
} KocrImage;

/* Stands in for Qwen2VLImageProcessor._preprocess (HF image_processing_qwen2_vl.py:148-232), reached from
 * karanta/training/pipeline_steps.py:289-294 and karanta/training/test_trained_model.py:82-87.
 *   pixel_values: DEVICE [sumN, C*tps*patch*patch] in out_dtype (F32 = drop-in; BF16 = feeds the tower directly)
 *   grid_thw_out: HOST  [n_images*3] int64 (t, h, w), input order preserved.
 * One fused kernel per call: resize (both passes) + normalise + patch-order write. */
KOCR_API int kocr_preprocess(KocrCtx* ctx, const KocrImage* images, int n_images, int64_t min_pixels, int64_t max_pixels,
                    int resize_mode, int out_dtype, void* pixel_values, int64_t pixel_values_capacity_rows,
                    int64_t* grid_thw_out, void* stream);

/* ------------------------------------------------------------------ vision tower (device) */

typedef struct KocrTowerConfig {
  int32_t arch; /* KOCR_ARCH_* */
  int32_t depth;
  int32_t embed_dim;  /* Qwen2VLVisionConfig.embed_dim / Qwen2_5_VLVisionConfig.hidden_size */
  int32_t num_heads;
  int32_t mlp_hidden; /* embed_dim*mlp_ratio / intermediate_size */
  int32_t out_hidden; /* Qwen2VLVisionConfig.hidden_size / Qwen2_5_VLVisionConfig.out_hidden_size */
  int32_t patch_size;
  int32_t temporal_patch_size;
  int32_t in_channels;
  int32_t spatial_merge_size;
  int32_t window_size;
  int32_t n_fullatt;
  int32_t fullatt_block_indexes[8];
} KocrTowerConfig;

/* Stands in for Qwen2VisionTransformerPretrainedModel.__init__ (HF modeling_qwen2_vl.py:687-722). */
KOCR_API int kocr_tower_create(KocrCtx* ctx, const KocrTowerConfig* cfg, KocrTower** out);
KOCR_API void kocr_tower_destroy(KocrTower* tower);

/* nn.Module.load_state_dict, one tensor at a time, under the HF key names ("blocks.3.attn.qkv.weight", ...).
 * `data` is a DEVICE pointer to a contiguous tensor of `dtype`; it is converted to bf16 and prepacked into
 * library-owned storage, so the caller may free it after the call returns. */
KOCR_API int kocr_tower_set_weight(KocrTower* tower, const char* name, const void* data, int dtype, const int64_t* shape,
                          int ndim);
/* KOCR_ERR_STATE (message lists the missing keys) unless every tensor of the architecture was set. */
KOCR_API int kocr_tower_finalize(KocrTower* tower);

/* Bytes of caller-allocated scratch kocr_tower_forward needs for this batch. */
KOCR_API int64_t kocr_tower_workspace_bytes(const KocrTower* tower, const int64_t* grid_thw, int n_images);

/* Stands in for visual(pixel_values, grid_thw=...) (HF modeling_qwen2_vl.py:757-795, reached from
 * karanta/training/ocr_training.py:86,670 via get_image_features :1118-1136; vLLM qwen2_vl.py:1376).
 *   pixel_values: DEVICE [sumN, patch_dim], pv_dtype F32 or BF16 (cast to bf16 like PatchEmbed.forward :306-309)
 *   grid_thw:     HOST [n_images*3] int64
 *   out:          DEVICE [sumN/merge^2, out_hidden] bf16 (pooler_output / the 4.53.3 return value)
 *   hidden_out:   optional DEVICE [sumN, embed_dim] bf16 (last_hidden_state), may be NULL */
KOCR_API int kocr_tower_forward(KocrTower* tower, const void* pixel_values, int pv_dtype, const int64_t* grid_thw,
                       int n_images, void* out, void* hidden_out, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* The tables kocr_tower_forward derives from grid_thw alone (patch positions for rot_pos_emb, HF modeling_qwen2_vl.py:
 * 725-752; cu_seqlens :772-780 as attention work lists; Qwen2.5-VL window index / per-row window bounds, HF
 * modeling_qwen2_5_vl.py:411-451) are planned once per distinct grid_thw and kept in HBM by the tower (up to 16 plans, least
 * recently used evicted): a forward over a grid seen before does no host planning and no table upload.  Counters for tests
 * and monitoring: forwards served from a cached plan / forwards that had to plan. */
KOCR_API int kocr_tower_plan_stats(const KocrTower* tower, int64_t* hits, int64_t* misses);

/* ------------------------------------------------------------------ page decode on the GPU */

/* Stands in for the host decode in front of the path: PIL.Image.open(BytesIO(base64.b64decode(...))) on the PNG pages the
 * reference passes around (karanta/data/utils.py:186-225 base64_to_grayscale and :228-251 prepare_image_and_text;
 * karanta/data/process_pdf_utils.py:50-75 render_pdf_to_base64png; karanta/pipeline.py:131-142), i.e. libpng + zlib. */
typedef struct KocrPngInfo {
  int32_t height, width;
  int32_t channels;      /* of the decoded page: 1 (gray, LAYOUT_GRAY) or 3 (RGB, LAYOUT_HWC); alpha is dropped like .convert() */
  int32_t src_channels;  /* samples per pixel in the file: 1, 2 (gray+alpha), 3, 4 (RGBA) */
  int32_t bit_depth, color_type, interlace;
  int32_t reserved;
  int64_t idat_bytes;    /* compressed payload */
  int64_t raw_bytes;     /* filtered scan lines = height * (1 + width * src_channels) */
} KocrPngInfo;

/* Host only: walk the container (signature, IHDR, IDAT..., IEND; chunk CRCs checked).  KOCR_ERR_INVALID for a damaged file,
 * KOCR_ERR_UNSUPPORTED for PNG flavours the kernels do not decode (palette, 16-bit, interlaced): decode those on the host. */
KOCR_API int kocr_png_info(const uint8_t* file, int64_t size, KocrPngInfo* info);
KOCR_API int64_t kocr_png_scratch_bytes(const KocrPngInfo* infos, int n);
/* files: HOST pointers to n whole PNG files; out_dev[i]: DEVICE buffer of height*width*channels bytes ([H][W][C], C = 1 or 3),
 * ready to be handed to kocr_preprocess as KOCR_LAYOUT_GRAY / KOCR_LAYOUT_HWC; scratch: DEVICE, kocr_png_scratch_bytes;
 * status_dev: DEVICE int32[n], 0 = decoded, else which stream check failed (read it after synchronising the stream).
 * Enqueues one H2D copy of the compressed bytes, inflate_kernel and unfilter_kernel on `stream`. */
KOCR_API int kocr_png_decode(KocrCtx* ctx, const uint8_t* const* files, const int64_t* sizes, int n, void* const* out_dev,
                             void* scratch, int64_t scratch_bytes, int32_t* status_dev, void* stream);

/* ------------------------------------------------------------------ LLM hand-off */

/* Host planning (no GPU): 3-D M-RoPE position ids of a text+image prompt.  Stands in for
 * Qwen2VLModel.get_rope_index (HF modeling_qwen2_vl.py:990-1092, get_vision_position_ids :934-988), reached from
 * karanta/training/ocr_training.py:86,670 (model(**batch)) and test_trained_model.py:91 (model.generate).
 *   input_ids [batch*seq_len] int64; attention_mask [batch*seq_len] int64 or NULL; image runs are the maximal runs of
 *   image_token_id among the unmasked tokens, consumed in order against image_grid_thw (t must be 1: still pages).
 *   position_ids [3*batch*seq_len] int64 (masked positions stay 0), deltas [batch] int64 (max position + 1 - length).
 * KOCR_ERR_INVALID when the number of image runs / their lengths disagree with image_grid_thw. */
KOCR_API int kocr_mrope_position_ids(const int64_t* input_ids, const int64_t* attention_mask, int batch, int seq_len,
                                     const int64_t* image_grid_thw, int n_images, int64_t image_token_id, int merge,
                                     int64_t* position_ids, int64_t* deltas);

/* Same, with the transformers line selectable.  KOCR_MROPE_TRANSFORMERS_5 is kocr_mrope_position_ids (pinned by
 * tests/golden/g6_llm_handoff.npz, minted from transformers 5.5.0).  KOCR_MROPE_TRANSFORMERS_4_5 follows the 4.5x
 * get_rope_index the reference pins (4.53.3, /root/reference/uv.lock:2168-2169): padded positions hold 1, deltas are taken
 * against the PADDED row length (what generate() adds to cache_position for left-padded batches), images are located by
 * vision_start_token_id (< 0: by runs of image tokens) and consume exactly their grid's token count, so adjacent images need
 * no separator.  4.5x is not installed in the build image: that mode is restated from its source and has no golden. */
#define KOCR_MROPE_TRANSFORMERS_5 0
#define KOCR_MROPE_TRANSFORMERS_4_5 1
KOCR_API int kocr_mrope_position_ids_v2(const int64_t* input_ids, const int64_t* attention_mask, int batch, int seq_len,
                                        const int64_t* image_grid_thw, int n_images, int64_t image_token_id,
                                        int64_t vision_start_token_id, int merge, int semantics, int64_t* position_ids,
                                        int64_t* deltas);

/* inputs_embeds.masked_scatter(input_ids == image_token_id, image_embeds) on the device, in place.  Stands in for
 * get_placeholder_mask + masked_scatter (HF modeling_qwen2_vl.py:1138-1177 and the forward's
 * `inputs_embeds.masked_scatter(image_mask, image_embeds)`).
 *   inputs_embeds: DEVICE bf16 [batch*seq_len, hidden]; image_embeds: DEVICE bf16 [n_rows, hidden]; input_ids: HOST.
 * KOCR_ERR_INVALID ("Image features and image tokens do not match, ...") when the placeholder count != n_rows. */
KOCR_API int kocr_scatter_image_embeds(KocrCtx* ctx, void* inputs_embeds, const void* image_embeds, int64_t n_rows, int hidden,
                                       const int64_t* input_ids, int batch, int seq_len, int64_t image_token_id, void* stream);

/* Number of kernels the last kocr_preprocess / kocr_tower_forward on this thread launched. */
KOCR_API int64_t kocr_last_launch_count(void);

/* Per-kernel-class device timing (CUDA events on the launching stream) for bench.py's roofline numbers.
 * kocr_profile_begin starts recording around every launch made through this ctx; kocr_profile_end waits for the
 * recorded events, fills ms_per_class / launches_per_class (max_classes entries) and returns the class count. */
KOCR_API int kocr_profile_begin(KocrCtx* ctx);
KOCR_API int kocr_profile_end(KocrCtx* ctx, int max_classes, double* ms_per_class, int64_t* launches_per_class);
KOCR_API const char* kocr_profile_class_name(int cls);

/* ------------------------------------------------------------------ single kernels (unit-level parity tests) */

#define KOCR_EPI_NONE 0          /* C = A.B^T                                   (PatchEmbed)          */
#define KOCR_EPI_BIAS 1          /* C = A.B^T + bias                                                   */
#define KOCR_EPI_BIAS_QUICKGELU 2 /* x*sigmoid(1.702x)                          (VisionMlp.fc1)       */
#define KOCR_EPI_BIAS_GELU 3     /* erf GELU                                    (PatchMerger.mlp[1])  */
#define KOCR_EPI_BIAS_RESIDUAL 4 /* C = residual + A.B^T + bias                 (attn.proj, fc2)      */
#define KOCR_EPI_BIAS_SWIGLU 5   /* interleaved gate/up columns -> silu(g)*u    (Qwen2_5_VLMLP)       */

/* C[M,N] (bf16, row pitch ldc elements) = epilogue(A[M,K] . B[N,K]^T); A, B bf16 row-major, K-contiguous. */
KOCR_API int kocr_op_gemm(KocrCtx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                 const void* residual, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue,
                 void* stream);

/* y = LayerNorm(x)*w + b (b may be NULL -> RMSNorm when rms != 0); rows of `dim` bf16. */
KOCR_API int kocr_op_norm(KocrCtx* ctx, const void* x, const float* weight, const float* bias, void* y, int64_t rows,
                 int dim, float eps, int rms, void* stream);

/* Varlen non-causal attention over packed [S, heads, 3, head_dim] q|k|v (the tower's internal QKV layout:
 * q already rotated and pre-scaled by head_dim^-0.5*log2(e)); out [S, heads*head_dim] bf16. */
KOCR_API int kocr_op_attention(KocrCtx* ctx, const void* qkv, void* out, const int32_t* cu_seqlens_host, int n_seqs,
                      int num_heads, int head_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KOCR_H_ */
