// Host-side planning of the page-image path: the integer / index work the reference reaches through
// transformers, restated in C++ behind the C ABI (include/kocr.h). No GPU is touched in this file except
// by the context helpers at the bottom.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "kocr_common.cuh"

namespace kocr {

static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

// ---- smart_resize: HF models/qwen2_vl/image_processing_qwen2_vl.py:62-88.
// Python's round() on a float is round-half-even == nearbyint() in the default FP environment.
int smart_resize(int h, int w, int factor, int64_t minp, int64_t maxp, int* oh, int* ow) {
  if (h <= 0 || w <= 0 || factor <= 0) return fail(KOCR_ERR_INVALID, "smart_resize: non-positive size");
  double ratio = (double)std::max(h, w) / (double)std::min(h, w);
  if (ratio > 200.0) {
    char buf[128];
    snprintf(buf, sizeof buf, "absolute aspect ratio must be smaller than 200, got %.17g", ratio);
    return fail(KOCR_ERR_ASPECT, buf);
  }
  int64_t hb = (int64_t)nearbyint((double)h / factor) * factor;
  int64_t wb = (int64_t)nearbyint((double)w / factor) * factor;
  if (hb * wb > maxp) {
    double beta = sqrt(((double)h * (double)w) / (double)maxp);
    hb = std::max<int64_t>(factor, (int64_t)floor((double)h / beta / factor) * factor);
    wb = std::max<int64_t>(factor, (int64_t)floor((double)w / beta / factor) * factor);
  } else if (hb * wb < minp) {
    double beta = sqrt((double)minp / ((double)h * (double)w));
    hb = (int64_t)ceil((double)h * beta / factor) * factor;
    wb = (int64_t)ceil((double)w * beta / factor) * factor;
  }
  *oh = (int)hb;
  *ow = (int)wb;
  return KOCR_OK;
}

// ---- antialiased bicubic filter bank (a = -0.5)
static double cubic_pil(double x) {  // Pillow Resample.c bicubic_filter
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
static double cubic_aten(double x) {  // ATen UpSample.h cubic_convolution1/2 via HelperInterpCubic::aa_filter
  const double a = -0.5;
  x = fabs(x);
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a;
  return 0.0;
}

int resample_ksize(int in_size, int out_size) {
  double scale = (double)in_size / (double)out_size;
  double support = 2.0 * std::max(scale, 1.0);
  return (int)ceil(support) * 2 + 1;
}

int resample_coeffs(int in_size, int out_size, int mode, int32_t* bounds, int32_t* coeffs, int* precision) {
  if (in_size <= 0 || out_size <= 0) return fail(KOCR_ERR_INVALID, "resample_coeffs: non-positive size");
  if (mode != KOCR_RESIZE_PIL && mode != KOCR_RESIZE_ATEN) return fail(KOCR_ERR_INVALID, "resample_coeffs: bad mode");
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = std::max(scale, 1.0);
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double inv = 1.0 / filterscale;
  std::vector<double> kk((size_t)out_size * ksize, 0.0);
  double wt_max = 0.0;
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (mode == KOCR_RESIZE_PIL) ? (0 + (xx + 0.5) * scale) : (scale * (xx + 0.5));
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    int cnt = xmax - xmin;
    if (mode == KOCR_RESIZE_ATEN) cnt = std::min(std::max(cnt, 0), ksize);
    double* k = &kk[(size_t)xx * ksize];
    double ww = 0.0;
    for (int x = 0; x < cnt; ++x) {
      double w = (mode == KOCR_RESIZE_PIL) ? cubic_pil((x + xmin - center + 0.5) * inv)
                                           : cubic_aten((x + xmin - center + 0.5) * inv);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < cnt; ++x) {
      if (ww != 0.0) k[x] /= ww;
      wt_max = std::max(wt_max, k[x]);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = cnt;
  }
  int prec;
  if (mode == KOCR_RESIZE_PIL) {
    prec = 32 - 8 - 2;
  } else {
    for (prec = 0; prec < 22; ++prec) {
      int next_value = (int)(0.5 + wt_max * (double)(1 << (prec + 1)));
      if (next_value >= (1 << 15)) break;
    }
  }
  const double mul = (double)(1 << prec);
  for (size_t i = 0; i < kk.size(); ++i) {
    double v = kk[i] * mul;
    int c = (v < 0) ? (int)(-0.5 + v) : (int)(0.5 + v);
    if (mode == KOCR_RESIZE_ATEN) c = (int)(int16_t)c;
    coeffs[i] = c;
  }
  *precision = prec;
  return KOCR_OK;
}

static const float kClipMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
static const float kClipStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};

int normalize_lut(int mode, float* lut) {
  if (mode != KOCR_RESIZE_PIL && mode != KOCR_RESIZE_ATEN) return fail(KOCR_ERR_INVALID, "normalize_lut: bad mode");
  for (int c = 0; c < 3; ++c) {
    for (int v = 0; v < 256; ++v) {
      volatile float x, m, s, d;
      if (mode == KOCR_RESIZE_ATEN) {
        // HF image_processing_backends.py:291-331: mean*255, std*255 as f32, then f32 sub and f32 div
        m = kClipMean[c] * 255.0f;
        s = kClipStd[c] * 255.0f;
        x = (float)v;
      } else {
        // HF image_transforms.py: rescale in f64 then cast to f32, then (x - mean) / std in f32
        m = kClipMean[c];
        s = kClipStd[c];
        x = (float)((double)v * (1.0 / 255.0));
      }
      d = x - m;
      lut[c * 256 + v] = d / s;
    }
  }
  return KOCR_OK;
}

// ---- index work
int pos_ids(const int64_t* grid, int n, int merge, int32_t* out) {
  size_t o = 0;
  for (int i = 0; i < n; ++i) {
    int64_t t = grid[3 * i], h = grid[3 * i + 1], w = grid[3 * i + 2];
    if (t <= 0 || h <= 0 || w <= 0 || h % merge || w % merge) return fail(KOCR_ERR_INVALID, "pos_ids: bad grid_thw");
    for (int64_t tt = 0; tt < t; ++tt)
      for (int64_t hb = 0; hb < h / merge; ++hb)
        for (int64_t wb = 0; wb < w / merge; ++wb)
          for (int mh = 0; mh < merge; ++mh)
            for (int mw = 0; mw < merge; ++mw) {
              out[o++] = (int32_t)(hb * merge + mh);
              out[o++] = (int32_t)(wb * merge + mw);
            }
  }
  return KOCR_OK;
}

int cu_seqlens(const int64_t* grid, int n, int32_t* cu, int* n_cu) {
  int k = 0;
  int64_t acc = 0;
  cu[k++] = 0;
  for (int i = 0; i < n; ++i) {
    int64_t t = grid[3 * i], hw = grid[3 * i + 1] * grid[3 * i + 2];
    if (t <= 0 || hw <= 0) return fail(KOCR_ERR_INVALID, "cu_seqlens: bad grid_thw");
    for (int64_t tt = 0; tt < t; ++tt) {
      acc += hw;
      if (acc > INT32_MAX) return fail(KOCR_ERR_UNSUPPORTED, "cu_seqlens: more than 2^31 patches");
      cu[k++] = (int32_t)acc;
    }
  }
  *n_cu = k;
  return KOCR_OK;
}

int window_index(const int64_t* grid, int n, int window_size, int merge, int patch, int32_t* widx, int32_t* cuw,
                 int* n_cuw) {
  const int win = window_size / merge / patch;
  if (win <= 0) return fail(KOCR_ERR_INVALID, "window_index: window smaller than one merged cell");
  size_t o = 0;
  int k = 0;
  int64_t base = 0, last = 0;
  cuw[k++] = 0;
  for (int i = 0; i < n; ++i) {
    int64_t t = grid[3 * i], lh = grid[3 * i + 1] / merge, lw = grid[3 * i + 2] / merge;
    int64_t pad_h = win - lh % win, pad_w = win - lw % win;  // a full extra window when divisible (HF :424-425)
    int64_t nwh = (lh + pad_h) / win, nww = (lw + pad_w) / win;
    for (int64_t tt = 0; tt < t; ++tt)
      for (int64_t wy = 0; wy < nwh; ++wy)
        for (int64_t wx = 0; wx < nww; ++wx) {
          int64_t cnt = 0;
          for (int y = 0; y < win; ++y)
            for (int x = 0; x < win; ++x) {
              int64_t yy = wy * win + y, xx = wx * win + x;
              if (yy < lh && xx < lw) {
                widx[o++] = (int32_t)(base + (tt * lh + yy) * lw + xx);
                ++cnt;
              }
            }
          int64_t v = last + cnt * merge * merge;
          if (v != last) cuw[k++] = (int32_t)v;  // unique_consecutive (HF :476)
          last = v;
        }
    base += t * lh * lw;
  }
  *n_cuw = k;
  return KOCR_OK;
}

// ---- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static std::atomic<EncodeTiledFn> g_encode{nullptr};

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo) {
  EncodeTiledFn fn = g_encode.load();
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      return fail(KOCR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = (EncodeTiledFn)p;
    g_encode.store(fn);
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] stride0 %llu box [%u,%u,%u]",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return fail(KOCR_ERR_CUDA, buf);
  }
  return KOCR_OK;
}

// ---- context
int Ctx::stage_begin(size_t bytes, void** h_out, int* slot) {
  if (bytes > kSlotBytes) return fail(KOCR_ERR_UNSUPPORTED, "planning table larger than the staging slot");
  int s = -1;
  {
    std::lock_guard<std::mutex> lk(slot_mu);
    for (int k = 0; k < kSlots; ++k) {  // next slot in ring order that is not still waiting for its consumers to be enqueued
      const int c = (next_slot + k) % kSlots;
      if (!slot_open[c]) {
        s = c;
        break;
      }
    }
    if (s < 0) return fail(KOCR_ERR_STATE, "staging ring exhausted: more than 8 calls staged and not yet released");
    next_slot = (s + 1) % kSlots;
    slot_open[s] = true;
  }
  cudaError_t e = cudaEventSynchronize(slot_ev[s]);  // recorded after the slot's last consumer: normally long complete
  if (e != cudaSuccess) {
    std::lock_guard<std::mutex> lk(slot_mu);
    slot_open[s] = false;
    return fail(KOCR_ERR_CUDA, cudaGetErrorString(e));
  }
  *h_out = h_slot[s];
  *slot = s;
  return KOCR_OK;
}
int Ctx::stage_commit(int s, size_t bytes, cudaStream_t stream, void** d_out) {
  cudaError_t e = cudaMemcpyAsync(d_slot[s], h_slot[s], bytes, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) {
    stage_release(s, stream);
    return fail(KOCR_ERR_CUDA, cudaGetErrorString(e));
  }
  *d_out = d_slot[s];
  return KOCR_OK;
}
void Ctx::stage_release(int s, cudaStream_t stream) {
  if (s < 0 || s >= kSlots) return;
  cudaEventRecord(slot_ev[s], stream);
  std::lock_guard<std::mutex> lk(slot_mu);
  slot_open[s] = false;
}
int Ctx::stage(const void* src, size_t bytes, cudaStream_t stream, void** d_out, int* slot_out) {
  void* h;
  int s;
  int rc = stage_begin(bytes, &h, &s);
  if (rc) return rc;
  memcpy(h, src, bytes);
  rc = stage_commit(s, bytes, stream, d_out);
  if (rc) return rc;
  if (slot_out) *slot_out = s;       // the caller releases it after enqueueing the consumers (StageGuard)
  else stage_release(s, stream);     // no consumer beyond the copy itself
  return KOCR_OK;
}
int Ctx::opt_in_smem(const void* func, int bytes) {
  std::lock_guard<std::mutex> lk(attr_mu);
  if (smem_attr_done.count(func)) return KOCR_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();  // not sticky: do not let the next launch check report it
    return fail(KOCR_ERR_CUDA, std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize, ") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  }
  smem_attr_done.insert(func);
  return KOCR_OK;
}

cudaEvent_t Ctx::prof_event() {
  if (!prof_pool.empty()) {
    cudaEvent_t e = prof_pool.back();
    prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

static const char* kProfNames[kProfClasses] = {"preprocess", "gemm_patch_embed", "norm", "gemm_qkv_rope", "attention",
                                               "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_merger", "other", "attention_windowed"};

}  // namespace kocr

using namespace kocr;

extern "C" {

int kocr_profile_begin(KocrCtx* ctx_) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx_);
  if (!c) return fail(KOCR_ERR_INVALID, "kocr_profile_begin: null ctx");
  for (auto& r : c->prof_recs) {
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  c->prof_on = true;
  return KOCR_OK;
}

int kocr_profile_end(KocrCtx* ctx_, int max_classes, double* ms_per_class, int64_t* launches_per_class) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx_);
  if (!c || !ms_per_class || !launches_per_class) return fail(KOCR_ERR_INVALID, "kocr_profile_end: null argument");
  c->prof_on = false;
  for (int i = 0; i < max_classes; ++i) {
    ms_per_class[i] = 0;
    launches_per_class[i] = 0;
  }
  for (auto& r : c->prof_recs) {
    KOCR_CUDA_CHECK(cudaEventSynchronize(r.b));
    float ms = 0;
    KOCR_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.cls < max_classes) {
      ms_per_class[r.cls] += ms;
      launches_per_class[r.cls] += 1;
    }
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  return kProfClasses;
}

const char* kocr_profile_class_name(int cls) { return (cls >= 0 && cls < kProfClasses) ? kProfNames[cls] : ""; }

const char* kocr_last_error(void) { return g_err.c_str(); }
const char* kocr_version(void) { return "kocr 0.1 (sm_100a)"; }
int64_t kocr_last_launch_count(void) { return g_launches; }

int kocr_smart_resize(int height, int width, int factor, int64_t min_pixels, int64_t max_pixels, int* out_height,
                      int* out_width) {
  if (!out_height || !out_width) return fail(KOCR_ERR_INVALID, "null output");
  return smart_resize(height, width, factor, min_pixels, max_pixels, out_height, out_width);
}
int kocr_resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return fail(KOCR_ERR_INVALID, "resample_ksize: non-positive size");
  return resample_ksize(in_size, out_size);
}
int kocr_resample_coeffs(int in_size, int out_size, int mode, int32_t* bounds, int32_t* coeffs, int* precision) {
  if (!bounds || !coeffs || !precision) return fail(KOCR_ERR_INVALID, "null output");
  return resample_coeffs(in_size, out_size, mode, bounds, coeffs, precision);
}
int kocr_normalize_lut(int mode, float* lut) {
  if (!lut) return fail(KOCR_ERR_INVALID, "null output");
  return normalize_lut(mode, lut);
}
int64_t kocr_num_patches(int height, int width, int patch, int merge, int64_t min_pixels, int64_t max_pixels) {
  int oh, ow;
  int rc = smart_resize(height, width, patch * merge, min_pixels, max_pixels, &oh, &ow);
  if (rc) return rc;
  return (int64_t)(oh / patch) * (ow / patch);
}
int kocr_pos_ids(const int64_t* grid_thw, int n_images, int merge, int32_t* pos_hw) {
  if (!grid_thw || !pos_hw || n_images < 0 || merge <= 0) return fail(KOCR_ERR_INVALID, "pos_ids: bad argument");
  return pos_ids(grid_thw, n_images, merge, pos_hw);
}
int kocr_cu_seqlens(const int64_t* grid_thw, int n_images, int32_t* cu, int* n_cu) {
  if (!grid_thw || !cu || !n_cu || n_images < 0) return fail(KOCR_ERR_INVALID, "cu_seqlens: bad argument");
  return cu_seqlens(grid_thw, n_images, cu, n_cu);
}
int kocr_window_index(const int64_t* grid_thw, int n_images, int window_size, int merge, int patch,
                      int32_t* window_index_out, int32_t* cu_window, int* n_cu_window) {
  if (!grid_thw || !window_index_out || !cu_window || !n_cu_window || n_images < 0 || merge <= 0 || patch <= 0)
    return fail(KOCR_ERR_INVALID, "window_index: bad argument");
  return window_index(grid_thw, n_images, window_size, merge, patch, window_index_out, cu_window, n_cu_window);
}

int kocr_create(int device, KocrCtx** out) {
  if (!out) return fail(KOCR_ERR_INVALID, "null output");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(KOCR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                   " (libkocr has no CPU fallback)");
  if (device < 0 || device >= count) return fail(KOCR_ERR_INVALID, "device index out of range");
  KOCR_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  KOCR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(KOCR_ERR_UNSUPPORTED, std::string("libkocr is built for sm_100a only; device is ") + prop.name);
  Ctx* c = new Ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  for (int i = 0; i < Ctx::kSlots; ++i) {
    KOCR_CUDA_CHECK(cudaMallocHost(&c->h_slot[i], Ctx::kSlotBytes));
    KOCR_CUDA_CHECK(cudaMalloc(&c->d_slot[i], Ctx::kSlotBytes));
    KOCR_CUDA_CHECK(cudaEventCreateWithFlags(&c->slot_ev[i], cudaEventDisableTiming));
  }
  for (int mode = 0; mode < 2; ++mode) {
    float lut[768];
    normalize_lut(mode, lut);
    KOCR_CUDA_CHECK(cudaMalloc(&c->d_lut[mode], sizeof lut));
    KOCR_CUDA_CHECK(cudaMemcpy(c->d_lut[mode], lut, sizeof lut, cudaMemcpyHostToDevice));
  }
  *out = reinterpret_cast<KocrCtx*>(c);
  return KOCR_OK;
}

int kocr_set_reserved_sms(KocrCtx* ctx_, int n) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || n < 0 || n > ctx->num_sms / 2) return fail(KOCR_ERR_INVALID, "kocr_set_reserved_sms: n must be in [0, SMs / 2]");
  ctx->reserved_sms = n;
  return KOCR_OK;
}

void kocr_destroy(KocrCtx* ctx) {
  if (!ctx) return;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaSetDevice(c->device);
  for (int i = 0; i < Ctx::kSlots; ++i) {
    if (c->h_slot[i]) cudaFreeHost(c->h_slot[i]);
    if (c->d_slot[i]) cudaFree(c->d_slot[i]);
    if (c->slot_ev[i]) cudaEventDestroy(c->slot_ev[i]);
  }
  for (int mode = 0; mode < 2; ++mode)
    if (c->d_lut[mode]) cudaFree(c->d_lut[mode]);
  if (c->png_pinned) cudaFreeHost(c->png_pinned);
  if (c->png_pinned_ev) cudaEventDestroy(c->png_pinned_ev);
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  delete c;
}

}  // extern "C"
