// Fused page preprocess kernel: bicubic-AA resize (horizontal + vertical fixed-point passes, uint8
// intermediate) -> normalise (LUT) -> temporal duplicate -> write in Qwen2-VL patch order.
//
// Stands in for Qwen2VLImageProcessor._preprocess (HF models/qwen2_vl/image_processing_qwen2_vl.py:148-232)
// which the reference reaches from karanta/training/pipeline_steps.py:289-294.
//
// HBM-bound by design: each input byte is read ~once (plus the filter-support halo between strips),
// each output element is written exactly once with fully coalesced 8-byte (f32) / 4-byte (bf16) stores.
//   tile  = (page, strip of 28 output rows = one merge row, chunk of tile_w output columns)
//   phase 1 horizontal taps from global (L1-resident rows) -> smem mid[3][rows_in][tile_w]  (uint8)
//   phase 2 vertical taps from smem                      -> smem res[3][28][tile_w]        (uint8)
//   phase 3 LUT + patch-order gather from smem; a tile's tokens are one contiguous run of pixel_values.
#include <map>
#include <tuple>
#include <vector>

#include "kocr_common.cuh"

namespace kocr {

struct PageJob {
  const uint8_t* src;
  const int32_t* hb;  // horizontal bounds [out_w][2], or null when the width does not change
  const int32_t* hc;  // horizontal coeffs [out_w][hk]
  const int32_t* vb;  // vertical bounds [out_h][2], or null
  const int32_t* vc;  // vertical coeffs [out_h][vk]
  long long token_base;
  long long chan_stride;  // bytes between channels of one pixel (0 for gray: do_convert_rgb replicates)
  int row_pitch;          // bytes between rows
  int pix_stride;         // bytes between horizontally adjacent pixels
  int in_h, in_w, layout;
  int out_h, out_w;
  int hk, hprec, vk, vprec;
  int tile_w, tiles_x, tile_base;
};

static constexpr int kStrip = 28;       // output rows per tile = patch * merge
static constexpr int kPatchDim = 1176;  // 3 * 2 * 14 * 14
static constexpr int kThreads = 256;
static constexpr int kRB = 14;           // rows of horizontal-pass accumulators held in registers per thread

__device__ __forceinline__ uint8_t load_px(const PageJob& j, int c, int r, int x) {
  return __ldg(j.src + c * j.chan_stride + (long long)r * j.row_pitch + x * j.pix_stride);
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 3) preprocess_kernel(const PageJob* __restrict__ jobs, int n_jobs,
                                                              int n_tiles, const float* __restrict__ lut_g,
                                                              void* __restrict__ out, int max_mid_bytes, int coef_off) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float lut[768];
  __shared__ PageJob job;
  for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = lut_g[i];

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();  // previous tile's smem and `job` fully consumed
    if (threadIdx.x == 0) {
      int lo = 0, hi = n_jobs - 1;  // last job with tile_base <= tile
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].tile_base <= tile) lo = mid; else hi = mid - 1;
      }
      job = jobs[lo];
    }
    __syncthreads();
    const PageJob& j = job;
    const int local = tile - j.tile_base;
    const int sy = local / j.tiles_x, tx = local % j.tiles_x;
    const int y0 = sy * kStrip;
    const int x0 = tx * j.tile_w;
    const int tw = min(j.tile_w, j.out_w - x0);  // multiple of 28
    int r0 = y0, rows_in = kStrip;
    if (j.vb) {
      r0 = j.vb[2 * y0];
      rows_in = j.vb[2 * (y0 + kStrip - 1)] + j.vb[2 * (y0 + kStrip - 1) + 1] - r0;
    }
    uint8_t* mid = smem;                                   // [3][rows_in][tw]
    uint8_t* res = j.vb ? smem + max_mid_bytes : smem;     // [3][28][tw]

    // ---- phase 1: horizontal pass (or plain copy) into mid. One thread per output column: its tap window and
    // coefficients are read once, then reused for every row and channel (kRB rows of accumulators in registers);
    // consecutive threads read consecutive input bytes of the same row.
    if (j.hb) {
      int32_t* kcoef = reinterpret_cast<int32_t*>(smem + coef_off);  // [hk][tw] transposed: conflict-free
      for (int i = threadIdx.x; i < j.hk * tw; i += kThreads) {
        const int t = i / tw, x = i % tw;
        kcoef[i] = j.hc[(size_t)(x0 + x) * j.hk + t];
      }
      __syncthreads();
      const int x = threadIdx.x;
      if (x < tw) {
        const int xmin = j.hb[2 * (x0 + x)], cnt = j.hb[2 * (x0 + x) + 1];
        const int round0 = 1 << (j.hprec - 1);
        for (int rb = 0; rb < rows_in; rb += kRB) {
          int acc[3][kRB];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < kRB; ++r) acc[c][r] = round0;
          const uint8_t* col = j.src + (long long)(r0 + rb) * j.row_pitch + (long long)xmin * j.pix_stride;
          const int nr = min(kRB, rows_in - rb);
          for (int t = 0; t < cnt; ++t, col += j.pix_stride) {
            const int kt = kcoef[t * tw + x];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint8_t* pc = col + c * j.chan_stride;
#pragma unroll
              for (int r = 0; r < kRB; ++r, pc += j.row_pitch)
                if (r < nr) acc[c][r] += (int)__ldg(pc) * kt;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < kRB; ++r)
              if (rb + r < rows_in) mid[((size_t)c * rows_in + rb + r) * tw + x] = (uint8_t)min(max(acc[c][r] >> j.hprec, 0), 255);
        }
      }
    } else {
      const int n1 = 3 * rows_in * tw;
      for (int i = threadIdx.x; i < n1; i += kThreads) {
        const int x = i % tw;
        const int rc = i / tw;
        mid[i] = load_px(j, rc / rows_in, r0 + rc % rows_in, x0 + x);
      }
    }
    __syncthreads();

    // ---- phase 2: vertical pass mid -> res
    if (j.vb) {
      const int n2 = 3 * kStrip * tw;
      for (int i = threadIdx.x; i < n2; i += kThreads) {
        const int x = i % tw;
        const int yc = i / tw;
        const int y = yc % kStrip, c = yc / kStrip;
        const int ymin = j.vb[2 * (y0 + y)] - r0, cnt = j.vb[2 * (y0 + y) + 1];
        const int32_t* k = j.vc + (size_t)(y0 + y) * j.vk;
        const uint8_t* col = mid + ((size_t)c * rows_in + ymin) * tw + x;
        int acc = 1 << (j.vprec - 1);
        for (int t = 0; t < cnt; ++t) acc += (int)col[(size_t)t * tw] * k[t];
        res[i] = (uint8_t)min(max(acc >> j.vprec, 0), 255);
      }
      __syncthreads();
    }

    // ---- phase 3: normalise + patch-order write. unit = 2 px of one (token, c, py) run, written for tp = 0 and 1.
    const int cells = tw / kStrip;
    const int gw2 = j.out_w / kStrip;  // merge cells per row
    const long long n0 = j.token_base + ((long long)sy * gw2 + x0 / kStrip) * 4;
    const int units = cells * 4 * 294;
    for (int u = threadIdx.x; u < units; u += kThreads) {
      const int t = u / 294, r = u % 294;
      const int c = r / 98, r2 = r % 98;
      const int py = r2 / 7, jx = r2 % 7;
      const int cell = t >> 2, mh = (t >> 1) & 1, mw = t & 1;
      const int y = mh * 14 + py, x = cell * kStrip + mw * 14 + 2 * jx;
      const uint8_t* p = res + ((size_t)c * kStrip + y) * tw + x;
      const float v0 = lut[c * 256 + p[0]], v1 = lut[c * 256 + p[1]];
      const long long e = (n0 + t) * kPatchDim + c * 392 + py * 14 + 2 * jx;
      if (kBf16) {
        uint32_t* o = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + e);
        const uint32_t pk = pack_bf16(v0, v1);
        o[0] = pk;
        o[98] = pk;  // tp = 1 copy, 196 elements further
      } else {
        float2* o = reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + e);
        o[0] = make_float2(v0, v1);
        o[98] = make_float2(v0, v1);
      }
    }
  }
}

struct AxisTable {
  std::vector<int32_t> bounds, coeffs;
  int ksize = 0, prec = 0, max_rows = 0;  // max_rows: input rows a 28-row output strip can touch
};

static std::map<std::tuple<int, int, int>, AxisTable>& table_cache() {
  static thread_local std::map<std::tuple<int, int, int>, AxisTable> cache;
  return cache;
}

static int get_table(int in_size, int out_size, int mode, const AxisTable** out) {
  auto key = std::make_tuple(in_size, out_size, mode);
  auto& cache = table_cache();
  auto it = cache.find(key);
  if (it == cache.end()) {
    AxisTable t;
    t.ksize = resample_ksize(in_size, out_size);
    t.bounds.resize((size_t)out_size * 2);
    t.coeffs.resize((size_t)out_size * t.ksize);
    int rc = resample_coeffs(in_size, out_size, mode, t.bounds.data(), t.coeffs.data(), &t.prec);
    if (rc) return rc;
    for (int y0 = 0; y0 + kStrip <= out_size; y0 += kStrip) {
      int last = y0 + kStrip - 1;
      t.max_rows = std::max(t.max_rows, t.bounds[2 * last] + t.bounds[2 * last + 1] - t.bounds[2 * y0]);
    }
    it = cache.emplace(key, std::move(t)).first;
  }
  *out = &it->second;
  return KOCR_OK;
}

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_preprocess(KocrCtx* ctx_, const KocrImage* images, int n_images, int64_t min_pixels,
                               int64_t max_pixels, int resize_mode, int out_dtype, void* pixel_values,
                               int64_t capacity_rows, int64_t* grid_thw_out, void* stream_) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  reset_launch_count();
  if (!ctx || !images || n_images <= 0 || !pixel_values || !grid_thw_out)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: null argument or empty batch");
  if (resize_mode != KOCR_RESIZE_PIL && resize_mode != KOCR_RESIZE_ATEN)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: bad resize_mode");
  if (out_dtype != KOCR_DTYPE_F32 && out_dtype != KOCR_DTYPE_BF16)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: out_dtype must be F32 or BF16");
  if ((reinterpret_cast<uintptr_t>(pixel_values) & 7) != 0)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: pixel_values must be 8-byte aligned");

  // ---- plan on the host: sizes, filter banks (deduplicated), tiles
  if (table_cache().size() > 256) table_cache().clear();  // bound the per-thread cache (pointers below stay valid)
  std::vector<PageJob> jobs(n_images);
  struct Need { const AxisTable* t; size_t off_b, off_c; };
  std::map<const AxisTable*, Need> needs;
  size_t table_bytes = 0;
  auto want = [&](const AxisTable* t) {
    if (needs.count(t)) return;
    Need n{t, table_bytes, 0};
    table_bytes += t->bounds.size() * 4;
    n.off_c = table_bytes;
    table_bytes += t->coeffs.size() * 4;
    needs[t] = n;
  };
  std::vector<const AxisTable*> ht(n_images, nullptr), vt(n_images, nullptr);
  int64_t tokens = 0;
  int tiles = 0, max_mid = 0, max_res = 0;
  const int kSmemBudget = 96 * 1024;
  for (int i = 0; i < n_images; ++i) {
    const KocrImage& im = images[i];
    if (!im.data || im.layout < 0 || im.layout > 2) return fail(KOCR_ERR_INVALID, "kocr_preprocess: bad image");
    int oh, ow;
    int rc = smart_resize(im.height, im.width, kStrip, min_pixels, max_pixels, &oh, &ow);
    if (rc) return rc;
    if (im.width != ow) { rc = get_table(im.width, ow, resize_mode, &ht[i]); if (rc) return rc; want(ht[i]); }
    if (im.height != oh) { rc = get_table(im.height, oh, resize_mode, &vt[i]); if (rc) return rc; want(vt[i]); }
    PageJob& j = jobs[i];
    memset(&j, 0, sizeof j);
    j.src = im.data;
    j.in_h = im.height; j.in_w = im.width; j.layout = im.layout;
    if (im.layout == KOCR_LAYOUT_CHW) { j.pix_stride = 1; j.row_pitch = im.width; j.chan_stride = (long long)im.height * im.width; }
    else if (im.layout == KOCR_LAYOUT_HWC) { j.pix_stride = 3; j.row_pitch = 3 * im.width; j.chan_stride = 1; }
    else { j.pix_stride = 1; j.row_pitch = im.width; j.chan_stride = 0; }
    j.out_h = oh; j.out_w = ow;
    j.token_base = tokens;
    const int rows_in = vt[i] ? vt[i]->max_rows : kStrip;
    int tw = 252;
    while (tw > kStrip && 3 * tw * (rows_in + (vt[i] ? kStrip : 0)) > kSmemBudget) tw -= kStrip;
    if (3 * tw * (rows_in + (vt[i] ? kStrip : 0)) > 200 * 1024)
      return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: vertical downscale factor too large for one tile");
    tw = std::min(tw, ow);
    j.tile_w = tw;
    j.tiles_x = (ow + tw - 1) / tw;
    j.tile_base = tiles;
    tiles += j.tiles_x * (oh / kStrip);
    max_mid = std::max(max_mid, 3 * tw * rows_in);
    if (vt[i]) max_res = std::max(max_res, 3 * tw * kStrip);
    grid_thw_out[3 * i] = 1;
    grid_thw_out[3 * i + 1] = oh / 14;
    grid_thw_out[3 * i + 2] = ow / 14;
    tokens += (int64_t)(oh / 14) * (ow / 14);
  }
  if (tokens > capacity_rows)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: pixel_values buffer too small for this batch");
  max_mid = (max_mid + 15) & ~15;

  // ---- stage tables + jobs
  const size_t jobs_off = (table_bytes + 15) & ~size_t(15);
  const size_t total = jobs_off + sizeof(PageJob) * n_images;
  void* h;
  int slot;
  int rc = ctx->stage_begin(total, &h, &slot);
  if (rc) return rc;
  uint8_t* hb = static_cast<uint8_t*>(h);
  uint8_t* db = static_cast<uint8_t*>(ctx->d_slot[slot]);
  for (auto& kv : needs) {
    memcpy(hb + kv.second.off_b, kv.first->bounds.data(), kv.first->bounds.size() * 4);
    memcpy(hb + kv.second.off_c, kv.first->coeffs.data(), kv.first->coeffs.size() * 4);
  }
  for (int i = 0; i < n_images; ++i) {
    PageJob& j = jobs[i];
    if (ht[i]) {
      j.hb = reinterpret_cast<const int32_t*>(db + needs[ht[i]].off_b);
      j.hc = reinterpret_cast<const int32_t*>(db + needs[ht[i]].off_c);
      j.hk = ht[i]->ksize; j.hprec = ht[i]->prec;
    }
    if (vt[i]) {
      j.vb = reinterpret_cast<const int32_t*>(db + needs[vt[i]].off_b);
      j.vc = reinterpret_cast<const int32_t*>(db + needs[vt[i]].off_c);
      j.vk = vt[i]->ksize; j.vprec = vt[i]->prec;
    }
  }
  memcpy(hb + jobs_off, jobs.data(), sizeof(PageJob) * n_images);
  void* d;
  rc = ctx->stage_commit(slot, total, stream, &d);
  if (rc) return rc;

  int max_coef = 0;
  for (int i = 0; i < n_images; ++i)
    if (ht[i]) max_coef = std::max(max_coef, ht[i]->ksize * jobs[i].tile_w * 4);
  const int coef_off = (max_mid + max_res + 15) & ~15;
  const int smem = coef_off + max_coef;
  const int grid = std::min(tiles, ctx->num_sms * 8);
  const PageJob* d_jobs = reinterpret_cast<const PageJob*>(db + jobs_off);
  ProfScope ps(ctx, kProfPreprocess, stream);
  if (out_dtype == KOCR_DTYPE_BF16) {
    KOCR_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    preprocess_kernel<true><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode],
                                                              pixel_values, max_mid, coef_off);
  } else {
    KOCR_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    preprocess_kernel<false><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode],
                                                               pixel_values, max_mid, coef_off);
  }
  KOCR_LAUNCH_CHECK("preprocess_kernel");
  return KOCR_OK;
}
