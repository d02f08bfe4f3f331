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
//   phase 3 LUT + patch-order gather smem -> smem; a tile's tokens are one contiguous span of pixel_values, which
//           leaves the SM as a single cp.async.bulk (TMA engine) copy per tile.
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
  long long src_bytes;    // size of the image buffer (staging never reads past it)
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
static constexpr int kMaxStageRows = 768;  // staged input rows per tile (3 planes x up to 256 rows)
static constexpr int kMaxDynSmem = 208 * 1024;  // dynamic shared memory opt-in; + ~10 KB static (LUT, job, row pointers) <= 227 KB
static constexpr int kRB = 14;           // rows of horizontal-pass accumulators held in registers per thread

// 1-D bulk async copy shared -> global (TMA engine, no tensor map): size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint8_t load_px(const PageJob& j, int c, int r, int x) {
  return __ldg(j.src + c * j.chan_stride + (long long)r * j.row_pitch + x * j.pix_stride);
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 3) preprocess_kernel(const PageJob* __restrict__ jobs, int n_jobs,
                                                              int n_tiles, const float* __restrict__ lut_g,
                                                              void* __restrict__ out, int max_mid_bytes, int coef_off, int out_off) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float lut[768];
  __shared__ PageJob job;
  __shared__ uintptr_t row_ptr[kMaxStageRows];
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

    // ---- phase 0: stage the input rows of this tile in smem with aligned 32-bit loads, all in flight at once (one
    // global-latency round trip per tile instead of one per filter tap). Rows are re-aligned on the way (funnel shift of
    // two aligned words): row r of segment s starts exactly at stage + (s*rows_in + r)*Lp.
    int xin0 = x0, ncols_in = tw;
    if (j.hb) {
      xin0 = j.hb[2 * x0];
      ncols_in = j.hb[2 * (x0 + tw - 1)] + j.hb[2 * (x0 + tw - 1) + 1] - xin0;
    }
    const int nseg = j.layout == KOCR_LAYOUT_CHW ? 3 : 1;     // planar: one segment per channel; interleaved / gray: one
    const int Lp = ((ncols_in * j.pix_stride + 6) & ~3) + 4;  // staged bytes per row, a multiple of 4
    const int wpr = Lp >> 2;
    uint8_t* stage = smem + out_off;  // aliases the output staging buffer (free until phase 3)
    const uint8_t* seg0 = j.src + (long long)r0 * j.row_pitch + (long long)xin0 * j.pix_stride;
    if (threadIdx.x == 0) bulk_store_wait_read();  // the previous tile's bulk copy has finished reading that buffer
    __syncthreads();
    {
      const uintptr_t img_end = (reinterpret_cast<uintptr_t>(j.src) + j.src_bytes + 3) & ~uintptr_t(3);
      uint32_t* sw = reinterpret_cast<uint32_t*>(stage);
      const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
      const int nrows = nseg * rows_in;
      constexpr int kWarps = kThreads / 32, kRowsInFlight = 6;
      // the 64-bit address of every staged row is worked out once per tile (one thread per row), not per lane and word
      for (int sr = threadIdx.x; sr < nrows; sr += kThreads) {
        const int sg = sr >= 2 * rows_in ? 2 : (sr >= rows_in ? 1 : 0);
        row_ptr[sr] = reinterpret_cast<uintptr_t>(seg0 + sg * j.chan_stride + (long long)(sr - sg * rows_in) * j.row_pitch);
      }
      __syncthreads();
      // warp per staged row, lanes over its words; kRowsInFlight rows are requested before the first is consumed, and
      // the upper word of each funnel shift comes from the neighbouring lane instead of a second load
      for (int wd0 = 0; wd0 < wpr; wd0 += 32) {
        const int wd = wd0 + lane;
        for (int sr0 = wrp; sr0 < nrows; sr0 += kWarps * kRowsInFlight) {
          uint32_t lo[kRowsInFlight], nx[kRowsInFlight];
          int sh[kRowsInFlight];
#pragma unroll
          for (int b = 0; b < kRowsInFlight; ++b) {
            const int sr = sr0 + b * kWarps;
            lo[b] = nx[b] = 0u;
            sh[b] = 0;
            if (sr < nrows) {
              const uintptr_t p = row_ptr[sr];
              const uintptr_t a = (p & ~uintptr_t(3)) + 4u * wd;
              sh[b] = (int)(p & 3) * 8;
              if (wd <= wpr && a < img_end) lo[b] = __ldg(reinterpret_cast<const uint32_t*>(a));
              if (lane == 31 && sh[b] != 0 && a + 4 < img_end) nx[b] = __ldg(reinterpret_cast<const uint32_t*>(a + 4));
            }
          }
#pragma unroll
          for (int b = 0; b < kRowsInFlight; ++b) {
            const int sr = sr0 + b * kWarps;
            const uint32_t up = __shfl_down_sync(0xffffffffu, lo[b], 1);
            const uint32_t hi = lane == 31 ? nx[b] : up;
            if (sr < nrows && wd < wpr) sw[sr * wpr + wd] = __funnelshift_r(lo[b], hi, sh[b]);  // byte k = byte k of the image row
          }
        }
      }
    }
    if (j.hb) {
      int32_t* kcoef = reinterpret_cast<int32_t*>(smem + coef_off);  // [hk][tw] transposed: conflict-free
      for (int i = threadIdx.x; i < j.hk * tw; i += kThreads) {
        const int t = i / tw, x = i % tw;
        kcoef[i] = j.hc[(size_t)(x0 + x) * j.hk + t];
      }
    }
    __syncthreads();

    // ---- phase 1: horizontal pass (or plain copy) stage -> mid. One thread per output column: its tap window and
    // coefficients are read once and reused for every row and channel (kRB rows of accumulators in registers).
    {
      const int groups = kThreads / tw;  // thread = (output column, row-block group): keeps all lanes busy at any tile width
      const int x = threadIdx.x % tw, grp = threadIdx.x / tw;
      const int chan_in_row = j.layout == KOCR_LAYOUT_HWC ? 1 : 0;  // byte step between channels inside a staged row
      const uint32_t stage_u32 = smem_u32(stage);
      if (grp < groups) {
        int xmin = x0 + x, cnt = 1;
        if (j.hb) { xmin = j.hb[2 * (x0 + x)]; cnt = j.hb[2 * (x0 + x) + 1]; }
        const int32_t* kcoef = reinterpret_cast<const int32_t*>(smem + coef_off);
        const int round0 = j.hb ? 1 << (j.hprec - 1) : 0;
        const int shr = j.hb ? j.hprec : 0;
        const int col_off = (xmin - xin0) * j.pix_stride;
        for (int rb = grp * kRB; rb < rows_in; rb += kRB * groups) {
          int acc[3][kRB];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < kRB; ++r) acc[c][r] = round0;
          const int nr = min(kRB, rows_in - rb);
          // rows past the strip's last input row read padding of the staging buffer; their sums are discarded below
          for (int t = 0; t < cnt; ++t) {
            const int kt = j.hb ? kcoef[t * tw + x] : 1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int sg = nseg == 3 ? c : 0;
              uint32_t a = stage_u32 + (uint32_t)((sg * rows_in + rb) * Lp + col_off + t * j.pix_stride + c * chan_in_row);
#pragma unroll
              for (int r = 0; r < kRB; ++r, a += Lp) acc[c][r] += (int)lds_u8(a) * kt;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < kRB; ++r)
              if (r < nr) mid[((size_t)c * rows_in + rb + r) * tw + x] = (uint8_t)min(max(acc[c][r] >> shr, 0), 255);
        }
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

    // ---- phase 3: normalise + patch order, staged in smem. One thread per (token, channel, patch row) run of 14 px:
    // 7 two-byte loads, 14 LUT reads, the run written for both temporal copies. The tile's tokens are ONE contiguous
    // span of pixel_values, so the whole staged tile leaves with a single bulk async copy (no LSU store traffic).
    const int cells = tw / kStrip;
    const int gw2 = j.out_w / kStrip;  // merge cells per row
    const long long n0 = j.token_base + ((long long)sy * gw2 + x0 / kStrip) * 4;
    constexpr int kElt = kBf16 ? 2 : 4;
    uint8_t* obuf = smem + out_off;  // the staged input it aliased was consumed in phase 1 (two barriers ago)
    const int runs = cells * 4 * 42;
    for (int run = threadIdx.x; run < runs; run += kThreads) {
      const int t = run / 42, rr = run % 42;
      const int c = rr / 14, py = rr % 14;
      const int cell = t >> 2, mh = (t >> 1) & 1, mw = t & 1;
      const uint16_t* p = reinterpret_cast<const uint16_t*>(res + ((size_t)c * kStrip + mh * 14 + py) * tw + cell * kStrip + mw * 14);
      uint8_t* d = obuf + ((size_t)t * kPatchDim + c * 392 + py * 14) * kElt;
      const float* l = lut + c * 256;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const uint32_t u = p[k];
        const float v0 = l[u & 255], v1 = l[u >> 8];
        if (kBf16) {
          const uint32_t pk = pack_bf16(v0, v1);
          reinterpret_cast<uint32_t*>(d)[k] = pk;
          reinterpret_cast<uint32_t*>(d + 196 * kElt)[k] = pk;  // tp = 1 copy
        } else {
          reinterpret_cast<float2*>(d)[k] = make_float2(v0, v1);
          reinterpret_cast<float2*>(d + 196 * kElt)[k] = make_float2(v0, v1);
        }
      }
    }
    fence_proxy_async();  // make the generic-proxy smem writes visible to the bulk copy engine
    __syncthreads();
    if (threadIdx.x == 0)
      bulk_store(reinterpret_cast<uint8_t*>(out) + n0 * kPatchDim * kElt, obuf, (uint32_t)(cells * 4 * kPatchDim * kElt));
  }
  if (threadIdx.x == 0) bulk_store_wait_all();
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
  if ((reinterpret_cast<uintptr_t>(pixel_values) & 15) != 0)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: pixel_values must be 16-byte aligned");

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
  const int kSmemBudget = 64 * 1024;
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
    j.src_bytes = (long long)im.height * im.width * (im.layout == KOCR_LAYOUT_GRAY ? 1 : 3);
    if (im.layout == KOCR_LAYOUT_CHW) { j.pix_stride = 1; j.row_pitch = im.width; j.chan_stride = (long long)im.height * im.width; }
    else if (im.layout == KOCR_LAYOUT_HWC) { j.pix_stride = 3; j.row_pitch = 3 * im.width; j.chan_stride = 1; }
    else { j.pix_stride = 1; j.row_pitch = im.width; j.chan_stride = 0; }
    j.out_h = oh; j.out_w = ow;
    j.token_base = tokens;
    const int rows_in = vt[i] ? vt[i]->max_rows : kStrip;
    if ((vt[i] ? 3 : 3) * rows_in > kMaxStageRows)
      return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: vertical downscale factor too large for one tile");
    int tw = 112;  // staged output tile: 16 tokens = 37.6 KB (bf16) / 75.3 KB (f32); keeps 3+ CTAs resident per SM
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
  StageGuard guard(ctx, slot, stream);  // released after preprocess_kernel (the tables' only reader) is enqueued
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
  if (rc) {
    guard.slot = -1;  // stage_commit released it
    return rc;
  }

  int max_coef = 0;
  for (int i = 0; i < n_images; ++i)
    if (ht[i]) max_coef = std::max(max_coef, ht[i]->ksize * jobs[i].tile_w * 4);
  int max_tw = 0;
  for (int i = 0; i < n_images; ++i) max_tw = std::max(max_tw, jobs[i].tile_w);
  const int coef_off = (max_mid + max_res + 15) & ~15;
  const int out_off = (coef_off + max_coef + 127) & ~127;
  int stage_bytes = 0;
  for (int i = 0; i < n_images; ++i) {
    const PageJob& j = jobs[i];
    const int rows_in = vt[i] ? vt[i]->max_rows : kStrip;
    const double scale = std::max(1.0, (double)j.in_w / j.out_w);
    const int ncols = (int)(j.tile_w * scale) + 2 * (ht[i] ? ht[i]->ksize : 0) + 8;  // generous bound on the tap span
    const int lp = ((ncols * j.pix_stride + 6) & ~3) + 4;
    stage_bytes = std::max(stage_bytes, ((j.layout == KOCR_LAYOUT_CHW ? 3 : 1) * rows_in + kRB) * lp);  // + padding rows
  }
  const int out_bytes = (max_tw / kStrip) * 4 * kPatchDim * (out_dtype == KOCR_DTYPE_BF16 ? 2 : 4);
  const int smem = out_off + std::max(out_bytes, stage_bytes);
  if (smem > kMaxDynSmem) return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: tile does not fit in shared memory");
  const int grid = std::min(tiles, ctx->num_sms * 8);
  const PageJob* d_jobs = reinterpret_cast<const PageJob*>(db + jobs_off);
  ProfScope ps(ctx, kProfPreprocess, stream);
  if (out_dtype == KOCR_DTYPE_BF16) {
    if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&preprocess_kernel<true>), kMaxDynSmem))) return rc;
    preprocess_kernel<true><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode],
                                                              pixel_values, max_mid, coef_off, out_off);
  } else {
    if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&preprocess_kernel<false>), kMaxDynSmem))) return rc;
    preprocess_kernel<false><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode],
                                                               pixel_values, max_mid, coef_off, out_off);
  }
  KOCR_LAUNCH_CHECK("preprocess_kernel");
  return KOCR_OK;
}
