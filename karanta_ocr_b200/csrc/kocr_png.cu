// PNG page decode on the GPU: the step in front of the hot path. The reference hands pages
// around as base64 PNG - 8-bit grayscale 'L' (karanta/data/utils.py:186-225, base64_to_grayscale) or RGB
// (karanta/data/process_pdf_utils.py:50-75, pdftoppm -png) - and decodes them with Pillow on the host
// (karanta/data/utils.py:228-251, karanta/pipeline.py:131-142). Here the file bytes go to the GPU as they are and the pixels
// land in HBM where kocr_preprocess reads them: no decoded image ever crosses PCIe and no host core spends 20 ms per page.
//
//   host   kocr_png_info: container walk (signature, IHDR, IDAT chunks with their CRCs, IEND)
//   kernel inflate_kernel: zlib / DEFLATE, one warp per page - lane 0 runs the Huffman state machine of
//          kocr_inflate_core.h, the 32 lanes refill its input window from HBM with 16-byte loads and copy LZ77 matches
//   kernel unfilter_kernel: scan-line reconstruction (None / Sub / Up / Average / Paeth), one warp per page as a 32-row
//          wavefront - lane k works one pixel behind lane k-1 and receives the pixel above it through a shuffle
// Byte work, HBM/latency bound by nature; many pages are in flight at once and the kernels are meant to run on a side stream
// under the previous batch's tower.
#include <string.h>

#include <algorithm>

#include <vector>

#include "kocr_common.cuh"
#include "kocr_inflate_core.h"

namespace kocr {

struct PngJob {
  const uint8_t* comp;   // zlib stream (IDAT payloads, concatenated), 16-byte aligned, padded with >= 16 readable bytes
  uint8_t* raw;          // filtered scan lines: height x (1 + width * bpp)
  uint8_t* out;          // pixels: [height][width][out_ch]
  long long comp_bytes, raw_bytes;
  int height, width, bpp, out_ch;
};

static constexpr int kPngWarps = 8;  // pages per CTA: few CTAs, so that a decode under a running tower ties up few SMs

__global__ void __launch_bounds__(kPngWarps * 32) inflate_kernel(const PngJob* __restrict__ jobs, int n_jobs, int32_t* __restrict__ status) {
  __shared__ inflate::Tables tabs[kPngWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int job = blockIdx.x * kPngWarps + warp; job < n_jobs; job += gridDim.x * kPngWarps) {  // the grid may be capped (reserved SMs)
  const PngJob j = jobs[job];
  inflate::Tables& t = tabs[warp];
  inflate::State s;
  s.in_size = j.comp_bytes;
  s.out_size = j.raw_bytes;
  // the compressed stream sits 16-byte aligned in the scratch buffer with slack behind it: whole uint4 loads are safe
  const uint4* src = reinterpret_cast<const uint4*>(j.comp);
  uint4* win = reinterpret_cast<uint4*>(t.window);
  for (int v = lane; v < inflate::kWindow / 16; v += 32)
    win[v] = (long long)v * 16 < j.comp_bytes + 16 ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
  __syncwarp();
  volatile uint8_t* out = j.raw;  // matches read bytes that other lanes (or lane 0, as literals) stored a moment ago
  for (;;) {
    int ev = 0;
    if (lane == 0) ev = inflate::run(s, t, j.raw);
    __syncwarp();  // orders lane 0's literal stores before the copies below
    ev = __shfl_sync(0xffffffffu, ev, 0);
    if (ev == inflate::kEvMatch) {
      const int len = __shfl_sync(0xffffffffu, s.match_len, 0), dist = __shfl_sync(0xffffffffu, s.match_dist, 0);
      const long long pos = __shfl_sync(0xffffffffu, (long long)s.out_pos, 0);
      if (dist >= 32 || dist >= len) {
        // source and destination chunks of 32 never overlap inside one step; later chunks may read what earlier ones wrote
        for (int j0 = 0; j0 < len; j0 += 32) {
          const int k = j0 + lane;
          if (k < len) out[pos + k] = out[pos - dist + k];
          if (dist < len) __syncwarp();
        }
      } else {
        // short period: the match repeats the last `dist` bytes, every source byte was written before the match began
        for (int k = lane; k < len; k += 32) out[pos + k] = out[pos - dist + k % dist];
      }
      __syncwarp();
      if (lane == 0) s.out_pos += len;
    } else if (ev == inflate::kEvRefill) {
      const long long base = __shfl_sync(0xffffffffu, (long long)s.win_base, 0);
      const long long from = base + inflate::kWindow;  // next kHalf bytes of input replace the half that was consumed
      for (int v = lane; v < inflate::kHalf / 16; v += 32) {
        const long long b = from + (long long)v * 16;
        win[(b & (inflate::kWindow - 1)) >> 4] = b < j.comp_bytes + 16 ? __ldg(src + (b >> 4)) : make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      if (lane == 0) s.win_base += inflate::kHalf;
    } else {
      break;
    }
  }
  if (lane == 0) status[job] = s.status;
  __syncwarp();
  }
}

// Scan-line reconstruction in place in `raw`, pixels (alpha dropped) to `out`. Lane k of the warp owns row r0 + k of a
// 32-row band and runs one pixel behind lane k - 1, so that the pixel above (b) arrives by shuffle from the lane that has
// just produced it and the one above-left (c) is last step's b; lane 0 reads the band's upper neighbour row from memory
// (the previous band finished it). A pixel is up to 4 bytes, carried packed in one register.
__global__ void __launch_bounds__(kPngWarps * 32) unfilter_kernel(const PngJob* __restrict__ jobs, int n_jobs, int32_t* __restrict__ status) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int job = blockIdx.x * kPngWarps + warp; job < n_jobs; job += gridDim.x * kPngWarps) {
  const PngJob j = jobs[job];
  if (status[job] != 0) continue;  // inflate failed: nothing to reconstruct
  const long long pitch = 1 + (long long)j.width * j.bpp;
  int bad = 0;
  for (int r0 = 0; r0 < j.height; r0 += 32) {
    const int r = r0 + lane;
    const bool live = r < j.height;
    uint8_t* row = j.raw + (long long)(live ? r : 0) * pitch + 1;
    const uint8_t* up_row = (lane == 0 && r0 > 0) ? j.raw + (long long)(r0 - 1) * pitch + 1 : nullptr;
    const int type = live ? row[-1] : 0;
    if (type > 4) bad = 1;
    uint8_t* orow = j.out + (long long)(live ? r : 0) * j.width * j.out_ch;
    uint32_t mine = 0, left = 0, above_left = 0;  // packed pixels: this lane's newest, its left neighbour, last step's `above`
    for (int step = 0; step < j.width + 31; ++step) {
      uint32_t above = __shfl_up_sync(0xffffffffu, mine, 1);  // lane k-1's pixel of the previous step = the pixel above ours
      const int x = step - lane;
      const bool on = live && x >= 0 && x < j.width;
      if (lane == 0) {
        above = 0;
        if (up_row && on)
          for (int c = 0; c < j.bpp; ++c) above |= (uint32_t)up_row[(long long)x * j.bpp + c] << (8 * c);
      }
      if (on) {
        if (x == 0) left = above_left = 0;
        uint32_t px = 0;
        for (int c = 0; c < j.bpp; ++c) {
          const int filt = row[(long long)x * j.bpp + c];
          const int v = pngfilter::recon(type, filt, (left >> (8 * c)) & 255, (above >> (8 * c)) & 255, (above_left >> (8 * c)) & 255);
          px |= (uint32_t)v << (8 * c);
          row[(long long)x * j.bpp + c] = (uint8_t)v;
          if (c < j.out_ch) orow[(long long)x * j.out_ch + c] = (uint8_t)v;
        }
        left = px;
        above_left = above;
        mine = px;
      }
    }
    __syncwarp();  // the band's last row is complete in memory before lane 0 of the next band reads it
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0 && bad) status[job] = inflate::kErrFilter;
  }
}

// ---------------------------------------------------------------------------------------------- host: container parsing
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) table[s][i] = table[0][table[s - 1][i] & 255] ^ (table[s - 1][i] >> 8);
    init = true;
  }
  crc = ~crc;
  while (n >= 8) {  // slicing by 8
    uint32_t a, b;
    memcpy(&a, p, 4);
    memcpy(&b, p + 4, 4);
    a ^= crc;
    crc = table[7][a & 255] ^ table[6][(a >> 8) & 255] ^ table[5][(a >> 16) & 255] ^ table[4][a >> 24] ^ table[3][b & 255] ^
          table[2][(b >> 8) & 255] ^ table[1][(b >> 16) & 255] ^ table[0][b >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = table[0][(crc ^ *p++) & 255] ^ (crc >> 8);
  return ~crc;
}

struct Segment { const uint8_t* p; size_t n; };

// Walks the chunks; fills `info` and, when `segs` is given, the IDAT payload segments. Checks the CRC of every chunk it uses.
static int parse_png(const uint8_t* f, int64_t size, KocrPngInfo* info, std::vector<Segment>* segs) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (!f || size < 8 + 25 || memcmp(f, sig, 8) != 0) return fail(KOCR_ERR_INVALID, "png: not a PNG file");
  int64_t pos = 8;
  bool have_ihdr = false, have_iend = false;
  int64_t idat = 0;
  memset(info, 0, sizeof *info);
  while (pos + 12 <= size) {
    const uint32_t len = be32(f + pos);
    const uint8_t* type = f + pos + 4;
    if ((int64_t)len > size - pos - 12) return fail(KOCR_ERR_INVALID, "png: truncated chunk");
    const uint8_t* data = f + pos + 8;
    const bool is_ihdr = !memcmp(type, "IHDR", 4), is_idat = !memcmp(type, "IDAT", 4), is_iend = !memcmp(type, "IEND", 4);
    if (is_ihdr || is_idat || is_iend) {
      if (crc32_update(0, type, 4 + (size_t)len) != be32(data + len)) return fail(KOCR_ERR_INVALID, "png: chunk CRC mismatch");
    }
    if (is_ihdr) {
      if (len != 13 || have_ihdr) return fail(KOCR_ERR_INVALID, "png: bad IHDR");
      have_ihdr = true;
      const uint32_t w = be32(data), h = be32(data + 4);
      const int depth = data[8], ctype = data[9], comp = data[10], filt = data[11], lace = data[12];
      if (w == 0 || h == 0 || w > 65535 || h > 65535 || comp != 0 || filt != 0) return fail(KOCR_ERR_INVALID, "png: bad IHDR fields");
      info->width = (int32_t)w;
      info->height = (int32_t)h;
      info->bit_depth = depth;
      info->color_type = ctype;
      info->interlace = lace;
      if (depth != 8 || lace != 0 || !(ctype == 0 || ctype == 2 || ctype == 4 || ctype == 6))
        return fail(KOCR_ERR_UNSUPPORTED, "png: only 8-bit non-interlaced gray / RGB (with or without alpha) is decoded on the GPU");
      info->src_channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : 4;
      info->channels = (ctype == 0 || ctype == 4) ? 1 : 3;
    } else if (is_idat) {
      if (!have_ihdr) return fail(KOCR_ERR_INVALID, "png: IDAT before IHDR");
      idat += len;
      if (segs && len) segs->push_back(Segment{data, (size_t)len});
    } else if (is_iend) {
      have_iend = true;
      break;
    }
    pos += 12 + (int64_t)len;
  }
  if (!have_ihdr || !have_iend || idat < 6) return fail(KOCR_ERR_INVALID, "png: missing IHDR / IDAT / IEND");
  info->idat_bytes = idat;
  info->raw_bytes = (int64_t)info->height * (1 + (int64_t)info->width * info->src_channels);
  return KOCR_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static size_t comp_slot(int64_t idat) { return align_up((size_t)idat + 32, 16); }  // >= 16 readable bytes behind the stream

}  // namespace kocr

using namespace kocr;

extern "C" {

int kocr_png_info(const uint8_t* file, int64_t size, KocrPngInfo* info) {
  if (!info) return fail(KOCR_ERR_INVALID, "kocr_png_info: null output");
  return parse_png(file, size, info, nullptr);
}

int64_t kocr_png_scratch_bytes(const KocrPngInfo* infos, int n) {
  if (!infos || n <= 0) return fail(KOCR_ERR_INVALID, "kocr_png_scratch_bytes: bad argument");
  size_t total = 256;
  for (int i = 0; i < n; ++i) total += comp_slot(infos[i].idat_bytes) + align_up((size_t)infos[i].raw_bytes + 16, 256);
  return (int64_t)total;
}

int kocr_png_decode(KocrCtx* ctx_, const uint8_t* const* files, const int64_t* sizes, int n, void* const* out_dev, void* scratch,
                    int64_t scratch_bytes, int32_t* status_dev, void* stream_) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  reset_launch_count();
  if (!ctx || !files || !sizes || n <= 0 || !out_dev || !scratch || !status_dev)
    return fail(KOCR_ERR_INVALID, "kocr_png_decode: null argument or empty batch");
  KOCR_CUDA_CHECK(cudaSetDevice(ctx->device));
  std::vector<KocrPngInfo> infos(n);
  std::vector<std::vector<Segment>> segs(n);
  size_t comp_total = 0;
  for (int i = 0; i < n; ++i) {
    int rc = parse_png(files[i], sizes[i], &infos[i], &segs[i]);
    if (rc) return rc;
    if (!out_dev[i]) return fail(KOCR_ERR_INVALID, "kocr_png_decode: null output pointer");
    comp_total += comp_slot(infos[i].idat_bytes);
  }
  if (kocr_png_scratch_bytes(infos.data(), n) > scratch_bytes) return fail(KOCR_ERR_INVALID, "kocr_png_decode: scratch too small");
  uint8_t* sc = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(scratch), 256));

  // the IDAT payloads, packed, go through a pinned buffer the context owns (grown on demand; reused once the previous
  // call's copy has left it)
  {
    std::lock_guard<std::mutex> lk(ctx->png_mu);
    if (ctx->png_pinned_ev) KOCR_CUDA_CHECK(cudaEventSynchronize(ctx->png_pinned_ev));
    else KOCR_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->png_pinned_ev, cudaEventDisableTiming));
    if (ctx->png_pinned_cap < comp_total) {
      if (ctx->png_pinned) cudaFreeHost(ctx->png_pinned);
      ctx->png_pinned = nullptr;
      ctx->png_pinned_cap = 0;
      const size_t cap = align_up(comp_total + comp_total / 2, 1 << 20);
      KOCR_CUDA_CHECK(cudaMallocHost(&ctx->png_pinned, cap));
      ctx->png_pinned_cap = cap;
    }
    uint8_t* hp = static_cast<uint8_t*>(ctx->png_pinned);
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      size_t o = off;
      for (const Segment& sg : segs[i]) {
        memcpy(hp + o, sg.p, sg.n);
        o += sg.n;
      }
      memset(hp + o, 0, comp_slot(infos[i].idat_bytes) - (size_t)infos[i].idat_bytes);
      off += comp_slot(infos[i].idat_bytes);
    }
    KOCR_CUDA_CHECK(cudaMemcpyAsync(sc, hp, comp_total, cudaMemcpyHostToDevice, stream));
    KOCR_CUDA_CHECK(cudaEventRecord(ctx->png_pinned_ev, stream));
  }
  std::vector<PngJob> jobs(n);
  size_t coff = 0, roff = align_up(comp_total, 256);
  for (int i = 0; i < n; ++i) {
    PngJob& j = jobs[i];
    j.comp = sc + coff;
    j.comp_bytes = infos[i].idat_bytes;
    j.raw = sc + roff;
    j.raw_bytes = infos[i].raw_bytes;
    j.out = static_cast<uint8_t*>(out_dev[i]);
    j.height = infos[i].height;
    j.width = infos[i].width;
    j.bpp = infos[i].src_channels;
    j.out_ch = infos[i].channels;
    coff += comp_slot(infos[i].idat_bytes);
    roff += align_up((size_t)infos[i].raw_bytes + 16, 256);
  }
  void* d_jobs;
  int slot = -1;
  int rc = ctx->stage(jobs.data(), jobs.size() * sizeof(PngJob), stream, &d_jobs, &slot);
  if (rc) return rc;
  StageGuard guard(ctx, slot, stream);
  ProfScope ps(ctx, kProfOther, stream);
  int grid = (n + kPngWarps - 1) / kPngWarps;
  if (ctx->reserved_sms > 0) grid = std::min(grid, ctx->reserved_sms);  // the tower's persistent grids leave exactly these SMs free
  inflate_kernel<<<grid, kPngWarps * 32, 0, stream>>>(static_cast<const PngJob*>(d_jobs), n, status_dev);
  KOCR_LAUNCH_CHECK("inflate_kernel");
  unfilter_kernel<<<grid, kPngWarps * 32, 0, stream>>>(static_cast<const PngJob*>(d_jobs), n, status_dev);
  KOCR_LAUNCH_CHECK("unfilter_kernel");
  return KOCR_OK;
}

}  // extern "C"
