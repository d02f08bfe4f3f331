// Host build of the DEFLATE state machine and the PNG scan-line reconstruction that the CUDA kernels run
// (karanta_ocr_b200/csrc/kocr_inflate_core.h), with the warp's two cooperative steps (window refill, match copy) done by a
// plain loop. Compiled by tests/test_png_host.py with g++ and checked against zlib / Pillow in the CPU suite.
#include <stdint.h>
#include <string.h>

#include "../../karanta_ocr_b200/csrc/kocr_inflate_core.h"

using namespace kocr;

extern "C" int kocr_test_inflate(const uint8_t* in, int64_t in_size, uint8_t* out, int64_t out_size, int64_t* produced) {
  static inflate::Tables t;
  inflate::State s;
  s.in_size = in_size;
  s.out_size = out_size;
  for (int64_t i = 0; i < inflate::kWindow; ++i) t.window[i] = i < in_size ? in[i] : 0;
  for (;;) {
    const int ev = inflate::run(s, t, out);
    if (ev == inflate::kEvMatch) {
      for (int j = 0; j < s.match_len; ++j) out[s.out_pos + j] = out[s.out_pos - s.match_dist + j];
      s.out_pos += s.match_len;
    } else if (ev == inflate::kEvRefill) {
      const int64_t from = s.win_base + inflate::kWindow;
      for (int64_t i = 0; i < inflate::kHalf; ++i) t.window[(from + i) & (inflate::kWindow - 1)] = from + i < in_size ? in[from + i] : 0;
      s.win_base += inflate::kHalf;
    } else {
      break;
    }
  }
  *produced = s.out_pos;
  return s.status;
}

// raw: h rows of (1 filter byte + w*bpp bytes), reconstructed in place; out: [h][w][out_ch] (alpha dropped when bpp = out_ch + 1)
extern "C" int kocr_test_unfilter(uint8_t* raw, int h, int w, int bpp, int out_ch, uint8_t* out) {
  const int64_t pitch = 1 + (int64_t)w * bpp;
  for (int r = 0; r < h; ++r) {
    uint8_t* row = raw + r * pitch + 1;
    const uint8_t* up = r ? raw + (r - 1) * pitch + 1 : nullptr;
    const int type = row[-1];
    if (type > 4) return inflate::kErrFilter;
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < bpp; ++c) {
        const int a = x ? row[(x - 1) * bpp + c] : 0, b = up ? up[x * bpp + c] : 0, cc = (x && up) ? up[(x - 1) * bpp + c] : 0;
        row[x * bpp + c] = (uint8_t)pngfilter::recon(type, row[x * bpp + c], a, b, cc);
        if (c < out_ch) out[((int64_t)r * w + x) * out_ch + c] = row[x * bpp + c];
      }
  }
  return 0;
}
