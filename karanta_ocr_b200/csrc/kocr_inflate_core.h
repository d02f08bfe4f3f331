// DEFLATE (RFC 1951) decoder state machine and PNG scan-line reconstruction (PNG spec section 9), written once for host
// and device: the CUDA kernels in kocr_png.cu run it with lane 0 of a warp as the decoder and the whole warp as the copy
// engine; tests/native/inflate_host.cpp runs the very same code on the CPU against zlib (`-m "not gpu"` suite).
//
// Stands in for the page decode the reference does on the host before the hot path: PIL.Image.open(BytesIO(base64...)) in
// karanta/data/utils.py:186-225 (base64_to_grayscale), :228-251 (prepare_image_and_text) and
// karanta/data/process_pdf_utils.py:50-75 (pdftoppm -png), i.e. libpng + zlib inside Pillow.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KOCR_HD __host__ __device__ __forceinline__
#else
#define KOCR_HD inline
#endif

namespace kocr {
namespace inflate {

static constexpr int kWindow = 2048;       // compressed-input window (bytes), refilled in halves
static constexpr int kHalf = kWindow / 2;
static constexpr int kLitFastBits = 10;    // primary lookup width of the literal/length code
static constexpr int kDistFastBits = 8;    // ... of the distance code
static constexpr int kMaxBits = 15;

enum Status : int32_t {
  kOk = 0,
  kErrHeader = 1,        // bad zlib header / block type
  kErrCode = 2,          // invalid or incomplete Huffman code / symbol
  kErrDistance = 3,      // match reaches before the start of the output
  kErrOverflow = 4,      // more output than the image holds
  kErrTruncated = 5,     // input ended early / output short
  kErrFilter = 6,        // scan line with an unknown filter type
};

// What lane 0 hands to the warp when it stops decoding.
enum Event : int32_t { kEvMatch = 0, kEvRefill = 1, kEvDone = 2, kEvError = 3, kEvFlush = 4 };

struct Tables {            // per stream, in shared memory on the device
  uint16_t lit_fast[1 << kLitFastBits];   // (symbol << 4) | code length, 0 = longer than the fast width (or unused)
  uint16_t dist_fast[1 << kDistFastBits];
  uint16_t lit_count[kMaxBits + 1], dist_count[kMaxBits + 1];  // canonical code: codes per length
  uint16_t lit_symbol[288], dist_symbol[32];                   // symbols ordered by code
  uint8_t lengths[320];                                        // scratch while a dynamic header is read
  alignas(16) uint8_t window[kWindow];                         // compressed bytes [base, base + kWindow), read as aligned 32-bit words
};

struct State {
  uint64_t bitbuf = 0;
  int bitcnt = 0;
  int64_t in_pos = 0;       // next input byte to load into the bit buffer
  int64_t in_size = 0;
  int64_t win_base = 0;     // window holds input bytes [win_base, win_base + kWindow)
  int64_t out_pos = 0;
  int64_t out_size = 0;
  // `out` may be a ring (the device keeps the last 32 KB of output in shared memory and streams it to HBM behind the decoder):
  // bytes go to out[out_pos & out_mask], and run() yields kEvFlush once out_pos reaches flush_at. A plain buffer: mask -1.
  int64_t out_mask = -1;
  int64_t flush_at = INT64_MAX;
  int phase = 0;            // 0 = zlib header, 1 = block header, 2 = stored block, 3 = Huffman block, 4 = finished
  int last_block = 0;
  int stored_left = 0;
  int match_len = 0, match_dist = 0;
  int status = kOk;
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

#if defined(__CUDACC__)
__device__ static const uint16_t kLenBase_d[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ static const uint8_t kLenExtra_d[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ static const uint16_t kDistBase_d[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ static const uint8_t kDistExtra_d[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ static const uint8_t kClOrder_d[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
#endif
#if defined(__CUDA_ARCH__)
#define KOCR_TAB(name) name##_d
#else
#define KOCR_TAB(name) name
#endif

// ---- bit reader over the window (LSB first, RFC 1951 section 3.1.1). The buffer is topped up four bytes at a time with one
// aligned word read (in_pos stays a multiple of 4), which leaves at least 32 valid bits behind every call - enough for any one
// step below (a literal/length code with its extra bits is at most 20 bits, a distance code with extra bits 28, a stored
// block's LEN/NLEN 32). Bytes past the end of the input read as zero (the window is zero-filled there); running past the end
// is detected by the caller through in_pos.
KOCR_HD void refill_bits(State& s, const Tables& t) {
  if (s.bitcnt <= 32) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&t.window[s.in_pos & (kWindow - 1)]);
    s.bitbuf |= (uint64_t)w << s.bitcnt;
    s.bitcnt += 32;
    s.in_pos += 4;
  }
}
KOCR_HD uint32_t peek(const State& s, int n) { return (uint32_t)(s.bitbuf & ((1ull << n) - 1)); }
KOCR_HD void drop(State& s, int n) {
  s.bitbuf >>= n;
  s.bitcnt -= n;
}
KOCR_HD uint32_t take(State& s, int n) {
  const uint32_t v = peek(s, n);
  drop(s, n);
  return v;
}
// input bytes really consumed so far (the bit buffer holds whole bytes ahead of the read point)
KOCR_HD int64_t consumed(const State& s) { return s.in_pos - (s.bitcnt >> 3); }

KOCR_HD uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

// Canonical Huffman code from code lengths (RFC 1951 section 3.2.2): count / symbol arrays for the bit-serial decoder and a
// direct lookup for codes up to fast_bits. Returns false for an over-subscribed code (incomplete codes are allowed only
// where deflate allows them: a single distance code).
KOCR_HD bool build_code(const uint8_t* lengths, int n, uint16_t* count, uint16_t* symbol, uint16_t* fast, int fast_bits) {
  for (int l = 0; l <= kMaxBits; ++l) count[l] = 0;
  for (int i = 0; i < n; ++i) ++count[lengths[i]];
  for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
  if (count[0] == n) return true;  // no codes at all: legal for the distance code of a literal-only block
  int left = 1;
  for (int l = 1; l <= kMaxBits; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return false;
  }
  uint16_t offs[kMaxBits + 2];
  offs[1] = 0;
  for (int l = 1; l <= kMaxBits; ++l) offs[l + 1] = offs[l] + count[l];
  for (int i = 0; i < n; ++i)
    if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
  // direct table: walk the symbols in canonical order, code increments by one within a length
  uint32_t code = 0;
  int idx = 0;
  for (int l = 1; l <= kMaxBits; ++l) {
    for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
      if (l <= fast_bits) {
        const uint32_t rev = reverse_bits(code, l);
        const uint16_t e = (uint16_t)((symbol[idx] << 4) | l);
        for (uint32_t x = rev; x < (1u << fast_bits); x += (1u << l)) fast[x] = e;
      }
    }
    code <<= 1;
  }
  return true;
}

// One symbol. Direct lookup first; codes longer than the fast width are walked bit by bit (canonical decode).
KOCR_HD int decode_symbol(State& s, const uint16_t* fast, int fast_bits, const uint16_t* count, const uint16_t* symbol) {
  const uint16_t e = fast[peek(s, fast_bits)];
  if (e) {
    drop(s, e & 15);
    return e >> 4;
  }
  int code = 0, first = 0, index = 0;
  for (int l = 1; l <= kMaxBits; ++l) {
    code |= (int)((s.bitbuf >> (l - 1)) & 1u);
    const int c = count[l];
    if (code - c < first) {
      drop(s, l);
      return symbol[index + (code - first)];
    }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

KOCR_HD bool fixed_tables(Tables& t) {
  for (int i = 0; i < 144; ++i) t.lengths[i] = 8;
  for (int i = 144; i < 256; ++i) t.lengths[i] = 9;
  for (int i = 256; i < 280; ++i) t.lengths[i] = 7;
  for (int i = 280; i < 288; ++i) t.lengths[i] = 8;
  bool ok = build_code(t.lengths, 288, t.lit_count, t.lit_symbol, t.lit_fast, kLitFastBits);
  for (int i = 0; i < 30; ++i) t.lengths[i] = 5;
  return ok && build_code(t.lengths, 30, t.dist_count, t.dist_symbol, t.dist_fast, kDistFastBits);
}

// Dynamic block header (RFC 1951 section 3.2.7). Consumes at most ~600 input bytes.
KOCR_HD bool dynamic_tables(State& s, Tables& t) {
  refill_bits(s, t);
  const int nlen = (int)take(s, 5) + 257, ndist = (int)take(s, 5) + 1, ncode = (int)take(s, 4) + 4;
  if (nlen > 286 || ndist > 30) return false;
  uint8_t cl[19];
  for (int i = 0; i < 19; ++i) cl[i] = 0;
  for (int i = 0; i < ncode; ++i) {
    refill_bits(s, t);
    cl[KOCR_TAB(kClOrder)[i]] = (uint8_t)take(s, 3);
  }
  // the code-length code reuses the distance arrays (they are rebuilt right after)
  if (!build_code(cl, 19, t.dist_count, t.dist_symbol, t.dist_fast, 7)) return false;
  int i = 0;
  while (i < nlen + ndist) {
    refill_bits(s, t);
    const int sym = decode_symbol(s, t.dist_fast, 7, t.dist_count, t.dist_symbol);
    if (sym < 0) return false;
    if (sym < 16) {
      t.lengths[i++] = (uint8_t)sym;
    } else {
      int rep, val = 0;
      if (sym == 16) {
        if (i == 0) return false;
        val = t.lengths[i - 1];
        rep = 3 + (int)take(s, 2);
      } else if (sym == 17) {
        rep = 3 + (int)take(s, 3);
      } else {
        rep = 11 + (int)take(s, 7);
      }
      if (i + rep > nlen + ndist) return false;
      while (rep--) t.lengths[i++] = (uint8_t)val;
    }
  }
  if (t.lengths[256] == 0) return false;  // no end-of-block code
  uint8_t dl[32];
  for (int k = 0; k < ndist; ++k) dl[k] = t.lengths[nlen + k];
  if (!build_code(t.lengths, nlen, t.lit_count, t.lit_symbol, t.lit_fast, kLitFastBits)) return false;
  return build_code(dl, ndist, t.dist_count, t.dist_symbol, t.dist_fast, kDistFastBits);
}

// Lane 0: decode until something needs the whole warp - a match to copy (match_len / match_dist set, out_pos not yet
// advanced), the input window to be refilled, the end of the stream, or an error. Literals and stored bytes are written
// straight to `out` on the way.
KOCR_HD int run(State& s, Tables& t, uint8_t* out) {
  for (;;) {
    if (consumed(s) > s.in_size) {
      s.status = kErrTruncated;
      return kEvError;
    }
    if (s.in_pos >= s.win_base + kHalf && s.win_base + kWindow < s.in_size && s.phase != 4) return kEvRefill;
    if (s.out_pos >= s.flush_at) return kEvFlush;
    refill_bits(s, t);
    switch (s.phase) {
      case 0: {  // zlib header (RFC 1950): CM = 8, window <= 32K, no preset dictionary, header checksum
        const uint32_t cmf = take(s, 8), flg = take(s, 8);
        if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 0x20) || ((cmf << 8) | flg) % 31) {
          s.status = kErrHeader;
          return kEvError;
        }
        s.phase = 1;
        break;
      }
      case 1: {
        s.last_block = (int)take(s, 1);
        const uint32_t type = take(s, 2);
        if (type == 0) {
          drop(s, s.bitcnt & 7);  // to the byte boundary
          refill_bits(s, t);
          const uint32_t len = take(s, 16), nlen = take(s, 16);
          if ((len ^ 0xffffu) != nlen) {
            s.status = kErrHeader;
            return kEvError;
          }
          s.stored_left = (int)len;
          s.phase = 2;
        } else if (type == 1) {
          if (!fixed_tables(t)) {
            s.status = kErrCode;
            return kEvError;
          }
          s.phase = 3;
        } else if (type == 2) {
          if (!dynamic_tables(s, t)) {
            s.status = kErrCode;
            return kEvError;
          }
          s.phase = 3;
        } else {
          s.status = kErrHeader;
          return kEvError;
        }
        break;
      }
      case 2: {  // stored bytes, a few per round so that the refill check above stays in charge
        int n = s.stored_left < 4 ? s.stored_left : 4;
        if (s.out_pos + n > s.out_size) {
          s.status = kErrOverflow;
          return kEvError;
        }
        s.stored_left -= n;
        while (n--) out[s.out_pos++ & s.out_mask] = (uint8_t)take(s, 8);
        if (s.stored_left == 0) s.phase = s.last_block ? 4 : 1;
        break;
      }
      case 3: {
        // literals are taken in a short inner loop (up to 8 symbols, < 50 input bytes: far inside the window's look-ahead) so
        // that the checks at the top of the outer loop are paid once per round, not once per byte
        int sym = 0;
        for (int k = 0; k < 8; ++k) {
          if (k) refill_bits(s, t);
          sym = decode_symbol(s, t.lit_fast, kLitFastBits, t.lit_count, t.lit_symbol);
          if (sym < 0 || sym >= 256) break;
          if (s.out_pos >= s.out_size) {
            s.status = kErrOverflow;
            return kEvError;
          }
          out[s.out_pos++ & s.out_mask] = (uint8_t)sym;
        }
        if (sym < 0) {
          s.status = kErrCode;
          return kEvError;
        }
        if (sym < 256) {
          // eight literals: next round
        } else if (sym == 256) {
          s.phase = s.last_block ? 4 : 1;
        } else {
          if (sym > 285) {
            s.status = kErrCode;
            return kEvError;
          }
          const int len = KOCR_TAB(kLenBase)[sym - 257] + (int)take(s, KOCR_TAB(kLenExtra)[sym - 257]);
          refill_bits(s, t);
          const int ds = decode_symbol(s, t.dist_fast, kDistFastBits, t.dist_count, t.dist_symbol);
          if (ds < 0 || ds > 29) {
            s.status = kErrCode;
            return kEvError;
          }
          const int dist = KOCR_TAB(kDistBase)[ds] + (int)take(s, KOCR_TAB(kDistExtra)[ds]);
          if (dist > s.out_pos) {
            s.status = kErrDistance;
            return kEvError;
          }
          if (s.out_pos + len > s.out_size) {
            s.status = kErrOverflow;
            return kEvError;
          }
          s.match_len = len;
          s.match_dist = dist;
          return kEvMatch;
        }
        break;
      }
      default:  // finished: the Adler-32 trailer is not checked (PNG's own chunk CRCs cover transport errors)
        if (s.out_pos != s.out_size) {
          s.status = kErrTruncated;
          return kEvError;
        }
        return kEvDone;
    }
  }
}

}  // namespace inflate

// ---- PNG scan-line reconstruction (PNG spec 9.2): Recon(x) from Filt(x) and the neighbours a (left), b (above), c (above left)
namespace pngfilter {
KOCR_HD int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
KOCR_HD int recon(int type, int filt, int a, int b, int c) {
  switch (type) {
    case 0: return filt;
    case 1: return (filt + a) & 255;
    case 2: return (filt + b) & 255;
    case 3: return (filt + ((a + b) >> 1)) & 255;
    default: return (filt + paeth(a, b, c)) & 255;
  }
}
}  // namespace pngfilter
}  // namespace kocr
