// Fused page preprocess kernel: bicubic-AA resize (horizontal + vertical fixed-point passes, uint8
// intermediate) -> normalise (LUT) -> temporal duplicate -> write in Qwen2-VL patch order.
//
// Stands in for Qwen2VLImageProcessor._preprocess (HF models/qwen2_vl/image_processing_qwen2_vl.py:148-232)
// which the reference reaches from karanta/training/pipeline_steps.py:289-294.
//
// HBM-bound by design (3.84 MB read + 15.6 MB bf16 written per letter page). Each input byte crosses HBM once (plus the
// filter halo between tiles), each output element is written once.
//   tile  = (page, strip of 28 output rows = one merge row, chunk of tile_w output columns); CTAs walk tiles persistently
//   load  the tile's input rows arrive in shared memory as whole 16-byte vectors (cp.async.cg, aligned 128-bit global
//         reads; a row keeps its sub-vector misalignment, recorded per row) and the NEXT tile's rows are requested before
//         this tile is computed: the global round trip hides behind a tile of arithmetic
//   pass 1 horizontal taps: one thread per output column keeps its filter as packed int16 pairs in registers and runs
//         down rows of equal alignment, two aligned words -> funnel shift -> dp2a (two taps per instruction); Pillow's
//         22-bit coefficients are split into two int16 halves (exact: the sum is linear)       -> mid[plane][row][x] u8
//   pass 2 vertical taps (only when the height changes): four columns per thread, row pairs interleaved with prmt so that
//         dp2a again takes two taps at a time                                                  -> res[plane][28][x] u8
//   pass 3 normalise (3x256 f32 LUT) + patch order: one thread per 28-pixel row of a merge cell, both temporal copies
//         written from one read into a staged copy of the tile's tokens, which are ONE contiguous span of pixel_values:
//         the tile leaves the SM as a single cp.async.bulk (TMA engine, no LSU store traffic)
// Gray pages (do_convert_rgb) are filtered once and normalised three times; interleaved RGB is split into planes in
// shared memory first.
#include <stdlib.h>

#include <map>
#include <tuple>
#include <vector>

#include "kocr_common.cuh"

namespace kocr {

struct PageJob {
  const uint8_t* src;
  const int32_t* hb;   // horizontal bounds [out_w][2] (first input column, taps), or null when the width does not change
  const uint32_t* hp;  // horizontal taps as int16 pairs [out_w][hkp * hsets]
  const int32_t* vb;   // vertical bounds [out_h][2], or null
  const uint32_t* vp;  // vertical pairs [out_h][vkp * vsets]
  long long token_base;
  long long src_bytes;    // size of the image buffer (loads never start past its 16-byte rounded end)
  long long chan_stride;  // bytes between the planes of a planar image
  int row_pitch;          // bytes between rows
  int pix_stride;         // bytes between horizontally adjacent pixels (3 = interleaved RGB)
  int in_h, in_w, layout;
  int out_h, out_w;
  int hkp, hsets, hprec, vkp, vsets, vprec;  // pairs per output, 1 set (int16 taps) or 2 (hi / lo halves of 22-bit taps)
  int tile_w, tiles_x, tile_base;
  int hcols;  // output columns per thread in the horizontal pass: 4 or 2 when their tap windows start within 4 input pixels of each other, else 1
};

struct TileInfo {
  int job, valid;
  int y0, x0, tw, r0, rows_in, xin0, ncols_in;
  int nseg, planes, Lp, nvec;   // staged segments (3 planar / 1), filtered planes (1 gray / 3), staged row pitch, vectors per row
  int lg_lpr;                   // log2 of the lanes that share one staged row in request_rows (>= nvec of them)
  int ng, nrg, n_iter, magic;   // pass1_multi: column groups, row slices, iterations per thread, ceil(2^16 / ng) (tid / ng without a division)
  long long n0;                 // first token of the tile
};

static constexpr int kStrip = 28;       // output rows per tile = patch * merge
static constexpr int kPatchDim = 1176;  // 3 * 2 * 14 * 14
static constexpr int kThreads = 256;
static constexpr int kMaxPairs = 4;     // fast passes hold up to 8 taps per output; larger filters take the generic loops
static constexpr int kMaxDynSmem = 200 * 1024;

// 1-D bulk async copy shared -> global (TMA engine, no tensor map): size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }
// two taps per instruction: d = c + s16(a.lo) * u8(b.byte[0|2]) + s16(a.hi) * u8(b.byte[1|3])  (lo: bytes 0,1; hi: bytes 2,3)
__device__ __forceinline__ int dp2a_lo_su(uint32_t taps, uint32_t px, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(px), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t taps, uint32_t px, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(px), "r"(c));
  return d;
}

// Which tile is it, and what does it read: evaluated by one thread, a tile ahead of its use. A CTA visits its tiles in
// increasing order, so the job is found by walking on from the previous one (`cur`, `next_base` live in that thread's registers).
__device__ void plan_tile(const PageJob* __restrict__ jobs, int n_jobs, int tile, int n_tiles, TileInfo& t, int& cur, int& next_base) {
  t.valid = tile < n_tiles;
  if (!t.valid) return;
  while (tile >= next_base) {
    ++cur;
    next_base = cur + 1 < n_jobs ? jobs[cur + 1].tile_base : 0x7fffffff;
  }
  const int lo = cur;
  const PageJob& j = jobs[lo];
  t.job = lo;
  const int local = tile - j.tile_base;
  const int sy = local / j.tiles_x, tx = local % j.tiles_x;
  t.y0 = sy * kStrip;
  t.x0 = tx * j.tile_w;
  t.tw = min(j.tile_w, j.out_w - t.x0);  // multiple of 28
  t.r0 = t.y0;
  t.rows_in = kStrip;
  if (j.vb) {
    t.r0 = j.vb[2 * t.y0];
    t.rows_in = j.vb[2 * (t.y0 + kStrip - 1)] + j.vb[2 * (t.y0 + kStrip - 1) + 1] - t.r0;
  }
  t.xin0 = t.x0;
  t.ncols_in = t.tw;
  if (j.hb) {
    t.xin0 = j.hb[2 * t.x0];
    t.ncols_in = j.hb[2 * (t.x0 + t.tw - 1)] + j.hb[2 * (t.x0 + t.tw - 1) + 1] - t.xin0;
  }
  t.nseg = j.layout == KOCR_LAYOUT_CHW ? 3 : 1;
  t.planes = j.layout == KOCR_LAYOUT_GRAY ? 1 : 3;
  // a staged row = the 16-byte vectors that cover its bytes, plus 16 bytes of slack for the passes' whole-word reads
  t.nvec = (15 + t.ncols_in * j.pix_stride + 15) >> 4;
  t.Lp = (t.nvec + 1) * 16;
  t.lg_lpr = 32 - __clz(t.nvec - 1);  // nvec >= 2
  if (j.hcols > 1 || !j.hb) {
    t.ng = t.tw / (j.hb ? j.hcols : 4);
    t.nrg = kThreads / t.ng;
    t.magic = (65536 + t.ng - 1) / t.ng;
    t.n_iter = (t.planes * t.rows_in + t.nrg - 1) / t.nrg;
  }
  t.n0 = j.token_base + ((long long)sy * (j.out_w / kStrip) + t.x0 / kStrip) * 4;
}

// Request the tile's input rows: every thread issues whole 16-byte vectors; row i of segment s lands at stage + (s*rows_in+i)*Lp
// starting `rowoff` bytes in (the row's misalignment against 16 bytes in global memory).
__device__ __forceinline__ void request_rows(const PageJob& j, const TileInfo& t, uint8_t* stage, uint8_t* rowoff) {
  const int nrows = t.nseg * t.rows_in;
  const uintptr_t img_end = (reinterpret_cast<uintptr_t>(j.src) + (uintptr_t)j.src_bytes + 15) & ~uintptr_t(15);
  const uint32_t stage_u32 = smem_u32(stage);
  const int lg = t.lg_lpr, nvec = t.nvec, rows_in = t.rows_in;
  // a power-of-two group of lanes per row: row and vector index are a shift and a mask (no integer division on this path)
  const int v = threadIdx.x & ((1 << lg) - 1);
  if (v >= nvec) return;
  const uintptr_t col0 = reinterpret_cast<uintptr_t>(j.src) + (uintptr_t)((long long)t.xin0 * j.pix_stride);
  for (int row = threadIdx.x >> lg; row < nrows; row += kThreads >> lg) {
    const int sg = (row >= rows_in) + (row >= 2 * rows_in), r = row - sg * rows_in;
    const uintptr_t p = col0 + (uintptr_t)(sg * j.chan_stride) + (uintptr_t)((long long)(t.r0 + r) * j.row_pitch);
    const uintptr_t a = (p & ~uintptr_t(15)) + 16u * (uintptr_t)v;
    if (v == 0) rowoff[row] = (uint8_t)(p & 15);
    if (a < img_end) cp_async16(stage_u32 + (uint32_t)(row * t.Lp + v * 16), reinterpret_cast<const void*>(a));
  }
}

// Two s32 -> saturated u8, packed under the upper half of c: d = sat(lo) | sat(hi) << 8 | c << 16
__device__ __forceinline__ uint32_t pack_sat_u8(int hi, int lo, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t shf_r_clamp(uint32_t lo, uint32_t hi, uint32_t sh) {
  uint32_t d;
  asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sh));
  return d;
}

// Horizontal pass, NC adjacent output columns per thread. Their tap windows start within four input pixels of each other
// (any resize up to ~1.3x down: the host checks it per page), so one row costs four aligned words and three funnel shifts that
// bring the first column's window to byte 0, then per column two clamped shifts by its own (loop-constant) offset, one dp2a per tap
// pair and a shift; the NC results leave as one packed store. Rows are walked with one flat index over (plane, row).
template <int NC, bool kSplit>
__device__ __forceinline__ void pass1_multi(const PageJob& j, const TileInfo& t, const uint8_t* rows_base, int row_pitch_s, int seg_rows,
                                            bool aligned_rows, const uint8_t* rowoff, const int32_t* xoff, const uint32_t* htab,
                                            uint8_t* mid, int mid_plane) {
  const int tw = t.tw, rows_in = t.rows_in, planes = t.planes, nseg = t.nseg;
  const int ng = t.ng;                               // column groups
  const int nrg = t.nrg;                             // row slices side by side
  const int grp = (int)(((uint32_t)threadIdx.x * (uint32_t)t.magic) >> 16), g = (int)threadIdx.x - grp * ng;
  if (grp >= nrg) return;
  const bool four = j.hkp > 3;
  // the tile's tap table is staged as [column][set][4 pairs] (zero padded): one 16-byte read per column and set
  const int4 xo4 = NC == 4 ? *reinterpret_cast<const int4*>(xoff + 4 * g) : make_int4(xoff[2 * g], xoff[2 * g + 1], 0, 0);
  const int xo0 = xo4.x;
  const int xoc[4] = {xo4.x, xo4.y, xo4.z, xo4.w};
  uint32_t dsh[NC], k[NC][4], l[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    dsh[c] = (uint32_t)(xoc[c] - xo0) * 8u;   // 0, 8, ... 32
    const uint4* kx = reinterpret_cast<const uint4*>(htab) + (NC * g + c) * (kSplit ? 2 : 1);
    const uint4 kk = kx[0];
    k[c][0] = kk.x; k[c][1] = kk.y; k[c][2] = kk.z; k[c][3] = kk.w;
    if (kSplit) {
      const uint4 ll = kx[1];
      l[c][0] = ll.x; l[c][1] = ll.y; l[c][2] = ll.z; l[c][3] = ll.w;
    } else {
      l[c][0] = l[c][1] = l[c][2] = l[c][3] = 0u;
    }
  }
  const int prec = j.hprec, round0 = 1 << (prec - 1);
  const int total = planes * rows_in;
  // (plane, row) is walked as one flat index `it`: the staged segments follow each other (seg_rows == rows_in), so the row's
  // address and its misalignment entry are linear in `it`; only the output plane pitch (rows_in + 1 rows) needs a wrap step
  const int nseg3 = nseg == 3 || aligned_rows;
  (void)nseg3;
  (void)seg_rows;
  uint32_t rp = smem_u32(rows_base) + (uint32_t)(grp * row_pitch_s);
  const uint32_t rstep = (uint32_t)(nrg * row_pitch_s);
  uint32_t rop = smem_u32(rowoff) + (uint32_t)grp;
  int r = grp, plw = 0;
  while (r >= rows_in) { r -= rows_in; ++plw; }
  uint32_t mp = smem_u32(mid) + (uint32_t)(plw * mid_plane + r * tw + NC * g);
  const uint32_t mstep = (uint32_t)(nrg * tw);
#pragma unroll 2
  for (int it = grp; it < total; it += nrg) {
    uint32_t ro = 0;
    if (!aligned_rows) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ro) : "r"(rop));
    const uint32_t boff = ro + (uint32_t)xo0;
    const uint32_t wa = rp + (boff & ~3u);
    const uint32_t sh = (boff & 3u) * 8u;
    const uint32_t w0 = lds_u32(wa), w1 = lds_u32(wa + 4), w2 = lds_u32(wa + 8), w3 = lds_u32(wa + 12);
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
    int o[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const uint32_t p = shf_r_clamp(v0, v1, dsh[c]), q = shf_r_clamp(v1, v2, dsh[c]);  // taps 0..3, 4..7 of column c
      int acc;
      if (!kSplit) {
        acc = dp2a_lo_su(k[c][0], p, round0);
        acc = dp2a_hi_su(k[c][1], p, acc);
        acc = dp2a_lo_su(k[c][2], q, acc);
        if (four) acc = dp2a_hi_su(k[c][3], q, acc);
      } else {  // 22-bit taps = hi * 2^11 + lo, both halves int16: sum = (sum_hi << 11) + sum_lo, exactly
        int hi = dp2a_lo_su(k[c][0], p, 0);
        hi = dp2a_hi_su(k[c][1], p, hi);
        hi = dp2a_lo_su(k[c][2], q, hi);
        int lo = dp2a_lo_su(l[c][0], p, round0);
        lo = dp2a_hi_su(l[c][1], p, lo);
        lo = dp2a_lo_su(l[c][2], q, lo);
        if (four) {
          hi = dp2a_hi_su(k[c][3], q, hi);
          lo = dp2a_hi_su(l[c][3], q, lo);
        }
        acc = hi * 2048 + lo;
      }
      o[c] = acc >> prec;
    }
    if (NC == 4) {
      const uint32_t v = pack_sat_u8(o[1], o[0], pack_sat_u8(o[NC - 1], o[NC - 2], 0u));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(mp), "r"(v) : "memory");
    } else {
      const uint32_t v = pack_sat_u8(o[1], o[0], 0u);
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(mp), "h"((unsigned short)v) : "memory");
    }
    rp += rstep;
    rop += (uint32_t)nrg;
    mp += mstep;
    r += nrg;
    while (r >= rows_in) {  // next plane of mid: its pitch is one row longer than the staged segment's
      r -= rows_in;
      mp += (uint32_t)tw;
    }
  }
}

// Width unchanged (no horizontal taps): the staged rows are re-aligned into mid, four pixels per thread and row - two aligned
// words, one funnel shift, one store. Same flat (plane, row) walk as pass1_multi.
__device__ __forceinline__ void pass1_copy(const TileInfo& t, const uint8_t* rows_base, int row_pitch_s, bool aligned_rows,
                                           const uint8_t* rowoff, uint8_t* mid, int mid_plane) {
  const int tw = t.tw, rows_in = t.rows_in;
  const int ng = t.ng, nrg = t.nrg;
  const int grp = (int)(((uint32_t)threadIdx.x * (uint32_t)t.magic) >> 16), g = (int)threadIdx.x - grp * ng;
  if (grp >= nrg) return;
  const int total = t.planes * rows_in;
  uint32_t rp = smem_u32(rows_base) + (uint32_t)(grp * row_pitch_s);
  const uint32_t rstep = (uint32_t)(nrg * row_pitch_s);
  uint32_t rop = smem_u32(rowoff) + (uint32_t)grp;
  int r = grp, plw = 0;
  while (r >= rows_in) { r -= rows_in; ++plw; }
  uint32_t mp = smem_u32(mid) + (uint32_t)(plw * mid_plane + r * tw + 4 * g);
  const uint32_t mstep = (uint32_t)(nrg * tw);
#pragma unroll 2
  for (int it = grp; it < total; it += nrg) {
    uint32_t ro = 0;
    if (!aligned_rows) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ro) : "r"(rop));
    const uint32_t boff = ro + 4u * (uint32_t)g;
    const uint32_t wa = rp + (boff & ~3u);
    const uint32_t v = __funnelshift_r(lds_u32(wa), lds_u32(wa + 4), (boff & 3u) * 8u);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(mp), "r"(v) : "memory");
    rp += rstep;
    rop += (uint32_t)nrg;
    mp += mstep;
    r += nrg;
    while (r >= rows_in) {
      r -= rows_in;
      mp += (uint32_t)tw;
    }
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 3) preprocess_kernel(const PageJob* __restrict__ jobs, int n_jobs,
                                                              int n_tiles, const float* __restrict__ lut_g,
                                                              void* __restrict__ out, int stage_bytes, int planar_off, int tab_off,
                                                              int mid_off, int res_off, int out_off) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(1024) float lut[768];  // 1 KB per channel: a table address is base | (byte << 2)
  __shared__ TileInfo tinfo[2];
  __shared__ uint8_t rowoff[2][3 * 256];
  for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = lut_g[i];

  int buf = 0;
  int cur = 0, next_base = 0;  // thread 0: job of the last planned tile, first tile of the job after it
  if (threadIdx.x == 0) {
    next_base = n_jobs > 1 ? jobs[1].tile_base : 0x7fffffff;
    plan_tile(jobs, n_jobs, blockIdx.x, n_tiles, tinfo[0], cur, next_base);
  }
  __syncthreads();
  if (tinfo[0].valid) request_rows(jobs[tinfo[0].job], tinfo[0], smem, rowoff[0]);
  cp_async_commit();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    // ---- a tile ahead: thread 0 plans the next tile (a few dependent global reads) while the others fetch this tile's tap tables
    // and wait for its rows; everybody meets at the barrier before pass 1
    if (threadIdx.x == 0) plan_tile(jobs, n_jobs, tile + gridDim.x, n_tiles, tinfo[buf ^ 1], cur, next_base);
    const TileInfo& t = tinfo[buf];
    const PageJob& j = jobs[t.job];
    const int tw = t.tw, rows_in = t.rows_in, planes = t.planes;
    uint8_t* stage = smem + buf * stage_bytes;
    uint8_t* mid = smem + mid_off;                            // [planes][rows_in + 1][tw]
    uint8_t* res = j.vb ? smem + res_off : mid;               // [planes][28][tw]
    const int mid_plane = (rows_in + 1) * tw;
    const int hpw = j.hkp * j.hsets;                          // u32 per output column in the tap table
    // ---- this tile's horizontal taps: (first input column relative to the tile, pairs) per output column, to shared memory
    int32_t* xoff = reinterpret_cast<int32_t*>(smem + tab_off);            // [tw]
    uint32_t* htab = reinterpret_cast<uint32_t*>(smem + tab_off) + 112;    // [tw][hpw]
    if (j.hb) {
      for (int x = threadIdx.x; x < tw; x += kThreads) xoff[x] = j.hb[2 * (t.x0 + x)] - t.xin0;
      if (j.hcols > 1) {  // pass1_multi: [column][set][4] with zero padding
        const int hkp = j.hkp, lg_per = j.hsets == 2 ? 3 : 2, per = 1 << lg_per;
        for (int i = threadIdx.x; i < tw * per; i += kThreads) {
          const int x = i >> lg_per, e = i & (per - 1), st = e >> 2, kk = e & 3;
          htab[i] = kk < hkp ? j.hp[(size_t)(t.x0 + x) * hpw + st * hkp + kk] : 0u;
        }
      } else {
        for (int i = threadIdx.x; i < tw * hpw; i += kThreads) htab[i] = j.hp[(size_t)t.x0 * hpw + i];
      }
    }
    cp_async_wait<0>();  // this tile's rows have landed
    __syncthreads();
    // the next tile's rows are requested now, into the other stage buffer (its last reader, pass 1 of the previous tile, is
    // several barriers behind): they have passes 1-3 of this tile to arrive
    if (tinfo[buf ^ 1].valid)
      request_rows(jobs[tinfo[buf ^ 1].job], tinfo[buf ^ 1], smem + (buf ^ 1) * stage_bytes, rowoff[buf ^ 1]);
    cp_async_commit();

    // ---- interleaved RGB: split the staged rows into planes (aligned rows, pitch Lq) so that pass 1 sees adjacent taps
    const uint8_t* rows_base = stage;
    int row_pitch_s = t.Lp, seg_rows = rows_in;
    bool aligned_rows = false;
    if (j.layout == KOCR_LAYOUT_HWC) {
      uint8_t* planar = smem + planar_off;
      const int Lq = (t.ncols_in + 20 + 3) & ~3;
      for (int i = threadIdx.x; i < rows_in * t.ncols_in; i += kThreads) {
        const int r = i / t.ncols_in, x = i - r * t.ncols_in;
        const uint8_t* p = stage + r * t.Lp + rowoff[buf][r] + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) planar[(c * rows_in + r) * Lq + x] = p[c];
      }
      __syncthreads();
      rows_base = planar;
      row_pitch_s = Lq;
      aligned_rows = true;
    }

    // ---- pass 1: horizontal. Thread = (output column x, alignment class rho): the rows r = rho, rho+4, ... of a segment
    // share their misalignment modulo 4 when walked with a constant pitch, so the word offset and the funnel-shift amount of
    // the column's tap window are loop constants; per row: three aligned words, two funnel shifts, one dp2a per tap pair.
    const int nseg = t.nseg;
    if (!j.hb) {
      pass1_copy(t, rows_base, row_pitch_s, aligned_rows, rowoff[buf], mid, mid_plane);
    } else if (j.hkp <= kMaxPairs && j.hcols > 1) {
      const uint8_t* ro = rowoff[buf];
      if (j.hcols == 4) {
        if (j.hsets == 2) pass1_multi<4, true>(j, t, rows_base, row_pitch_s, seg_rows, aligned_rows, ro, xoff, htab, mid, mid_plane);
        else pass1_multi<4, false>(j, t, rows_base, row_pitch_s, seg_rows, aligned_rows, ro, xoff, htab, mid, mid_plane);
      } else {
        if (j.hsets == 2) pass1_multi<2, true>(j, t, rows_base, row_pitch_s, seg_rows, aligned_rows, ro, xoff, htab, mid, mid_plane);
        else pass1_multi<2, false>(j, t, rows_base, row_pitch_s, seg_rows, aligned_rows, ro, xoff, htab, mid, mid_plane);
      }
    } else if (j.hb && j.hkp <= kMaxPairs) {
      // Thread = (output column x, row group): the column's taps (up to four int16 pairs, two sets for Pillow's split 22-bit
      // taps) stay in registers; per row: the row's misalignment + the column's offset give an aligned word address and a funnel-
      // shift amount, three aligned words -> two shifts -> the eight tap bytes, one dp2a per tap pair.
      const int nrg = kThreads / tw;                    // row groups side by side (84 columns -> 3, 56 -> 4, 112 -> 2)
      const int x = threadIdx.x % tw, grp = threadIdx.x / tw;
      if (grp < nrg) {
        const int xo = xoff[x];
        const uint32_t* kx = htab + x * hpw;
        const int hkp = j.hkp;
        const bool split = j.hsets == 2;
        const uint32_t k0 = kx[0], k1 = hkp > 1 ? kx[1] : 0u, k2 = hkp > 2 ? kx[2] : 0u, k3 = hkp > 3 ? kx[3] : 0u;
        const uint32_t l0 = split ? kx[hkp] : 0u, l1 = split && hkp > 1 ? kx[hkp + 1] : 0u, l2 = split && hkp > 2 ? kx[hkp + 2] : 0u,
                       l3 = split && hkp > 3 ? kx[hkp + 3] : 0u;
        const int prec = j.hprec, round0 = 1 << (prec - 1);
        const int rstep = nrg * row_pitch_s, mstep = nrg * tw;
        for (int pl = 0; pl < planes; ++pl) {
          const int sg = nseg == 3 ? pl : 0;
          const uint8_t* rp = rows_base + (size_t)((aligned_rows ? pl : sg) * seg_rows + grp) * row_pitch_s;
          const uint8_t* rop = &rowoff[buf][sg * rows_in + grp];
          uint8_t* mp = mid + pl * mid_plane + grp * tw + x;
#pragma unroll 2
          for (int r = grp; r < rows_in; r += nrg, rp += rstep, rop += nrg, mp += mstep) {
            const int boff = (aligned_rows ? 0 : (int)*rop) + xo;   // byte offset of the column's first tap in the staged row
            const uint32_t* w = reinterpret_cast<const uint32_t*>(rp + (boff & ~3));
            const int sh = (boff & 3) * 8;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            const uint32_t p = __funnelshift_r(w0, w1, sh), q = __funnelshift_r(w1, w2, sh);  // taps 0..3, 4..7
            int acc;
            if (!split) {
              acc = dp2a_lo_su(k0, p, round0);
              acc = dp2a_hi_su(k1, p, acc);
              acc = dp2a_lo_su(k2, q, acc);
              acc = dp2a_hi_su(k3, q, acc);
            } else {  // 22-bit taps = hi * 2^11 + lo, both halves int16: sum = (sum_hi << 11) + sum_lo, exactly
              int hi = dp2a_lo_su(k0, p, 0);
              hi = dp2a_hi_su(k1, p, hi);
              hi = dp2a_lo_su(k2, q, hi);
              hi = dp2a_hi_su(k3, q, hi);
              int lo = dp2a_lo_su(l0, p, round0);
              lo = dp2a_hi_su(l1, p, lo);
              lo = dp2a_lo_su(l2, q, lo);
              lo = dp2a_hi_su(l3, q, lo);
              acc = hi * 2048 + lo;
            }
            *mp = (uint8_t)sat_u8(acc >> prec);
          }
        }
      }
    } else {
      // generic: any filter length (strong downscales), or no horizontal resize at all (copy)
      for (int i = threadIdx.x; i < planes * rows_in * tw; i += kThreads) {
        const int x = i % tw, rr = i / tw;
        const int pl = rr / rows_in, r = rr - pl * rows_in;
        const int sg = nseg == 3 ? pl : 0;
        const uint8_t* rowp = aligned_rows ? rows_base + (size_t)(pl * seg_rows + r) * row_pitch_s
                                           : rows_base + (size_t)(sg * rows_in + r) * row_pitch_s + rowoff[buf][sg * rows_in + r];
        int v;
        if (j.hb) {
          const int xmin = j.hb[2 * (t.x0 + x)] - t.xin0, cnt = j.hb[2 * (t.x0 + x) + 1];
          const uint32_t* kk = j.hp + (size_t)(t.x0 + x) * hpw;
          int hi = 0, lo = 0;
          for (int k = 0; k < cnt; ++k) {
            const int px = rowp[xmin + k];
            const uint32_t w = kk[k >> 1];
            hi += px * (int)(int16_t)((k & 1) ? (w >> 16) : (w & 0xffff));
            if (j.hsets == 2) {
              const uint32_t w2 = kk[j.hkp + (k >> 1)];
              lo += px * (int)(int16_t)((k & 1) ? (w2 >> 16) : (w2 & 0xffff));
            }
          }
          const int acc = j.hsets == 2 ? hi * 2048 + lo : hi;
          v = sat_u8((acc + (1 << (j.hprec - 1))) >> j.hprec);
        } else {
          v = rowp[x];
        }
        mid[pl * mid_plane + r * tw + x] = (uint8_t)v;
      }
    }
    __syncthreads();

    // ---- pass 2: vertical mid -> res. Thread = (4 columns, plane, third of the strip): rows t and t+1 of the four columns are
    // interleaved by prmt into (t, t+1) byte pairs, so one dp2a applies a pair of taps to one column.
    if (j.vb) {
      const int groups = tw >> 2;
      const int vpw = j.vkp * j.vsets;
      const int items = planes * groups;
      const int ysplit = max(1, min(kStrip, kThreads / items));        // concurrent slices of the 28 output rows
      const int it = threadIdx.x % items, ys = threadIdx.x / items;
      if (ys < ysplit) {
        const int pl = it / groups, xg = it - pl * groups;
        const int round0 = 1 << (j.vprec - 1);
        for (int y = ys; y < kStrip; y += ysplit) {
          const int ymin = j.vb[2 * (t.y0 + y)] - t.r0, cnt = j.vb[2 * (t.y0 + y) + 1];
          const uint32_t* kk = j.vp + (size_t)(t.y0 + y) * vpw;
          const uint32_t a0 = smem_u32(mid + pl * mid_plane + ymin * tw + 4 * xg);
          int hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
          for (int k = 0; 2 * k < cnt; ++k) {
            const uint32_t wa = lds_u32(a0 + (uint32_t)(2 * k * tw)), wb = lds_u32(a0 + (uint32_t)((2 * k + 1) * tw));  // row ymin+cnt may be one past: its tap is 0
            const uint32_t ab = __byte_perm(wa, wb, 0x5140), cd = __byte_perm(wa, wb, 0x7362);   // (a0 b0 a1 b1), (a2 b2 a3 b3)
            const uint32_t k0 = __ldg(kk + k);
            hi[0] = dp2a_lo_su(k0, ab, hi[0]);
            hi[1] = dp2a_hi_su(k0, ab, hi[1]);
            hi[2] = dp2a_lo_su(k0, cd, hi[2]);
            hi[3] = dp2a_hi_su(k0, cd, hi[3]);
            if (j.vsets == 2) {
              const uint32_t k1 = __ldg(kk + j.vkp + k);
              lo[0] = dp2a_lo_su(k1, ab, lo[0]);
              lo[1] = dp2a_hi_su(k1, ab, lo[1]);
              lo[2] = dp2a_lo_su(k1, cd, lo[2]);
              lo[3] = dp2a_hi_su(k1, cd, lo[3]);
            }
          }
          uint32_t packed = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int acc = j.vsets == 2 ? hi[e] * 2048 + lo[e] : hi[e];
            packed |= (uint32_t)sat_u8((acc + round0) >> j.vprec) << (8 * e);
          }
          *reinterpret_cast<uint32_t*>(res + (pl * kStrip + y) * tw + 4 * xg) = packed;
        }
      }
      __syncthreads();
    }

    // ---- pass 3: normalise + patch order, staged in smem. Thread = (merge cell, row of the strip, channel): 28 pixels = seven
    // aligned words, 28 LUT reads, and the two 14-pixel runs (mw = 0 / 1) written for both temporal copies. The tile's tokens
    // are ONE contiguous span of pixel_values, so the whole staged tile leaves with a single bulk async copy.
    const int cells = tw / kStrip;
    constexpr int kElt = kBf16 ? 2 : 4;
    uint8_t* obuf = smem + out_off;
    if (threadIdx.x == 0) bulk_store_wait_read();  // the previous tile's bulk copy has finished reading obuf
    __syncthreads();
    const int res_plane = (j.vb ? kStrip : rows_in + 1) * tw;
    for (int item = threadIdx.x; item < cells * kStrip * 3; item += kThreads) {
      const int c = item / (cells * kStrip), rem = item - c * (cells * kStrip);
      const int cell = rem / kStrip, y = rem - cell * kStrip;
      const int mh = y >= 14 ? 1 : 0, py = y - 14 * mh;
      const uint32_t* p = reinterpret_cast<const uint32_t*>(res + (planes == 1 ? 0 : c) * res_plane + y * tw + cell * kStrip);
      const uint32_t lb = smem_u32(lut) + (uint32_t)c * 1024u;
      auto lut_at = [lb](uint32_t idx4) {  // idx4 = byte << 2 (already masked)
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(lb | idx4));
        return v;
      };
      uint8_t* d0 = obuf + ((size_t)(cell * 4 + mh * 2) * kPatchDim + c * 392 + py * 14) * kElt;  // token mw = 0; mw = 1 is kPatchDim further
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const uint32_t u = p[k];
        const float v0 = lut_at((u << 2) & 0x3fcu), v1 = lut_at((u >> 6) & 0x3fcu), v2 = lut_at((u >> 14) & 0x3fcu), v3 = lut_at((u >> 22) & 0x3fcu);
        // pixels 4k..4k+3 of the 28: pixel pairs never straddle the two tokens (14 is even)
        const int pa = 4 * k, pb = 4 * k + 2;
        uint8_t* da = d0 + ((pa >= 14 ? kPatchDim - 14 : 0) + pa) * kElt;
        uint8_t* db = d0 + ((pb >= 14 ? kPatchDim - 14 : 0) + pb) * kElt;
        if (kBf16) {
          const uint32_t ka = pack_bf16(v0, v1), kb = pack_bf16(v2, v3);
          *reinterpret_cast<uint32_t*>(da) = ka;
          *reinterpret_cast<uint32_t*>(da + 196 * kElt) = ka;  // tp = 1 copy
          *reinterpret_cast<uint32_t*>(db) = kb;
          *reinterpret_cast<uint32_t*>(db + 196 * kElt) = kb;
        } else {
          *reinterpret_cast<float2*>(da) = make_float2(v0, v1);
          *reinterpret_cast<float2*>(da + 196 * kElt) = make_float2(v0, v1);
          *reinterpret_cast<float2*>(db) = make_float2(v2, v3);
          *reinterpret_cast<float2*>(db + 196 * kElt) = make_float2(v2, v3);
        }
      }
    }
    fence_proxy_async();  // make the generic-proxy smem writes visible to the bulk copy engine
    __syncthreads();
    if (threadIdx.x == 0)
      bulk_store(reinterpret_cast<uint8_t*>(out) + t.n0 * kPatchDim * kElt, obuf, (uint32_t)(cells * 4 * kPatchDim * kElt));
  }
  cp_async_wait<0>();
  if (threadIdx.x == 0) bulk_store_wait_all();
}

struct AxisTable {
  std::vector<int32_t> bounds;    // [out][2]: first input index, taps
  std::vector<uint32_t> pairs;    // [out][kp * sets]: taps as int16 pairs (low half = even tap); 22-bit taps as hi set then lo set
  int ksize = 0, prec = 0, kp = 0, sets = 1;
  int max_rows = 0;               // input rows a 28-row output strip can touch
  int hcols = 1;                  // adjacent outputs whose windows start within 4 inputs of each other: 4, 2 or 1 (pass1_multi)
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
    std::vector<int32_t> coeffs((size_t)out_size * t.ksize);
    int rc = resample_coeffs(in_size, out_size, mode, t.bounds.data(), coeffs.data(), &t.prec);
    if (rc) return rc;
    for (int y0 = 0; y0 + kStrip <= out_size; y0 += kStrip) {
      int last = y0 + kStrip - 1;
      t.max_rows = std::max(t.max_rows, t.bounds[2 * last] + t.bounds[2 * last + 1] - t.bounds[2 * y0]);
    }
    for (int nc : {4, 2}) {
      if (out_size % nc) continue;
      bool ok = true;
      for (int o = 0; o < out_size && ok; o += nc) {
        const int d = t.bounds[2 * (o + nc - 1)] - t.bounds[2 * o];
        ok = d >= 0 && d <= 4;
        for (int c = 1; c < nc && ok; ++c) ok = t.bounds[2 * (o + c)] >= t.bounds[2 * (o + c - 1)];
      }
      if (ok) { t.hcols = nc; break; }
    }
    // dp2a operands. ATen's taps are int16 already; Pillow's (22-bit precision) are split as c = hi * 2^11 + lo with
    // lo in [0, 2047]: sum(c * p) = (sum(hi * p) << 11) + sum(lo * p) exactly, and both halves fit int16.
    t.kp = (t.ksize + 1) / 2;
    t.sets = mode == KOCR_RESIZE_PIL ? 2 : 1;
    t.pairs.assign((size_t)out_size * t.kp * t.sets, 0u);
    for (int o = 0; o < out_size; ++o) {
      const int cnt = t.bounds[2 * o + 1];
      for (int k = 0; k < cnt; ++k) {
        const int32_t c = coeffs[(size_t)o * t.ksize + k];
        const int32_t hi = t.sets == 2 ? (c >> 11) : c, lo = c & 2047;
        if (hi < -32768 || hi > 32767) return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: filter tap out of the int16 range");
        uint32_t* row = &t.pairs[(size_t)o * t.kp * t.sets];
        row[k >> 1] |= (uint32_t)(uint16_t)(int16_t)hi << (16 * (k & 1));
        if (t.sets == 2) row[t.kp + (k >> 1)] |= (uint32_t)(uint16_t)lo << (16 * (k & 1));
      }
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

  // ---- plan on the host: sizes, filter banks (deduplicated), tiles, shared-memory carve-up
  if (table_cache().size() > 256) table_cache().clear();  // bound the per-thread cache (pointers below stay valid)
  std::vector<PageJob> jobs(n_images);
  struct Need { size_t off_b, off_p; };
  std::map<const AxisTable*, Need> needs;
  size_t table_bytes = 0;
  auto want = [&](const AxisTable* t) {
    if (needs.count(t)) return;
    Need n{table_bytes, 0};
    table_bytes += t->bounds.size() * 4;
    n.off_p = table_bytes;
    table_bytes += t->pairs.size() * 4;
    needs[t] = n;
  };
  std::vector<const AxisTable*> ht(n_images, nullptr), vt(n_images, nullptr);
  int64_t tokens = 0;
  int tiles = 0;
  int stage_bytes = 0, planar_bytes = 0, tab_bytes = 112 * 4, mid_bytes = 0, res_bytes = 0, max_tw = 0;
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
    else if (im.layout == KOCR_LAYOUT_HWC) { j.pix_stride = 3; j.row_pitch = 3 * im.width; j.chan_stride = 0; }
    else { j.pix_stride = 1; j.row_pitch = im.width; j.chan_stride = 0; }
    j.out_h = oh; j.out_w = ow;
    j.token_base = tokens;
    const int rows_in = vt[i] ? vt[i]->max_rows : kStrip;
    const int nseg = im.layout == KOCR_LAYOUT_CHW ? 3 : 1, planes = im.layout == KOCR_LAYOUT_GRAY ? 1 : 3;
    if (3 * rows_in > 768)
      return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: vertical downscale factor too large for one tile");
    // shared memory of one CTA as a function of the tile width
    const double scale = std::max(1.0, (double)im.width / ow);
    const int elt = out_dtype == KOCR_DTYPE_BF16 ? 2 : 4;
    struct Carve { int stage, planar, tab, mid, res, outb, total; };
    auto carve = [&](int w) {
      Carve c{};
      // the widest span of input columns a tile can touch (a bound: the kernel works out the exact one per tile)
      const int ncols = ht[i] ? (int)(w * scale) + 2 * ht[i]->ksize + 8 : w;
      const int nvec = (15 + ncols * j.pix_stride + 15) >> 4;
      c.stage = nseg * rows_in * (nvec + 1) * 16;
      c.planar = im.layout == KOCR_LAYOUT_HWC ? 3 * rows_in * ((ncols + 20 + 3) & ~3) + 16 : 0;
      c.tab = 112 * 4 + (ht[i] ? w * std::max(ht[i]->kp, 4) * ht[i]->sets * 4 : 0);
      c.mid = planes * (rows_in + 1) * w + 16;
      c.res = vt[i] ? planes * kStrip * w : 0;
      c.outb = (w / kStrip) * 4 * kPatchDim * elt;
      c.total = 2 * c.stage + c.planar + c.tab + c.mid + c.res + c.outb + 256;
      return c;
    };
    // default tile: 84 columns = 12 tokens (28 KB of bf16 output); ~61 KB per CTA for a letter page, three CTAs per SM. Strong
    // downscales and the f32 drop-in output take narrower tiles; KOCR_PRE_TW overrides the default for experiments.
    static const int tw_default = [] {
      const char* e = getenv("KOCR_PRE_TW");
      const int v = e ? atoi(e) : 84;
      return (v >= kStrip && v <= 112 && v % kStrip == 0) ? v : 84;
    }();
    int tw = std::min(tw_default, ow);
    while (tw > kStrip && carve(tw).total > 72 * 1024) tw -= kStrip;
    const Carve cv = carve(tw);
    if (cv.total > kMaxDynSmem) return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: downscale factor too large for one tile");
    j.tile_w = tw;
    j.tiles_x = (ow + tw - 1) / tw;
    j.tile_base = tiles;
    tiles += j.tiles_x * (oh / kStrip);
    max_tw = std::max(max_tw, tw);
    stage_bytes = std::max(stage_bytes, cv.stage);
    planar_bytes = std::max(planar_bytes, cv.planar);
    tab_bytes = std::max(tab_bytes, cv.tab);
    mid_bytes = std::max(mid_bytes, cv.mid);
    res_bytes = std::max(res_bytes, cv.res);
    grid_thw_out[3 * i] = 1;
    grid_thw_out[3 * i + 1] = oh / 14;
    grid_thw_out[3 * i + 2] = ow / 14;
    tokens += (int64_t)(oh / 14) * (ow / 14);
  }
  if (tokens > capacity_rows)
    return fail(KOCR_ERR_INVALID, "kocr_preprocess: pixel_values buffer too small for this batch");
  auto up = [](int v, int a) { return (v + a - 1) / a * a; };
  stage_bytes = up(stage_bytes, 16);
  const int planar_off = 2 * stage_bytes;
  const int tab_off = planar_off + up(planar_bytes, 16);
  const int mid_off = tab_off + up(tab_bytes, 16);
  const int res_off = mid_off + up(mid_bytes, 16);
  const int out_off = up(res_off + res_bytes, 128);
  const int out_bytes = (max_tw / kStrip) * 4 * kPatchDim * (out_dtype == KOCR_DTYPE_BF16 ? 2 : 4);
  const int smem = out_off + out_bytes;
  if (smem > kMaxDynSmem) return fail(KOCR_ERR_UNSUPPORTED, "kocr_preprocess: tile does not fit in shared memory");

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
    memcpy(hb + kv.second.off_p, kv.first->pairs.data(), kv.first->pairs.size() * 4);
  }
  for (int i = 0; i < n_images; ++i) {
    PageJob& j = jobs[i];
    if (ht[i]) {
      j.hb = reinterpret_cast<const int32_t*>(db + needs[ht[i]].off_b);
      j.hp = reinterpret_cast<const uint32_t*>(db + needs[ht[i]].off_p);
      j.hkp = ht[i]->kp; j.hsets = ht[i]->sets; j.hprec = ht[i]->prec;
      j.hcols = ht[i]->hcols;
    }
    if (vt[i]) {
      j.vb = reinterpret_cast<const int32_t*>(db + needs[vt[i]].off_b);
      j.vp = reinterpret_cast<const uint32_t*>(db + needs[vt[i]].off_p);
      j.vkp = vt[i]->kp; j.vsets = vt[i]->sets; j.vprec = vt[i]->prec;
    }
  }
  memcpy(hb + jobs_off, jobs.data(), sizeof(PageJob) * n_images);
  void* d;
  rc = ctx->stage_commit(slot, total, stream, &d);
  if (rc) {
    guard.slot = -1;  // stage_commit released it
    return rc;
  }

  // persistent CTAs: as many as are resident at once (each prefetches its next tile while it works on the current one)
  int occ = 0;
  if (out_dtype == KOCR_DTYPE_BF16) {
    if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&preprocess_kernel<true>), kMaxDynSmem))) return rc;
    KOCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, preprocess_kernel<true>, kThreads, smem));
  } else {
    if ((rc = ctx->opt_in_smem(reinterpret_cast<const void*>(&preprocess_kernel<false>), kMaxDynSmem))) return rc;
    KOCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, preprocess_kernel<false>, kThreads, smem));
  }
  const int grid = std::min(tiles, ctx->num_sms * std::max(occ, 1));
  const PageJob* d_jobs = reinterpret_cast<const PageJob*>(db + jobs_off);
  ProfScope ps(ctx, kProfPreprocess, stream);
  if (out_dtype == KOCR_DTYPE_BF16) {
    preprocess_kernel<true><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode], pixel_values,
                                                              stage_bytes, planar_off, tab_off, mid_off, res_off, out_off);
  } else {
    preprocess_kernel<false><<<grid, kThreads, smem, stream>>>(d_jobs, n_images, tiles, ctx->d_lut[resize_mode], pixel_values,
                                                               stage_bytes, planar_off, tab_off, mid_off, res_off, out_off);
  }
  KOCR_LAUNCH_CHECK("preprocess_kernel");
  return KOCR_OK;
}
