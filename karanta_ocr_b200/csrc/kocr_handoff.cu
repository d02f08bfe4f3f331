// LLM hand-off after the vision tower: M-RoPE position ids (host planning) and the
// masked scatter of the image embeddings into inputs_embeds (one HBM-bound row-copy kernel).
#include <algorithm>
#include <vector>

#include "kocr_common.cuh"

namespace kocr {

// HF modeling_qwen2_vl.py:990-1092 (get_rope_index) + :934-988 (get_vision_position_ids), still images only.
int mrope_position_ids(const int64_t* ids, const int64_t* mask, int B, int L, const int64_t* grid, int n_images, int64_t image_token,
                       int merge, int64_t* pos, int64_t* deltas) {
  int img = 0;
  for (int64_t i = 0; i < (int64_t)3 * B * L; ++i) pos[i] = 0;
  std::vector<int> keep;
  for (int b = 0; b < B; ++b) {
    keep.clear();
    for (int i = 0; i < L; ++i)
      if (!mask || mask[(int64_t)b * L + i] != 0) keep.push_back(i);
    int64_t cur = 0, mx = -1;
    size_t k = 0;
    auto put = [&](int slot, int64_t t, int64_t h, int64_t w) {
      const int64_t o = (int64_t)b * L + keep[slot];
      pos[o] = t;
      pos[(int64_t)B * L + o] = h;
      pos[(int64_t)2 * B * L + o] = w;
      mx = std::max(mx, std::max(t, std::max(h, w)));
    };
    while (k < keep.size()) {
      const bool is_img = ids[(int64_t)b * L + keep[k]] == image_token;
      size_t e = k;
      while (e < keep.size() && (ids[(int64_t)b * L + keep[e]] == image_token) == is_img) ++e;
      const int64_t run = (int64_t)(e - k);
      if (!is_img) {
        for (int64_t j = 0; j < run; ++j) put((int)(k + j), cur + j, cur + j, cur + j);
        cur += run;
      } else {
        if (img >= n_images) return fail(KOCR_ERR_INVALID, "mrope_position_ids: more image runs in input_ids than rows in image_grid_thw");
        const int64_t t = grid[3 * img], gh = grid[3 * img + 1] / merge, gw = grid[3 * img + 2] / merge;
        if (t != 1) return fail(KOCR_ERR_UNSUPPORTED, "mrope_position_ids: only still images (t == 1) are supported on this path");
        if (gh * gw != run) {
          char buf[160];
          snprintf(buf, sizeof buf, "mrope_position_ids: image %d has %lld placeholder tokens but its grid needs %lld", img, (long long)run,
                   (long long)(gh * gw));
          return fail(KOCR_ERR_INVALID, buf);
        }
        for (int64_t j = 0; j < run; ++j) put((int)(k + j), cur, cur + j / gw, cur + j % gw);
        cur += std::max(gh, gw);
        ++img;
      }
      k = e;
    }
    deltas[b] = mx + 1 - (int64_t)keep.size();
  }
  return KOCR_OK;
}

// The same planning with the semantics of transformers 4.5x, the line the reference pins (4.53.3, /root/reference/uv.lock:
// 2168-2169; ==4.51.3 in requirements.txt:65). Differences from 5.x, restated from the 4.5x get_rope_index: (1) position_ids
// start as ones, so padded positions hold 1, not 0; (2) delta = max + 1 - L with L the PADDED length of the row, which is what
// generate() adds to cache_position for left-padded batches; (3) images are found as <|vision_start|> followed by an image
// token (vision_start < 0: every maximal run of image tokens starts one image) and each image consumes exactly its grid's
// t*h*w/merge^2 tokens from the next image token on, so two images with no separator between them are split by grid size
// instead of being rejected. That version is not installed in the build image: this mode has no golden ("parity unpinned").
int mrope_position_ids_v45(const int64_t* ids, const int64_t* mask, int B, int L, const int64_t* grid, int n_images, int64_t image_token,
                           int64_t vision_start, int merge, int64_t* pos, int64_t* deltas) {
  int img = 0;
  for (int64_t i = 0; i < (int64_t)3 * B * L; ++i) pos[i] = 1;
  std::vector<int> keep;
  std::vector<int64_t> tok;
  for (int b = 0; b < B; ++b) {
    keep.clear();
    tok.clear();
    for (int i = 0; i < L; ++i)
      if (!mask || mask[(int64_t)b * L + i] == 1) {
        keep.push_back(i);
        tok.push_back(ids[(int64_t)b * L + i]);
      }
    const size_t n = tok.size();
    int n_img_row = 0;
    for (size_t i = 0; i < n; ++i) {
      if (vision_start >= 0) n_img_row += (tok[i] == vision_start && i + 1 < n && tok[i + 1] == image_token);
      else n_img_row += (tok[i] == image_token && (i == 0 || tok[i - 1] != image_token));
    }
    int64_t mx = -1, next = 0;  // next = max position so far + 1
    size_t st = 0;
    auto put = [&](size_t slot, int64_t t, int64_t h, int64_t w) {
      const int64_t o = (int64_t)b * L + keep[slot];
      pos[o] = t;
      pos[(int64_t)B * L + o] = h;
      pos[(int64_t)2 * B * L + o] = w;
      mx = std::max(mx, std::max(t, std::max(h, w)));
    };
    for (int k = 0; k < n_img_row; ++k) {
      if (img >= n_images) return fail(KOCR_ERR_INVALID, "mrope_position_ids: more images in input_ids than rows in image_grid_thw");
      size_t ed = st;
      while (ed < n && tok[ed] != image_token) ++ed;
      if (ed >= n) return fail(KOCR_ERR_INVALID, "mrope_position_ids: image start without image tokens");
      const int64_t t = grid[3 * img], gh = grid[3 * img + 1] / merge, gw = grid[3 * img + 2] / merge;
      if (t != 1) return fail(KOCR_ERR_UNSUPPORTED, "mrope_position_ids: only still images (t == 1) are supported on this path");
      if (ed + (size_t)(gh * gw) > n) return fail(KOCR_ERR_INVALID, "mrope_position_ids: fewer placeholder tokens than the image grid needs");
      for (size_t j = st; j < ed; ++j) put(j, next + (int64_t)(j - st), next + (int64_t)(j - st), next + (int64_t)(j - st));
      const int64_t base = next + (int64_t)(ed - st);
      for (int64_t j = 0; j < gh * gw; ++j) put(ed + (size_t)j, base, base + j / gw, base + j % gw);
      next = mx + 1;
      st = ed + (size_t)(gh * gw);
      ++img;
    }
    for (size_t j = st; j < n; ++j) put(j, next + (int64_t)(j - st), next + (int64_t)(j - st), next + (int64_t)(j - st));
    deltas[b] = mx + 1 - (int64_t)L;
  }
  return KOCR_OK;
}

// one warp per image-embedding row: dst[pos[k], :] = src[k, :] with 16-byte accesses
__global__ void __launch_bounds__(256) scatter_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                                                           const int32_t* __restrict__ pos, int64_t n_rows, int vec_per_row) {
  const int lane = threadIdx.x & 31;
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n_rows; k += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const uint4* s = src + k * vec_per_row;
    uint4* d = dst + (int64_t)pos[k] * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) d[v] = __ldg(s + v);
  }
}

}  // namespace kocr

using namespace kocr;

extern "C" int kocr_mrope_position_ids(const int64_t* input_ids, const int64_t* attention_mask, int batch, int seq_len,
                                       const int64_t* image_grid_thw, int n_images, int64_t image_token_id, int merge,
                                       int64_t* position_ids, int64_t* deltas) {
  if (!input_ids || !position_ids || !deltas || batch <= 0 || seq_len <= 0 || merge <= 0 || (n_images > 0 && !image_grid_thw))
    return fail(KOCR_ERR_INVALID, "kocr_mrope_position_ids: bad argument");
  return mrope_position_ids(input_ids, attention_mask, batch, seq_len, image_grid_thw, n_images, image_token_id, merge, position_ids,
                            deltas);
}

extern "C" int kocr_mrope_position_ids_v2(const int64_t* input_ids, const int64_t* attention_mask, int batch, int seq_len,
                                          const int64_t* image_grid_thw, int n_images, int64_t image_token_id,
                                          int64_t vision_start_token_id, int merge, int semantics, int64_t* position_ids,
                                          int64_t* deltas) {
  if (!input_ids || !position_ids || !deltas || batch <= 0 || seq_len <= 0 || merge <= 0 || (n_images > 0 && !image_grid_thw))
    return fail(KOCR_ERR_INVALID, "kocr_mrope_position_ids_v2: bad argument");
  if (semantics == KOCR_MROPE_TRANSFORMERS_5)
    return mrope_position_ids(input_ids, attention_mask, batch, seq_len, image_grid_thw, n_images, image_token_id, merge, position_ids,
                              deltas);
  if (semantics == KOCR_MROPE_TRANSFORMERS_4_5)
    return mrope_position_ids_v45(input_ids, attention_mask, batch, seq_len, image_grid_thw, n_images, image_token_id,
                                  vision_start_token_id, merge, position_ids, deltas);
  return fail(KOCR_ERR_INVALID, "kocr_mrope_position_ids_v2: unknown semantics");
}

extern "C" int kocr_scatter_image_embeds(KocrCtx* ctx_, void* inputs_embeds, const void* image_embeds, int64_t n_rows, int hidden,
                                         const int64_t* input_ids, int batch, int seq_len, int64_t image_token_id, void* stream_) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  reset_launch_count();
  if (!ctx || !inputs_embeds || !image_embeds || !input_ids || batch <= 0 || seq_len <= 0 || hidden <= 0)
    return fail(KOCR_ERR_INVALID, "kocr_scatter_image_embeds: bad argument");
  if (hidden % 8) return fail(KOCR_ERR_UNSUPPORTED, "kocr_scatter_image_embeds: hidden must be a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(inputs_embeds) | reinterpret_cast<uintptr_t>(image_embeds)) & 15)
    return fail(KOCR_ERR_UNSUPPORTED, "kocr_scatter_image_embeds: tensors must be 16-byte aligned");
  std::vector<int32_t> pos;
  const int64_t total = (int64_t)batch * seq_len;
  if (total > INT32_MAX) return fail(KOCR_ERR_UNSUPPORTED, "kocr_scatter_image_embeds: more than 2^31 tokens");
  for (int64_t i = 0; i < total; ++i)
    if (input_ids[i] == image_token_id) pos.push_back((int32_t)i);
  if ((int64_t)pos.size() != n_rows) {
    char buf[160];
    snprintf(buf, sizeof buf, "Image features and image tokens do not match, tokens: %lld, features: %lld", (long long)pos.size(),
             (long long)n_rows);
    return fail(KOCR_ERR_INVALID, buf);
  }
  if (n_rows == 0) return KOCR_OK;
  void* d_pos;
  int slot = -1;
  int rc = ctx->stage(pos.data(), pos.size() * 4, stream, &d_pos, &slot);
  if (rc) return rc;
  StageGuard guard(ctx, slot, stream);  // the slot is reusable only after scatter_rows_kernel has read it
  ProfScope ps(ctx, kProfOther, stream);
  const unsigned grid = (unsigned)std::min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->num_sms * 16);
  scatter_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<uint4*>(inputs_embeds), static_cast<const uint4*>(image_embeds),
                                                static_cast<const int32_t*>(d_pos), n_rows, hidden / 8);
  KOCR_LAUNCH_CHECK("scatter_rows_kernel");
  return KOCR_OK;
}
