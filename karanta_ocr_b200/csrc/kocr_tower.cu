// Vision tower orchestration: weight prepack + the forward schedule of kernels.
// Stands in for Qwen2VisionTransformerPretrainedModel (HF models/qwen2_vl/modeling_qwen2_vl.py:687-795) and
// Qwen2_5_VisionTransformerPretrainedModel (HF models/qwen2_5_vl/modeling_qwen2_5_vl.py:345-518).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "kocr_common.cuh"
#include "kocr_kernels.h"

namespace kocr {

int build_attn_work(const int32_t* cu, int n_seqs, std::vector<AttnWork>* out);
int build_attn_work3(const int32_t* cu, int n_seqs, std::vector<AttnWork>* out);
int build_attn_work_windowed(const int32_t* cu, int n_seqs, const int32_t* cu_win, int n_win, std::vector<AttnWork>* out,
                             std::vector<int32_t>* row_win);

struct BlockW {
  float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;
  __nv_bfloat16 *w_qkv = nullptr, *w_proj = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;  // fc1 = gate|up interleaved, fc2 = down (2.5)
  float *b_qkv = nullptr, *b_proj = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *c1_qkv = nullptr, *c1_fc1 = nullptr;  // row sums of the gamma-scaled weights (norm1 -> qkv, norm2 -> fc1 folding)
};

// Per-grid plan ("kocr_plan"): everything kocr_tower_forward derives from grid_thw alone - patch
// positions, the attention work lists, the window permutation and per-row window bounds - built once on the host, kept in
// HBM and reused by every later forward with the same grid_thw (a serving loop sees the same page shape again and again).
struct Plan {
  std::vector<int64_t> key;  // grid_thw, flattened
  uint8_t* d = nullptr;      // device tables
  size_t cap = 0;            // bytes allocated at d
  size_t off_pos = 0, off_wf = 0, off_w3 = 0, off_ww = 0, off_wi = 0, off_rw = 0;
  int S = 0, max_pos = 0, n_work_full = 0, n_work3 = 0, n_work_win = 0;  // full attention: n_work3 three-tile blocks + n_work_full two-tile blocks
  cudaEvent_t used = nullptr;       // recorded after the last forward that reads the tables was enqueued
  cudaStream_t last_stream = nullptr;
  bool multi_stream = false;
  bool pinned = false;              // a captured CUDA graph reads the tables: never evicted or rebuilt
  uint64_t tick = 0;
};

struct Tower {
  Ctx* ctx = nullptr;
  static constexpr int kMaxPlans = 16;
  std::vector<Plan*> plans;
  std::mutex plan_mu;
  uint64_t plan_tick = 0;
  int64_t plan_hits = 0, plan_misses = 0;
  KocrTowerConfig cfg{};
  int D = 0, H = 0, F = 0, Fp = 0, O = 0, PD = 0, m2 = 4;
  bool q25 = false;
  __nv_bfloat16* w_patch = nullptr;
  std::vector<BlockW> blk;
  float *ln_w = nullptr, *ln_b = nullptr;
  __nv_bfloat16 *w_m0 = nullptr, *w_m2 = nullptr;
  float *b_m0 = nullptr, *b_m2 = nullptr;
  float2* rope_cs = nullptr;
  int rope_max_pos = 0;
  int32_t* d_qkv_perm = nullptr;
  int32_t* d_gu_perm = nullptr;
  int32_t* d_id_perm = nullptr;  // identity over D rows (down_proj is only re-pitched)
  std::set<std::string> loaded;
  bool folded = false;  // norm1/norm2 have been folded into the qkv / fc1 weights (kocr_tower_finalize)
  int stat_slots = 0;   // partial-statistics slots per row = 2 * ceil(D / 256)
  std::vector<void*> allocs;
  void* tmp = nullptr;  // weight-conversion scratch (set_weight), released by finalize
  size_t tmp_cap = 0;

  template <typename T>
  int alloc(T** p, size_t n, bool zero = false) {
    void* q = nullptr;
    KOCR_CUDA_CHECK(cudaMalloc(&q, n * sizeof(T)));
    if (zero) KOCR_CUDA_CHECK(cudaMemset(q, 0, n * sizeof(T)));
    allocs.push_back(q);
    *p = static_cast<T*>(q);
    return KOCR_OK;
  }
};

static std::vector<std::string> expected_keys(const Tower& t) {
  std::vector<std::string> k = {"patch_embed.proj.weight"};
  for (int i = 0; i < t.cfg.depth; ++i) {
    const std::string p = "blocks." + std::to_string(i) + ".";
    k.push_back(p + "norm1.weight");
    k.push_back(p + "norm2.weight");
    if (!t.q25) {
      k.push_back(p + "norm1.bias");
      k.push_back(p + "norm2.bias");
    }
    for (const char* s : {"attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias"}) k.push_back(p + s);
    if (!t.q25) {
      for (const char* s : {"mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"}) k.push_back(p + s);
    } else {
      for (const char* s : {"mlp.gate_proj.weight", "mlp.gate_proj.bias", "mlp.up_proj.weight", "mlp.up_proj.bias",
                            "mlp.down_proj.weight", "mlp.down_proj.bias"})
        k.push_back(p + s);
    }
  }
  k.push_back("merger.ln_q.weight");
  if (!t.q25) k.push_back("merger.ln_q.bias");
  for (const char* s : {"merger.mlp.0.weight", "merger.mlp.0.bias", "merger.mlp.2.weight", "merger.mlp.2.bias"}) k.push_back(s);
  return k;
}

static int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct WsLayout {
  size_t pv = 0, x = 0, xn = 0, qkv = 0, attn = 0, h = 0, x2 = 0, st_a = 0, st_b = 0, st_tmp = 0, total = 0;
};

static WsLayout ws_layout(const Tower& t, int64_t S, bool need_pv) {
  WsLayout w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
  w.pv = take(need_pv ? (size_t)S * t.PD * 2 : 0);
  w.x = take((size_t)S * t.D * 2);
  w.xn = take((size_t)S * t.D * 2);
  w.qkv = take((size_t)S * 3 * t.D * 2);
  w.attn = take((size_t)S * t.D * 2);
  w.h = take((size_t)S * std::max(t.Fp, t.O) * 2);
  w.x2 = take(t.q25 ? (size_t)S * t.D * 2 : 0);
  w.st_a = take((size_t)S * t.stat_slots * sizeof(float2));
  w.st_b = take((size_t)S * t.stat_slots * sizeof(float2));
  w.st_tmp = take(t.q25 ? (size_t)S * t.stat_slots * sizeof(float2) : 0);
  w.total = o + 256;
  return w;
}

// Look the plan up by grid_thw or build it: host planning + one H2D copy through the pinned staging ring.
static int get_plan(Tower* t, const int64_t* grid_thw, int n_images, cudaStream_t st, Plan** out) {
  Ctx* ctx = t->ctx;
  std::lock_guard<std::mutex> lk(t->plan_mu);
  const size_t nk = (size_t)n_images * 3;
  for (Plan* p : t->plans)
    if (p->key.size() == nk && memcmp(p->key.data(), grid_thw, nk * sizeof(int64_t)) == 0) {
      p->tick = ++t->plan_tick;
      ++t->plan_hits;
      *out = p;
      return KOCR_OK;
    }
  ++t->plan_misses;
  int64_t S64 = 0;
  int max_pos = 0, n_seq_max = 0;
  for (int i = 0; i < n_images; ++i) {
    const int64_t tt = grid_thw[3 * i], h = grid_thw[3 * i + 1], w = grid_thw[3 * i + 2];
    if (tt <= 0 || h <= 0 || w <= 0 || h % 2 || w % 2) return fail(KOCR_ERR_INVALID, "kocr_tower_forward: bad grid_thw");
    S64 += tt * h * w;
    max_pos = std::max<int>(max_pos, (int)std::max(h, w));
    n_seq_max += (int)tt;
  }
  if (S64 > (int64_t)1 << 30) return fail(KOCR_ERR_UNSUPPORTED, "kocr_tower_forward: batch too large, split it");
  const int S = (int)S64;
  int rc;
  std::vector<int32_t> pos((size_t)S * 2), cu(n_seq_max + 1);
  int n_cu = 0;
  if ((rc = pos_ids(grid_thw, n_images, 2, pos.data()))) return rc;
  if ((rc = cu_seqlens(grid_thw, n_images, cu.data(), &n_cu))) return rc;
  std::vector<int32_t> widx, cuw;
  int n_cuw = 0;
  if (t->q25) {
    widx.resize(S / 4);
    cuw.resize(S / 4 + 2);
    if ((rc = window_index(grid_thw, n_images, t->cfg.window_size, 2, t->cfg.patch_size, widx.data(), cuw.data(), &n_cuw)))
      return rc;
    std::vector<int32_t> p2(pos.size());
    for (int g = 0; g < S / 4; ++g) memcpy(&p2[(size_t)g * 8], &pos[(size_t)widx[g] * 8], 8 * sizeof(int32_t));
    pos.swap(p2);
  }
  std::vector<AttnWork> work_full, work3, work_win;
  if ((rc = build_attn_work3(cu.data(), n_cu - 1, &work3))) return rc;  // work_full (two-tile blocks) stays empty: see build_attn_work3
  std::vector<int32_t> row_win;
  if (t->q25 && (rc = build_attn_work_windowed(cu.data(), n_cu - 1, cuw.data(), n_cuw - 1, &work_win, &row_win))) return rc;

  const size_t off_pos = 0;
  const size_t off_wf = align256(pos.size() * 4);
  const size_t off_w3 = off_wf + align256(work_full.size() * sizeof(AttnWork));
  const size_t off_ww = off_w3 + align256(work3.size() * sizeof(AttnWork));
  const size_t off_wi = off_ww + align256(work_win.size() * sizeof(AttnWork));
  const size_t off_rw = off_wi + align256(widx.size() * 4);
  const size_t total = off_rw + align256(row_win.size() * 4);

  // a plan object: a fresh one while the cache has room, else the least recently used one (its buffer is reused when it
  // is large enough, so a stream of ever-changing batches settles into no allocation at all)
  Plan* p = nullptr;
  if ((int)t->plans.size() < Tower::kMaxPlans) {
    p = new Plan();
    KOCR_CUDA_CHECK(cudaEventCreateWithFlags(&p->used, cudaEventDisableTiming));
    t->plans.push_back(p);
  } else {
    p = nullptr;
    for (Plan* q : t->plans)
      if (!q->pinned && (!p || q->tick < p->tick)) p = q;
    if (!p) return fail(KOCR_ERR_STATE, "kocr_tower_forward: every cached plan belongs to a captured CUDA graph; destroy the tower to release them");
    if (p->multi_stream) KOCR_CUDA_CHECK(cudaDeviceSynchronize());
    else KOCR_CUDA_CHECK(cudaEventSynchronize(p->used));  // ~16 forwards old: normally long complete
  }
  p->key.clear();  // invalid until the tables are in place
  if (p->cap < total) {
    if (p->d) cudaFree(p->d);
    p->d = nullptr;
    p->cap = 0;
    const size_t cap = std::max<size_t>(align256(total + total / 4), 1 << 16);
    KOCR_CUDA_CHECK(cudaMalloc(&p->d, cap));
    p->cap = cap;
  }
  void* hs;
  int slot;
  if ((rc = ctx->stage_begin(total, &hs, &slot))) return rc;
  uint8_t* hb = static_cast<uint8_t*>(hs);
  memcpy(hb + off_pos, pos.data(), pos.size() * 4);
  if (!work_full.empty()) memcpy(hb + off_wf, work_full.data(), work_full.size() * sizeof(AttnWork));
  if (!work3.empty()) memcpy(hb + off_w3, work3.data(), work3.size() * sizeof(AttnWork));
  if (!work_win.empty()) memcpy(hb + off_ww, work_win.data(), work_win.size() * sizeof(AttnWork));
  if (!widx.empty()) memcpy(hb + off_wi, widx.data(), widx.size() * 4);
  if (!row_win.empty()) memcpy(hb + off_rw, row_win.data(), row_win.size() * 4);
  cudaError_t e = cudaMemcpyAsync(p->d, hs, total, cudaMemcpyHostToDevice, st);
  ctx->stage_release(slot, st);  // the pinned slot's only reader is this copy
  if (e != cudaSuccess) return fail(KOCR_ERR_CUDA, cudaGetErrorString(e));
  p->off_pos = off_pos; p->off_wf = off_wf; p->off_w3 = off_w3; p->off_ww = off_ww; p->off_wi = off_wi; p->off_rw = off_rw;
  p->S = S; p->max_pos = max_pos; p->n_work_full = (int)work_full.size(); p->n_work3 = (int)work3.size(); p->n_work_win = (int)work_win.size();
  p->last_stream = st;
  p->multi_stream = false;
  p->tick = ++t->plan_tick;
  p->key.assign(grid_thw, grid_thw + nk);
  // other streams must not read the tables before the copy lands: they wait on `used`, which is (re)recorded here
  cudaEventRecord(p->used, st);
  *out = p;
  return KOCR_OK;
}

}  // namespace kocr

using namespace kocr;

extern "C" {

int kocr_tower_create(KocrCtx* ctx_, const KocrTowerConfig* cfg, KocrTower** out) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !cfg || !out) return fail(KOCR_ERR_INVALID, "kocr_tower_create: null argument");
  if (cfg->arch != KOCR_ARCH_QWEN2_VL && cfg->arch != KOCR_ARCH_QWEN2_5_VL) return fail(KOCR_ERR_INVALID, "unknown arch");
  if (cfg->depth <= 0 || cfg->num_heads <= 0 || cfg->embed_dim % cfg->num_heads)
    return fail(KOCR_ERR_INVALID, "kocr_tower_create: bad depth / heads / embed_dim");
  if (cfg->embed_dim / cfg->num_heads != 80)
    return fail(KOCR_ERR_UNSUPPORTED, "kocr_tower_create: kernels are built for head_dim 80 (Qwen2-VL / Qwen2.5-VL towers)");
  if (cfg->spatial_merge_size != 2) return fail(KOCR_ERR_UNSUPPORTED, "kocr_tower_create: spatial_merge_size must be 2");
  if (cfg->embed_dim % 32 || cfg->out_hidden % 32) return fail(KOCR_ERR_UNSUPPORTED, "embed_dim and out_hidden must be multiples of 32");
  if (cfg->n_fullatt < 0 || cfg->n_fullatt > 8) return fail(KOCR_ERR_INVALID, "n_fullatt out of range");
  KOCR_CUDA_CHECK(cudaSetDevice(ctx->device));
  Tower* t = new Tower();
  t->ctx = ctx;
  t->cfg = *cfg;
  t->q25 = cfg->arch == KOCR_ARCH_QWEN2_5_VL;
  t->D = cfg->embed_dim;
  t->H = cfg->num_heads;
  t->F = cfg->mlp_hidden;
  t->Fp = (cfg->mlp_hidden + 31) / 32 * 32;
  t->O = cfg->out_hidden;
  t->PD = cfg->in_channels * cfg->temporal_patch_size * cfg->patch_size * cfg->patch_size;
  t->stat_slots = 2 * ((cfg->embed_dim + 255) / 256);
  if (t->PD % 8) { delete t; return fail(KOCR_ERR_UNSUPPORTED, "patch dim must be a multiple of 8"); }
  const int D = t->D, Fp = t->Fp, O = t->O;
  int rc = t->alloc(&t->w_patch, (size_t)D * t->PD);
  t->blk.resize(cfg->depth);
  for (int i = 0; i < cfg->depth && !rc; ++i) {
    BlockW& b = t->blk[i];
    rc = t->alloc(&b.n1w, D) || t->alloc(&b.n2w, D) || t->alloc(&b.n1b, D, true) || t->alloc(&b.n2b, D, true) ||
         t->alloc(&b.w_qkv, (size_t)3 * D * D) || t->alloc(&b.b_qkv, 3 * D) || t->alloc(&b.w_proj, (size_t)D * D) ||
         t->alloc(&b.b_proj, D) || t->alloc(&b.c1_qkv, 3 * D, true) ||
         t->alloc(&b.c1_fc1, (size_t)(cfg->arch == KOCR_ARCH_QWEN2_5_VL ? 2 : 1) * ((cfg->mlp_hidden + 31) / 32 * 32), true);
    if (rc) break;
    if (!t->q25) {
      rc = t->alloc(&b.w_fc1, (size_t)Fp * D, true) || t->alloc(&b.b_fc1, Fp, true) ||
           t->alloc(&b.w_fc2, (size_t)D * Fp, true) || t->alloc(&b.b_fc2, D);
    } else {
      rc = t->alloc(&b.w_fc1, (size_t)2 * Fp * D, true) || t->alloc(&b.b_fc1, 2 * Fp, true) ||
           t->alloc(&b.w_fc2, (size_t)D * Fp, true) || t->alloc(&b.b_fc2, D);
    }
  }
  if (!rc)
    rc = t->alloc(&t->ln_w, D) || t->alloc(&t->ln_b, D, true) || t->alloc(&t->w_m0, (size_t)16 * D * D) ||
         t->alloc(&t->b_m0, 4 * D) || t->alloc(&t->w_m2, (size_t)O * 4 * D) || t->alloc(&t->b_m2, O);
  if (!rc) {
    // qkv row permutation: new row h*240 + s*80 + d  <-  HF row s*D + h*80 + d   (HF :401-403 reshape(seq,3,H,hd))
    std::vector<int32_t> perm(3 * D);
    for (int h = 0; h < t->H; ++h)
      for (int s = 0; s < 3; ++s)
        for (int d = 0; d < 80; ++d) perm[h * 240 + s * 80 + d] = s * D + h * 80 + d;
    rc = t->alloc(&t->d_qkv_perm, 3 * D);
    if (!rc) KOCR_CUDA_CHECK(cudaMemcpy(t->d_qkv_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice));
    std::vector<int32_t> gp(Fp);
    for (int j = 0; j < Fp; ++j) gp[j] = j < t->F ? j : -1;
    if (!rc) rc = t->alloc(&t->d_gu_perm, Fp);
    if (!rc) KOCR_CUDA_CHECK(cudaMemcpy(t->d_gu_perm, gp.data(), gp.size() * 4, cudaMemcpyHostToDevice));
    std::vector<int32_t> idp(D);
    for (int r = 0; r < D; ++r) idp[r] = r;
    if (!rc) rc = t->alloc(&t->d_id_perm, D);
    if (!rc) KOCR_CUDA_CHECK(cudaMemcpy(t->d_id_perm, idp.data(), idp.size() * 4, cudaMemcpyHostToDevice));
  }
  if (rc) {
    for (void* p : t->allocs) cudaFree(p);
    delete t;
    return rc < 0 ? rc : KOCR_ERR_CUDA;  // message already set by the failing call
  }
  *out = reinterpret_cast<KocrTower*>(t);
  return KOCR_OK;
}

void kocr_tower_destroy(KocrTower* tower) {
  if (!tower) return;
  Tower* t = reinterpret_cast<Tower*>(tower);
  cudaSetDevice(t->ctx->device);
  cudaDeviceSynchronize();
  for (void* p : t->allocs) cudaFree(p);
  if (t->rope_cs) cudaFree(t->rope_cs);
  if (t->tmp) cudaFree(t->tmp);
  for (Plan* p : t->plans) {
    if (p->d) cudaFree(p->d);
    if (p->used) cudaEventDestroy(p->used);
    delete p;
  }
  delete t;
}

int kocr_tower_set_weight(KocrTower* tower, const char* name_, const void* data, int dtype, const int64_t* shape,
                          int ndim) {
  Tower* t = reinterpret_cast<Tower*>(tower);
  if (!t || !name_ || !data || !shape || ndim <= 0) return fail(KOCR_ERR_INVALID, "kocr_tower_set_weight: null argument");
  if (dtype < KOCR_DTYPE_F32 || dtype > KOCR_DTYPE_F16) return fail(KOCR_ERR_INVALID, "kocr_tower_set_weight: bad dtype");
  KOCR_CUDA_CHECK(cudaSetDevice(t->ctx->device));
  if (t->folded) return fail(KOCR_ERR_STATE, "kocr_tower_set_weight: weights were already finalized (norms folded); create a new tower");
  const std::string name(name_);
  const int64_t n = numel(shape, ndim);
  const int D = t->D, F = t->F, Fp = t->Fp, O = t->O;
  cudaStream_t st = 0;
  auto bad_shape = [&]() { return fail(KOCR_ERR_INVALID, "kocr_tower_set_weight: unexpected shape for " + name); };
  auto plain = [&](void* dst, int ddt, int64_t expect) -> int {
    if (n != expect) return bad_shape();
    return launch_convert(data, dtype, dst, ddt, n, st);
  };
  // convert into a temporary, then permute rows into place
  auto permuted = [&](void* dst, int ddt, int64_t src_rows, int64_t cols, const int32_t* perm, int64_t dst_rows,
                      int64_t dst_ld) -> int {
    if (n != src_rows * cols) return bad_shape();
    const int elt = ddt == KOCR_DTYPE_F32 ? 4 : 2;
    if (t->tmp_cap < (size_t)n * elt) {  // one conversion buffer for the whole load (grown, never per tensor), freed at finalize
      KOCR_CUDA_CHECK(cudaStreamSynchronize(st));
      if (t->tmp) cudaFree(t->tmp);
      t->tmp = nullptr;
      t->tmp_cap = 0;
      KOCR_CUDA_CHECK(cudaMalloc(&t->tmp, (size_t)n * elt));
      t->tmp_cap = (size_t)n * elt;
    }
    int rc = launch_convert(data, dtype, t->tmp, ddt, n, st);
    if (!rc) rc = launch_permute_rows(t->tmp, dst, perm, dst_rows, cols, cols, dst_ld, elt, st);
    return rc;  // stream-ordered: the next tensor's conversion follows this permute on the same stream
  };
  int rc = KOCR_ERR_INVALID;
  if (name == "patch_embed.proj.weight") {
    rc = plain(t->w_patch, KOCR_DTYPE_BF16, (int64_t)D * t->PD);
  } else if (name == "merger.ln_q.weight") {
    rc = plain(t->ln_w, KOCR_DTYPE_F32, D);
  } else if (name == "merger.ln_q.bias" && !t->q25) {
    rc = plain(t->ln_b, KOCR_DTYPE_F32, D);
  } else if (name == "merger.mlp.0.weight") {
    rc = plain(t->w_m0, KOCR_DTYPE_BF16, (int64_t)16 * D * D);
  } else if (name == "merger.mlp.0.bias") {
    rc = plain(t->b_m0, KOCR_DTYPE_F32, 4 * D);
  } else if (name == "merger.mlp.2.weight") {
    rc = plain(t->w_m2, KOCR_DTYPE_BF16, (int64_t)O * 4 * D);
  } else if (name == "merger.mlp.2.bias") {
    rc = plain(t->b_m2, KOCR_DTYPE_F32, O);
  } else if (name.rfind("blocks.", 0) == 0) {
    const size_t dot = name.find('.', 7);
    if (dot == std::string::npos) return fail(KOCR_ERR_INVALID, "unknown weight " + name);
    const int i = atoi(name.substr(7, dot - 7).c_str());
    if (i < 0 || i >= t->cfg.depth) return fail(KOCR_ERR_INVALID, "block index out of range in " + name);
    const std::string k = name.substr(dot + 1);
    BlockW& b = t->blk[i];
    if (k == "norm1.weight") rc = plain(b.n1w, KOCR_DTYPE_F32, D);
    else if (k == "norm1.bias" && !t->q25) rc = plain(b.n1b, KOCR_DTYPE_F32, D);
    else if (k == "norm2.weight") rc = plain(b.n2w, KOCR_DTYPE_F32, D);
    else if (k == "norm2.bias" && !t->q25) rc = plain(b.n2b, KOCR_DTYPE_F32, D);
    else if (k == "attn.qkv.weight") rc = permuted(b.w_qkv, KOCR_DTYPE_BF16, 3 * D, D, t->d_qkv_perm, 3 * D, D);
    else if (k == "attn.qkv.bias") rc = permuted(b.b_qkv, KOCR_DTYPE_F32, 3 * D, 1, t->d_qkv_perm, 3 * D, 1);
    else if (k == "attn.proj.weight") rc = plain(b.w_proj, KOCR_DTYPE_BF16, (int64_t)D * D);
    else if (k == "attn.proj.bias") rc = plain(b.b_proj, KOCR_DTYPE_F32, D);
    else if (!t->q25 && k == "mlp.fc1.weight") rc = plain(b.w_fc1, KOCR_DTYPE_BF16, (int64_t)F * D);
    else if (!t->q25 && k == "mlp.fc1.bias") rc = plain(b.b_fc1, KOCR_DTYPE_F32, F);
    else if (!t->q25 && k == "mlp.fc2.weight") {
      if (F == Fp) rc = plain(b.w_fc2, KOCR_DTYPE_BF16, (int64_t)D * F);
      else rc = KOCR_ERR_UNSUPPORTED;
    } else if (!t->q25 && k == "mlp.fc2.bias") rc = plain(b.b_fc2, KOCR_DTYPE_F32, D);
    // Qwen2.5-VL: gate/up rows interleaved (row 2j = gate_j, 2j+1 = up_j) == [Fp, 2D] with gate | up side by side
    else if (t->q25 && k == "mlp.gate_proj.weight") rc = permuted(b.w_fc1, KOCR_DTYPE_BF16, F, D, t->d_gu_perm, F, 2 * D);
    else if (t->q25 && k == "mlp.up_proj.weight") rc = permuted(b.w_fc1 + D, KOCR_DTYPE_BF16, F, D, t->d_gu_perm, F, 2 * D);
    else if (t->q25 && k == "mlp.gate_proj.bias") rc = permuted(b.b_fc1, KOCR_DTYPE_F32, F, 1, t->d_gu_perm, F, 2);
    else if (t->q25 && k == "mlp.up_proj.bias") rc = permuted(b.b_fc1 + 1, KOCR_DTYPE_F32, F, 1, t->d_gu_perm, F, 2);
    // down_proj [D, F] -> [D, Fp] (zero padded K): identity row "permutation" with a wider destination pitch
    else if (t->q25 && k == "mlp.down_proj.weight") {
      if (n != (int64_t)D * F) return bad_shape();
      rc = permuted(b.w_fc2, KOCR_DTYPE_BF16, D, F, t->d_id_perm, D, Fp);
    } else if (t->q25 && k == "mlp.down_proj.bias") rc = plain(b.b_fc2, KOCR_DTYPE_F32, D);
    else return fail(KOCR_ERR_INVALID, "unknown weight " + name);
  } else {
    return fail(KOCR_ERR_INVALID, "unknown weight " + name);
  }
  if (rc == KOCR_ERR_UNSUPPORTED) return fail(rc, "mlp hidden size must be a multiple of 32 for " + name);
  if (rc) return rc;
  KOCR_CUDA_CHECK(cudaStreamSynchronize(st));
  t->loaded.insert(name);
  return KOCR_OK;
}

int kocr_tower_finalize(KocrTower* tower) {
  Tower* t = reinterpret_cast<Tower*>(tower);
  if (!t) return fail(KOCR_ERR_INVALID, "kocr_tower_finalize: null tower");
  std::string missing;
  int n_missing = 0;
  for (const std::string& k : expected_keys(*t))
    if (!t->loaded.count(k)) {
      if (n_missing++ < 8) missing += (missing.empty() ? "" : ", ") + k;
    }
  if (n_missing) return fail(KOCR_ERR_STATE, "missing " + std::to_string(n_missing) + " weights: " + missing);
  if (!t->folded) {
    // Fold norm1 into qkv and norm2 into fc1 (gate/up): W' = W diag(gamma), b' = b + W beta, c1 = row sums of W'.
    // The GEMM then reads the raw residual stream and applies mean / rstd per row in its epilogue (no norm pass).
    KOCR_CUDA_CHECK(cudaSetDevice(t->ctx->device));
    const int D = t->D;
    for (BlockW& b : t->blk) {
      int rc = launch_fold_norm(b.w_qkv, D, b.b_qkv, b.c1_qkv, b.n1w, t->q25 ? nullptr : b.n1b, 3 * D, D, 0);
      if (!rc) rc = launch_fold_norm(b.w_fc1, D, b.b_fc1, b.c1_fc1, b.n2w, t->q25 ? nullptr : b.n2b, (t->q25 ? 2 : 1) * (int64_t)t->Fp, D, 0);
      if (rc) return rc;
    }
    KOCR_CUDA_CHECK(cudaStreamSynchronize(0));
    if (t->tmp) cudaFree(t->tmp);
    t->tmp = nullptr;
    t->tmp_cap = 0;
    t->folded = true;
  }
  return KOCR_OK;
}

int64_t kocr_tower_workspace_bytes(const KocrTower* tower, const int64_t* grid_thw, int n_images) {
  const Tower* t = reinterpret_cast<const Tower*>(tower);
  if (!t || !grid_thw || n_images <= 0) return fail(KOCR_ERR_INVALID, "kocr_tower_workspace_bytes: bad argument");
  int64_t S = 0;
  for (int i = 0; i < n_images; ++i) S += grid_thw[3 * i] * grid_thw[3 * i + 1] * grid_thw[3 * i + 2];
  return (int64_t)ws_layout(*t, S, true).total;
}

int kocr_tower_plan_stats(const KocrTower* tower, int64_t* hits, int64_t* misses) {
  const Tower* t = reinterpret_cast<const Tower*>(tower);
  if (!t || !hits || !misses) return fail(KOCR_ERR_INVALID, "kocr_tower_plan_stats: null argument");
  *hits = t->plan_hits;
  *misses = t->plan_misses;
  return KOCR_OK;
}

int kocr_tower_forward(KocrTower* tower, const void* pixel_values, int pv_dtype, const int64_t* grid_thw, int n_images,
                       void* out, void* hidden_out, void* workspace, int64_t workspace_bytes, void* stream_) {
  Tower* t = reinterpret_cast<Tower*>(tower);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  reset_launch_count();
  if (!t || !pixel_values || !grid_thw || n_images <= 0 || !out || !workspace)
    return fail(KOCR_ERR_INVALID, "kocr_tower_forward: null argument or empty batch");
  if (pv_dtype != KOCR_DTYPE_F32 && pv_dtype != KOCR_DTYPE_BF16)
    return fail(KOCR_ERR_INVALID, "kocr_tower_forward: pixel_values must be f32 or bf16");
  int rc = kocr_tower_finalize(tower);
  if (rc) return rc;
  Ctx* ctx = t->ctx;
  const int D = t->D, H = t->H, Fp = t->Fp, O = t->O;

  // ---- plan: positions, sequences, windows (cached per grid_thw)
  KOCR_CUDA_CHECK(cudaSetDevice(ctx->device));
  // Under stream capture (CUDA graphs: call the forward once eagerly with this grid_thw first, so that the plan exists and the
  // one-time attribute / table work is done) no event may tie the capture to work outside it: the plan is pinned instead.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  KOCR_CUDA_CHECK(cudaStreamIsCapturing(st, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  Plan* plan = nullptr;
  if (capturing) {
    std::lock_guard<std::mutex> lk(t->plan_mu);
    const size_t nk = (size_t)n_images * 3;
    for (Plan* p : t->plans)
      if (p->key.size() == nk && memcmp(p->key.data(), grid_thw, nk * sizeof(int64_t)) == 0) plan = p;
    if (!plan) return fail(KOCR_ERR_STATE, "kocr_tower_forward: capture needs a cached plan - run one eager forward with this grid_thw first");
    if (plan->max_pos > t->rope_max_pos) return fail(KOCR_ERR_STATE, "kocr_tower_forward: capture needs the rotary table in place - run one eager forward first");
    plan->pinned = true;
  } else {
    if ((rc = get_plan(t, grid_thw, n_images, st, &plan))) return rc;
    if (plan->last_stream != st) {  // built or last used on another stream: order this stream after it
      KOCR_CUDA_CHECK(cudaStreamWaitEvent(st, plan->used, 0));
      plan->multi_stream = true;
      plan->last_stream = st;
    }
  }
  struct PlanUse {  // the tables stay alive (not evicted / overwritten) until every forward that reads them has run
    Plan* p; cudaStream_t s; bool on;
    ~PlanUse() { if (on) cudaEventRecord(p->used, s); }
  } plan_use{plan, st, !capturing};
  const int S = plan->S, max_pos = plan->max_pos;
  const WsLayout wl = ws_layout(*t, S, pv_dtype == KOCR_DTYPE_F32);
  if ((int64_t)wl.total > workspace_bytes) return fail(KOCR_ERR_INVALID, "kocr_tower_forward: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  auto* x = reinterpret_cast<__nv_bfloat16*>(ws + wl.x);
  auto* xn = reinterpret_cast<__nv_bfloat16*>(ws + wl.xn);
  auto* qkv = reinterpret_cast<__nv_bfloat16*>(ws + wl.qkv);
  auto* attn = reinterpret_cast<__nv_bfloat16*>(ws + wl.attn);
  auto* hbuf = reinterpret_cast<__nv_bfloat16*>(ws + wl.h);
  auto* x2 = reinterpret_cast<__nv_bfloat16*>(ws + wl.x2);
  auto* st_a = reinterpret_cast<float2*>(ws + wl.st_a);   // row statistics of x after the attention residual (-> norm2)
  auto* st_b = reinterpret_cast<float2*>(ws + wl.st_b);   // row statistics of x after the MLP residual / patch embed (-> norm1)
  auto* st_tmp = reinterpret_cast<float2*>(ws + wl.st_tmp);
  const int slots = t->stat_slots;

  if (max_pos > t->rope_max_pos) {  // grow the (cos, sin) table; rare. Forwards on other streams may still read the old
    KOCR_CUDA_CHECK(cudaDeviceSynchronize());  // table, so the whole device is drained before it is freed
    if (t->rope_cs) cudaFree(t->rope_cs);
    t->rope_cs = nullptr;
    const int grown = std::max(max_pos, 256);
    KOCR_CUDA_CHECK(cudaMalloc(&t->rope_cs, (size_t)grown * 20 * sizeof(float2)));
    rc = launch_rope_table(t->rope_cs, grown, 20, 10000.0f, st);
    if (rc) return rc;
    KOCR_CUDA_CHECK(cudaStreamSynchronize(st));  // complete before any other stream's forward can pick it up
    t->rope_max_pos = grown;
  }

  const uint8_t* db = plan->d;
  const int2* d_pos = reinterpret_cast<const int2*>(db + plan->off_pos);
  const AttnWork* d_wf = reinterpret_cast<const AttnWork*>(db + plan->off_wf);
  const AttnWork* d_w3 = reinterpret_cast<const AttnWork*>(db + plan->off_w3);
  const AttnWork* d_ww = reinterpret_cast<const AttnWork*>(db + plan->off_ww);
  const int32_t* d_wi = reinterpret_cast<const int32_t*>(db + plan->off_wi);
  const int2* d_rw = reinterpret_cast<const int2*>(db + plan->off_rw);
  const int n_work_full = plan->n_work_full, n_work3 = plan->n_work3, n_work_win = plan->n_work_win;

  // ---- patch embed (HF :304-310; Conv3d with kernel == stride is a GEMM over the flattened patch)
  const void* pv = pixel_values;
  if (pv_dtype == KOCR_DTYPE_F32) {
    ProfScope ps(ctx, kProfOther, st);
    if ((rc = launch_cast_f32_bf16(static_cast<const float*>(pixel_values), ws + wl.pv, (int64_t)S * t->PD, st))) return rc;
    pv = ws + wl.pv;
  }
  GemmEpilogue ep{};
  __nv_bfloat16* x_embed = t->q25 ? x2 : x;
  ep.out = x_embed;
  ep.ldc = D;
  ep.stat_part = t->q25 ? st_tmp : st_b;
  ep.stat_slots = slots;
  {
    ProfScope ps(ctx, kProfPatchEmbed, st);
    if ((rc = launch_gemm(ctx, pv, t->PD, t->w_patch, t->PD, S, D, t->PD, KOCR_EPI_NONE, ep, st))) return rc;
  }
  if (t->q25) {  // window-major permutation of 4-patch groups (HF qwen2_5 :478-481)
    ProfScope ps(ctx, kProfOther, st);
    if ((rc = launch_gather_groups(x2, x, d_wi, S / 4, 4, D, false, st))) return rc;
    if ((rc = launch_gather_groups(st_tmp, st_b, d_wi, S / 4, 4, slots * 4, false, st))) return rc;  // the rows' statistics move with them
  }

  const float q_scale = (float)((1.0 / sqrt(80.0)) * 1.4426950408889634);
  for (int i = 0; i < t->cfg.depth; ++i) {
    const BlockW& b = t->blk[i];
    bool full = true;
    if (t->q25) {
      full = false;
      for (int k = 0; k < t->cfg.n_fullatt; ++k) full |= t->cfg.fullatt_block_indexes[k] == i;
    }
    // [norm1 folded] qkv (+bias, RoPE, q scale) -> attention -> proj (+bias, +residual, row statistics for norm2)
    GemmEpilogue e1{};
    e1.bias = b.b_qkv; e1.out = qkv; e1.ldc = 3 * D; e1.pos_hw = d_pos; e1.rope_cs = t->rope_cs; e1.q_scale = q_scale;
    e1.ln_part = st_b; e1.ln_c1 = b.c1_qkv; e1.ln_slots = slots; e1.ln_rms = t->q25; e1.ln_inv_dim = 1.0f / D; e1.ln_eps = 1e-6f;
    {
      ProfScope ps(ctx, kProfQkvRope, st);
      if ((rc = launch_gemm(ctx, x, D, b.w_qkv, D, S, 3 * D, D, kEpiQkvRope, e1, st))) return rc;
    }
    {
      ProfScope ps(ctx, full ? kProfAttention : kProfAttentionWin, st);
      if (full) {  // 384-row blocks on the three-tile kernel, what they leave (at most two blocks per sequence) on the two-tile one
        if (n_work3) rc = launch_attention3(ctx, qkv, attn, d_w3, n_work3, H, S, st);
        if (!rc && n_work_full) rc = launch_attention(ctx, qkv, attn, d_wf, n_work_full, H, S, st);
      } else {
        rc = launch_attention(ctx, qkv, attn, d_ww, n_work_win, H, S, st, d_rw);
      }
      if (rc) return rc;
    }
    GemmEpilogue e2{};
    e2.bias = b.b_proj; e2.residual = x; e2.ld_res = D; e2.out = x; e2.ldc = D; e2.stat_part = st_a; e2.stat_slots = slots;
    {
      ProfScope ps(ctx, kProfProj, st);
      if ((rc = launch_gemm(ctx, attn, D, b.w_proj, D, S, D, D, KOCR_EPI_BIAS_RESIDUAL, e2, st))) return rc;
    }
    // [norm2 folded] fc1 (+bias, activation) -> fc2 (+bias, +residual, row statistics for the next block's norm1)
    GemmEpilogue e3{};
    e3.bias = b.b_fc1; e3.out = hbuf; e3.ldc = Fp;
    e3.ln_part = st_a; e3.ln_c1 = b.c1_fc1; e3.ln_slots = slots; e3.ln_rms = t->q25; e3.ln_inv_dim = 1.0f / D; e3.ln_eps = 1e-6f;
    {
      ProfScope ps(ctx, kProfFc1, st);
      if (!t->q25) rc = launch_gemm(ctx, x, D, b.w_fc1, D, S, Fp, D, KOCR_EPI_BIAS_QUICKGELU, e3, st);
      else rc = launch_gemm(ctx, x, D, b.w_fc1, D, S, 2 * Fp, D, KOCR_EPI_BIAS_SWIGLU, e3, st);
      if (rc) return rc;
    }
    GemmEpilogue e4{};
    e4.bias = b.b_fc2; e4.residual = x; e4.ld_res = D; e4.out = x; e4.ldc = D; e4.stat_part = st_b; e4.stat_slots = slots;
    {
      ProfScope ps(ctx, kProfFc2, st);
      if ((rc = launch_gemm(ctx, hbuf, Fp, b.w_fc2, Fp, S, D, Fp, KOCR_EPI_BIAS_RESIDUAL, e4, st))) return rc;
    }
  }

  // ---- merger (HF :313-326): LN -> view(-1, 4D) (free: 4 consecutive tokens are one merge cell) -> GEMM+GELU -> GEMM
  ProfScope ps_merger(ctx, kProfMerger, st);
  if ((rc = launch_norm(x, D, t->ln_w, t->q25 ? nullptr : t->ln_b, xn, D, S, D, 1e-6f, t->q25, st))) return rc;
  GemmEpilogue e5{};
  e5.bias = t->b_m0; e5.out = qkv; e5.ldc = 4 * D;
  if ((rc = launch_gemm(ctx, xn, 4 * D, t->w_m0, 4 * D, S / 4, 4 * D, 4 * D, KOCR_EPI_BIAS_GELU, e5, st))) return rc;
  GemmEpilogue e6{};
  e6.bias = t->b_m2; e6.ldc = O;
  e6.out = t->q25 ? hbuf : static_cast<__nv_bfloat16*>(out);
  if ((rc = launch_gemm(ctx, qkv, 4 * D, t->w_m2, 4 * D, S / 4, O, 4 * D, KOCR_EPI_BIAS, e6, st))) return rc;
  if (t->q25) {  // undo the window permutation (HF qwen2_5 :511-513): out[window_index[g]] = merged[g]
    if ((rc = launch_gather_groups(hbuf, out, d_wi, S / 4, 1, O, true, st))) return rc;
  }
  if (hidden_out) {
    KOCR_CUDA_CHECK(cudaMemcpyAsync(hidden_out, x, (size_t)S * D * 2, cudaMemcpyDeviceToDevice, st));
  }
  return KOCR_OK;
}

}  // extern "C"
