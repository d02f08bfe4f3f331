"""CPU oracle for the Qwen2-VL / Qwen2.5-VL vision tower (TEST INFRASTRUCTURE ONLY).

Plain-torch fp32 restatement of the arithmetic the reference reaches through `transformers`
(pinned 4.53.3, /root/reference/uv.lock:2168-2169; call sites
/root/reference/karanta/training/ocr_training.py:86,670 and
/root/reference/karanta/training/test_trained_model.py:91).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import it; the product path never does.

Upstream lines followed (HF = site-packages/transformers 5.5.0):
  pos_ids / rot_pos_emb   HF models/qwen2_vl/modeling_qwen2_vl.py:725-752
  inv_freq                HF models/qwen2_vl/modeling_qwen2_vl.py:271-284 (dim = head_dim // 2)
  cu_seqlens              HF models/qwen2_vl/modeling_qwen2_vl.py:772-780
  rotate_half / rope      HF models/qwen2_vl/modeling_qwen2_vl.py:205-209,257-268
  PatchEmbed              HF models/qwen2_vl/modeling_qwen2_vl.py:304-310 (Conv3d k=s == GEMM)
  VisionAttention         HF models/qwen2_vl/modeling_qwen2_vl.py:392-458
  VisionMlp / QuickGELU   HF models/qwen2_vl/modeling_qwen2_vl.py:329-337, HF activations.py QuickGELUActivation
  block / merger          HF models/qwen2_vl/modeling_qwen2_vl.py:461-487,313-326
  tower forward           HF models/qwen2_vl/modeling_qwen2_vl.py:757-795
  Qwen2.5-VL deltas       HF models/qwen2_5_vl/modeling_qwen2_5_vl.py:57-88 (RMSNorm, gated MLP),
                          :411-451 (get_window_index), :455-518 (forward)

Parity status: pinned against the third-party implementation run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz); the reference's own tests hold no vector here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class TowerConfig:
    """Field names follow Qwen2VLVisionConfig / Qwen2_5_VLVisionConfig."""
    arch: str = "qwen2_vl"          # or "qwen2_5_vl"
    depth: int = 32
    embed_dim: int = 1280           # qwen2_5_vl calls this hidden_size
    num_heads: int = 16
    mlp_hidden: int = 5120          # embed_dim*mlp_ratio (qwen2_vl) or intermediate_size (qwen2_5_vl)
    out_hidden: int = 3584          # qwen2_vl hidden_size / qwen2_5_vl out_hidden_size
    patch_size: int = 14
    temporal_patch_size: int = 2
    in_channels: int = 3
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: tuple = field(default_factory=lambda: (7, 15, 23, 31))

    @property
    def head_dim(self):
        return self.embed_dim // self.num_heads

    @property
    def patch_dim(self):
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size


def qwen2_vl_7b(depth=32):
    return TowerConfig("qwen2_vl", depth, 1280, 16, 5120, 3584)


def qwen2_vl_2b(depth=32):
    return TowerConfig("qwen2_vl", depth, 1280, 16, 5120, 1536)


def qwen2_5_vl_7b(depth=32):
    return TowerConfig("qwen2_5_vl", depth, 1280, 16, 3420, 3584)


# ---------------------------------------------------------------- integer / index work (bit-exact parity)

def pos_ids(grid_thw, merge: int = 2) -> np.ndarray:
    """[sumN, 2] int32 (row, col) of each patch in merge-major token order (rot_pos_emb :727-748)."""
    out = []
    for t, h, w in np.asarray(grid_thw).tolist():
        hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // merge, merge, w // merge, merge)
        wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // merge, merge, w // merge, merge)
        hp = hp.transpose(0, 2, 1, 3).reshape(-1)
        wp = wp.transpose(0, 2, 1, 3).reshape(-1)
        out.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
    return np.concatenate(out, 0).astype(np.int32)


def cu_seqlens(grid_thw) -> np.ndarray:
    """int32 [sum(t)+1] (forward :772-780)."""
    g = np.asarray(grid_thw, dtype=np.int64)
    lens = np.repeat(g[:, 1] * g[:, 2], g[:, 0])
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


def window_index(grid_thw, window_size=112, merge=2, patch=14):
    """Qwen2.5-VL get_window_index :411-451 -> (window_index int32 [sumN/4], cu_window_seqlens int32 after
    unique_consecutive :476)."""
    win = window_size // merge // patch
    widx, cu = [], [0]
    base = 0
    for t, gh, gw in np.asarray(grid_thw).tolist():
        lh, lw = gh // merge, gw // merge
        index = np.arange(t * lh * lw).reshape(t, lh, lw)
        pad_h = win - lh % win
        pad_w = win - lw % win
        nwh, nww = (lh + pad_h) // win, (lw + pad_w) // win
        ip = np.pad(index, ((0, 0), (0, pad_h), (0, pad_w)), constant_values=-100)
        ip = ip.reshape(t, nwh, win, nww, win).transpose(0, 1, 3, 2, 4).reshape(t, nwh * nww, win, win)
        seqlens = (ip != -100).sum((2, 3)).reshape(-1)
        flat = ip.reshape(-1)
        widx.append(flat[flat != -100] + base)
        cu.extend((np.cumsum(seqlens) * merge * merge + cu[-1]).tolist())
        base += t * lh * lw
    cu = np.asarray(cu, dtype=np.int64)
    keep = np.concatenate([[True], cu[1:] != cu[:-1]])
    return np.concatenate(widx).astype(np.int32), cu[keep].astype(np.int32)


def inv_freq(head_dim: int, theta: float = 10000.0) -> torch.Tensor:
    dim = head_dim // 2
    return 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))


def rope_cos_sin(grid_thw, head_dim: int, merge: int = 2):
    """f32 cos/sin [sumN, head_dim]: angle vector = cat(row*f, col*f, row*f, col*f) (forward :768-770)."""
    p = torch.from_numpy(pos_ids(grid_thw, merge).astype(np.int64))
    g = np.asarray(grid_thw)
    seq = torch.arange(int(g[:, 1:].max()), dtype=torch.float)
    freqs = torch.outer(seq, inv_freq(head_dim))
    rot = freqs[p].flatten(1)
    emb = torch.cat((rot, rot), dim=-1)
    return emb.cos(), emb.sin()


# ---------------------------------------------------------------- weights

def init_weights(cfg: TowerConfig, seed: int = 0, std: float = 0.02) -> dict:
    """Seeded random weights under HF state-dict key names (no checkpoints exist offline).
    Generated one tensor at a time from a single torch CPU generator, so the same seed gives the
    same tensors on any box with this torch build. Norm weights are perturbed around 1 and biases are
    non-zero so every term of every epilogue is exercised."""
    g = torch.Generator().manual_seed(seed)
    D, Fh, O = cfg.embed_dim, cfg.mlp_hidden, cfg.out_hidden

    def rn(*shape, s=std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * s

    sd = {"patch_embed.proj.weight": rn(D, cfg.in_channels, cfg.temporal_patch_size, cfg.patch_size, cfg.patch_size)}
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + rn(D, s=0.1)
        sd[p + "norm2.weight"] = 1.0 + rn(D, s=0.1)
        if cfg.arch == "qwen2_vl":
            sd[p + "norm1.bias"] = rn(D, s=0.1)
            sd[p + "norm2.bias"] = rn(D, s=0.1)
        sd[p + "attn.qkv.weight"] = rn(3 * D, D)
        sd[p + "attn.qkv.bias"] = rn(3 * D, s=0.1)
        sd[p + "attn.proj.weight"] = rn(D, D)
        sd[p + "attn.proj.bias"] = rn(D, s=0.1)
        if cfg.arch == "qwen2_vl":
            sd[p + "mlp.fc1.weight"] = rn(Fh, D)
            sd[p + "mlp.fc1.bias"] = rn(Fh, s=0.1)
            sd[p + "mlp.fc2.weight"] = rn(D, Fh)
            sd[p + "mlp.fc2.bias"] = rn(D, s=0.1)
        else:
            for nm, shp in (("gate_proj", (Fh, D)), ("up_proj", (Fh, D)), ("down_proj", (D, Fh))):
                sd[p + f"mlp.{nm}.weight"] = rn(*shp)
                sd[p + f"mlp.{nm}.bias"] = rn(shp[0], s=0.1)
    sd["merger.ln_q.weight"] = 1.0 + rn(D, s=0.1)
    if cfg.arch == "qwen2_vl":
        sd["merger.ln_q.bias"] = rn(D, s=0.1)
    m = cfg.spatial_merge_size ** 2
    sd["merger.mlp.0.weight"] = rn(D * m, D * m)
    sd["merger.mlp.0.bias"] = rn(D * m, s=0.1)
    sd["merger.mlp.2.weight"] = rn(O, D * m)
    sd["merger.mlp.2.bias"] = rn(O, s=0.1)
    return sd


# ---------------------------------------------------------------- fp32 forward

def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def _rms_norm(x, w, eps=1e-6):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def _attention(q, k, v, cu, scale):
    """q,k,v [S, H, hd]; block-diagonal non-causal attention per cu_seqlens segment."""
    out = torch.empty_like(q)
    cu = [int(c) for c in cu]
    for a, b in zip(cu[:-1], cu[1:]):
        qs, ks, vs = (t[a:b].transpose(0, 1).unsqueeze(0) for t in (q, k, v))
        o = F.scaled_dot_product_attention(qs, ks, vs, scale=scale, is_causal=False)
        out[a:b] = o.squeeze(0).transpose(0, 1)
    return out


@torch.no_grad()
def tower_forward(cfg: TowerConfig, sd: dict, pixel_values: torch.Tensor, grid_thw, dtype=torch.float32,
                  return_hidden: bool = False):
    """pixel_values [..., patch_dim] -> merged embeddings [sumN/4, out_hidden] in `dtype`."""
    g = np.asarray(grid_thw, dtype=np.int64)
    w = {k: v.to(dtype) for k, v in sd.items()}
    D, H, hd = cfg.embed_dim, cfg.num_heads, cfg.head_dim
    x = pixel_values.reshape(-1, cfg.patch_dim).to(dtype)
    x = F.linear(x, w["patch_embed.proj.weight"].reshape(D, -1))
    S = x.shape[0]
    cos, sin = rope_cos_sin(g, hd, cfg.spatial_merge_size)
    cu_full = cu_seqlens(g)
    m2 = cfg.spatial_merge_size ** 2
    if cfg.arch == "qwen2_5_vl":
        widx, cu_win = window_index(g, cfg.window_size, cfg.spatial_merge_size, cfg.patch_size)
        wi = torch.from_numpy(widx.astype(np.int64))
        x = x.reshape(S // m2, m2, -1)[wi].reshape(S, -1)
        cos = cos.reshape(S // m2, m2, -1)[wi].reshape(S, -1)
        sin = sin.reshape(S // m2, m2, -1)[wi].reshape(S, -1)
    cos_, sin_ = cos.unsqueeze(-2).float(), sin.unsqueeze(-2).float()
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        if cfg.arch == "qwen2_vl":
            h = F.layer_norm(x, (D,), w[p + "norm1.weight"], w[p + "norm1.bias"], 1e-6)
        else:
            h = _rms_norm(x, w[p + "norm1.weight"])
        qkv = F.linear(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"]).reshape(S, 3, H, hd)
        q, k, v = qkv.unbind(1)
        qf, kf = q.float(), k.float()
        q = (qf * cos_ + _rotate_half(qf) * sin_).to(dtype)
        k = (kf * cos_ + _rotate_half(kf) * sin_).to(dtype)
        cu = cu_full
        if cfg.arch == "qwen2_5_vl" and i not in cfg.fullatt_block_indexes:
            cu = cu_win
        a = _attention(q, k, v, cu, hd ** -0.5).reshape(S, D)
        x = x + F.linear(a, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
        if cfg.arch == "qwen2_vl":
            h = F.layer_norm(x, (D,), w[p + "norm2.weight"], w[p + "norm2.bias"], 1e-6)
            h = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
            h = h * torch.sigmoid(1.702 * h)
            h = F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
        else:
            h = _rms_norm(x, w[p + "norm2.weight"])
            gate = F.linear(h, w[p + "mlp.gate_proj.weight"], w[p + "mlp.gate_proj.bias"])
            up = F.linear(h, w[p + "mlp.up_proj.weight"], w[p + "mlp.up_proj.bias"])
            h = F.linear(F.silu(gate) * up, w[p + "mlp.down_proj.weight"], w[p + "mlp.down_proj.bias"])
        x = x + h
    if cfg.arch == "qwen2_vl":
        h = F.layer_norm(x, (D,), w["merger.ln_q.weight"], w["merger.ln_q.bias"], 1e-6)
    else:
        h = _rms_norm(x, w["merger.ln_q.weight"])
    h = h.reshape(-1, D * m2)
    h = F.gelu(F.linear(h, w["merger.mlp.0.weight"], w["merger.mlp.0.bias"]))
    out = F.linear(h, w["merger.mlp.2.weight"], w["merger.mlp.2.bias"])
    if cfg.arch == "qwen2_5_vl":
        out = out[torch.argsort(wi)]
    if return_hidden:
        return out, x
    return out


def flops_per_batch(cfg: TowerConfig, grid_thw) -> float:
    """Algorithmic FLOPs of one tower forward (SURVEY.md section 8d formula)."""
    g = np.asarray(grid_thw, dtype=np.int64)
    N = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
    D, Fh, O = cfg.embed_dim, cfg.mlp_hidden, cfg.out_hidden
    cu_full = cu_seqlens(g).astype(np.int64)
    l2_full = float(((cu_full[1:] - cu_full[:-1]) ** 2).sum())
    mlp = 4.0 * N * D * Fh if cfg.arch == "qwen2_vl" else 6.0 * N * D * Fh
    lin = 2.0 * N * D * 3 * D + 2.0 * N * D * D + mlp
    total = 2.0 * N * cfg.patch_dim * D
    if cfg.arch == "qwen2_5_vl":
        _, cu_w = window_index(g, cfg.window_size, cfg.spatial_merge_size, cfg.patch_size)
        cu_w = cu_w.astype(np.int64)
        l2_win = float(((cu_w[1:] - cu_w[:-1]) ** 2).sum())
    for i in range(cfg.depth):
        l2 = l2_full
        if cfg.arch == "qwen2_5_vl" and i not in cfg.fullatt_block_indexes:
            l2 = l2_win
        total += lin + 4.0 * D * l2
    m = cfg.spatial_merge_size ** 2
    total += 2.0 * (N / m) * (m * D) ** 2 + 2.0 * (N / m) * (m * D) * O
    return total


# ---------------------------------------------------------------- LLM hand-off

def mrope_position_ids(input_ids, image_grid_thw, attention_mask=None, image_token_id=151655, merge=2):
    """numpy restatement of Qwen2VLModel.get_rope_index for still images (HF modeling_qwen2_vl.py:990-1092 with
    get_vision_position_ids :934-988): -> (position_ids int64 [3,B,L], deltas int64 [B,1])."""
    ids = np.asarray(input_ids, dtype=np.int64)
    B, L = ids.shape
    pos = np.zeros((3, B, L), dtype=np.int64)
    deltas = np.zeros((B, 1), dtype=np.int64)
    grids = iter(np.asarray(image_grid_thw, dtype=np.int64).reshape(-1, 3).tolist())
    for b in range(B):
        keep = np.arange(L) if attention_mask is None else np.nonzero(np.asarray(attention_mask)[b])[0]
        cur_ids = ids[b, keep]
        is_img = (cur_ids == image_token_id).tolist()
        chunks, cur, k = [], 0, 0
        while k < len(cur_ids):
            e = k
            while e < len(cur_ids) and is_img[e] == is_img[k]:
                e += 1
            if not is_img[k]:
                chunks.append(np.tile(np.arange(e - k) + cur, (3, 1)))
                cur += e - k
            else:
                t, gh, gw = next(grids)
                gh, gw = gh // merge, gw // merge
                assert t == 1 and gh * gw == e - k
                w = np.tile(np.arange(cur, cur + gw), gh * t)
                h = np.repeat(np.arange(cur, cur + gh), gw * t)
                chunks.append(np.stack([np.full(gh * gw * t, cur), h, w]))
                cur += max(gh, gw)
            k = e
        llm = np.concatenate(chunks, axis=1)
        pos[:, b, keep] = llm
        deltas[b, 0] = llm.max() + 1 - len(cur_ids)
    return pos, deltas


def scatter_image_features(inputs_embeds: torch.Tensor, input_ids, image_embeds: torch.Tensor, image_token_id=151655):
    """inputs_embeds.masked_scatter(input_ids == image_token_id, image_embeds) (HF modeling_qwen2_vl.py:1138-1177)."""
    mask = (torch.as_tensor(input_ids) == image_token_id)
    if int(mask.sum()) != image_embeds.shape[0]:
        raise ValueError(f"Image features and image tokens do not match, tokens: {int(mask.sum())}, features: {image_embeds.shape[0]}")
    m = mask.unsqueeze(-1).expand_as(inputs_embeds)
    return inputs_embeds.masked_scatter(m, image_embeds.to(inputs_embeds.dtype))
