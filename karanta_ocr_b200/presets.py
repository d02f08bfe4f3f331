"""Tower presets, random-initialised state dicts and FLOP accounting for benchmarks and smoke runs.

There are no checkpoints offline, so benchmark runs use random weights of the real architectures
(HF configuration_qwen2_vl.py / configuration_qwen2_5_vl.py; the 7B towers behind olmOCR-7B-0225 and -0725).
The FLOP formula is SURVEY.md section 8(d); sequence and window lengths come from the library's own planning functions
(kocr_cu_seqlens / kocr_window_index), which are host-only and need no GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

PRESETS = {
    "qwen2_vl_7b": dict(arch="qwen2_vl", depth=32, embed_dim=1280, num_heads=16, mlp_hidden=5120, out_hidden=3584,
                        window_size=112, fullatt_block_indexes=[]),
    "qwen2_vl_2b": dict(arch="qwen2_vl", depth=32, embed_dim=1280, num_heads=16, mlp_hidden=5120, out_hidden=1536,
                        window_size=112, fullatt_block_indexes=[]),
    "qwen2_5_vl_7b": dict(arch="qwen2_5_vl", depth=32, embed_dim=1280, num_heads=16, mlp_hidden=3420, out_hidden=3584,
                          window_size=112, fullatt_block_indexes=[7, 15, 23, 31]),
}


def preset(name: str, **overrides) -> dict:
    cfg = dict(PRESETS[name])
    cfg.update(overrides)
    return cfg


def random_state_dict(cfg: dict, seed: int = 0, std: float = 0.02) -> dict:
    """HF state-dict keys and shapes for `cfg`, N(0, std) weights, small non-zero biases, norm scales around 1."""
    g = torch.Generator().manual_seed(seed)
    D, Fh, O = cfg["embed_dim"], cfg["mlp_hidden"], cfg["out_hidden"]
    is25 = cfg["arch"] == "qwen2_5_vl"
    shapes = {"patch_embed.proj.weight": (D, 3, 2, 14, 14)}
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        for n in ("norm1", "norm2"):
            shapes[b + n + ".weight"] = (D,)
            if not is25:
                shapes[b + n + ".bias"] = (D,)
        shapes[b + "attn.qkv.weight"], shapes[b + "attn.qkv.bias"] = (3 * D, D), (3 * D,)
        shapes[b + "attn.proj.weight"], shapes[b + "attn.proj.bias"] = (D, D), (D,)
        lins = (("mlp.gate_proj", Fh, D), ("mlp.up_proj", Fh, D), ("mlp.down_proj", D, Fh)) if is25 else \
               (("mlp.fc1", Fh, D), ("mlp.fc2", D, Fh))
        for n, o, k in lins:
            shapes[b + n + ".weight"], shapes[b + n + ".bias"] = (o, k), (o,)
    shapes["merger.ln_q.weight"] = (D,)
    if not is25:
        shapes["merger.ln_q.bias"] = (D,)
    shapes["merger.mlp.0.weight"], shapes["merger.mlp.0.bias"] = (4 * D, 4 * D), (4 * D,)
    shapes["merger.mlp.2.weight"], shapes["merger.mlp.2.bias"] = (O, 4 * D), (O,)
    sd = {}
    for k, shp in shapes.items():
        x = torch.randn(*shp, generator=g, dtype=torch.float32)
        if k.endswith("bias"):
            sd[k] = x * 0.1
        elif len(shp) == 1:
            sd[k] = 1.0 + x * 0.1
        else:
            sd[k] = x * std
    return sd


def sequence_lengths(grid_thw, cfg: dict):
    """(full-attention sequence lengths, window lengths or None) for a batch, from the library's planning functions."""
    lib = _lib.load()
    g = np.ascontiguousarray(np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))
    n = len(g)
    cu = np.zeros(int(g[:, 0].sum()) + 1, dtype=np.int32)
    ncu = C.c_int(0)
    _lib.check(lib.kocr_cu_seqlens(g.ctypes.data, n, cu.ctypes.data, C.byref(ncu)))
    full = np.diff(cu[:ncu.value].astype(np.int64))
    if cfg["arch"] != "qwen2_5_vl":
        return full, None
    total = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
    wi = np.zeros(total // 4, dtype=np.int32)
    cuw = np.zeros(total // 4 + 1, dtype=np.int32)
    ncw = C.c_int(0)
    _lib.check(lib.kocr_window_index(g.ctypes.data, n, cfg["window_size"], 2, 14, wi.ctypes.data, cuw.ctypes.data, C.byref(ncw)))
    return full, np.diff(cuw[:ncw.value].astype(np.int64))


def flops_per_batch(cfg: dict, grid_thw) -> dict:
    """Algorithmic FLOPs of one tower forward over the batch: {'total', 'attention', 'attention_per_layer': [...]}."""
    g = np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3)
    N = float((g[:, 0] * g[:, 1] * g[:, 2]).sum())
    D, Fh, O = cfg["embed_dim"], cfg["mlp_hidden"], cfg["out_hidden"]
    full, win = sequence_lengths(g, cfg)
    l2_full = float((full ** 2).sum())
    l2_win = float((win ** 2).sum()) if win is not None else l2_full
    n_mlp_mats = 3 if cfg["arch"] == "qwen2_5_vl" else 2
    linear = 2.0 * N * D * (3 * D) + 2.0 * N * D * D + n_mlp_mats * 2.0 * N * D * Fh
    attn = [4.0 * D * (l2_full if (win is None or i in cfg["fullatt_block_indexes"]) else l2_win) for i in range(cfg["depth"])]
    merger = 2.0 * (N / 4) * (4 * D) ** 2 + 2.0 * (N / 4) * (4 * D) * O
    total = 2.0 * N * 1176 * D + cfg["depth"] * linear + sum(attn) + merger
    return {"total": total, "attention": sum(attn), "attention_per_layer": attn}
