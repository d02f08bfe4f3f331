"""LLM hand-off after the vision tower: the two steps Qwen2-VL runs between `visual(...)`
and the language model, behind the names transformers uses.

  get_rope_index(...)            Qwen2VLModel.get_rope_index (modeling_qwen2_vl.py:990-1092): 3-D M-RoPE position ids
  scatter_image_features(...)    get_placeholder_mask + inputs_embeds.masked_scatter (modeling_qwen2_vl.py:1138-1177)

Both sit on the path karanta-ocr reaches through model(**batch) (karanta/training/ocr_training.py:86,670) and
model.generate (karanta/training/test_trained_model.py:91). Position ids are host planning (integer work, bit-exact);
the scatter is one HBM-bound kernel in libkocr.so. Still images only (grid t == 1), as karanta-ocr sends pages.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

IMAGE_TOKEN_ID = 151655  # Qwen2VLConfig.image_token_id


VISION_START_TOKEN_ID = 151652  # Qwen2VLConfig.vision_start_token_id


def get_rope_index(input_ids, image_grid_thw=None, attention_mask=None, image_token_id: int = IMAGE_TOKEN_ID,
                   spatial_merge_size: int = 2, semantics: str = "5.x", vision_start_token_id: int | None = VISION_START_TOKEN_ID):
    """-> (position_ids int64 [3, batch, seq_len], mrope_position_deltas int64 [batch, 1]) on the CPU.

    `semantics` picks the transformers line to reproduce. "5.x" (default) is what the build image runs and what
    tests/golden/g6_llm_handoff.npz pins: padded positions are 0 and delta = max + 1 - (unmasked length). "4.5x" is the line
    karanta-ocr pins (4.53.3, /root/reference/uv.lock:2168-2169): padded positions are 1, delta = max + 1 - (padded length) -
    the value generate() adds to cache_position, so left-padded batches decode at the positions that stack expects - and
    images are located by <|vision_start|> (vision_start_token_id=None: by runs of image tokens) and consume exactly their
    grid's token count, so adjacent images need no separator. The 4.5x mode is restated from its source and has no golden here
    (that version is not installed offline): use it when the surrounding model code is transformers 4.5x."""
    if semantics not in ("5.x", "4.5x"):
        raise ValueError("semantics must be '5.x' or '4.5x'")
    ids = np.ascontiguousarray(torch.as_tensor(input_ids).detach().cpu().numpy().astype(np.int64))
    if ids.ndim != 2:
        raise ValueError("input_ids must be [batch, seq_len]")
    B, L = ids.shape
    mask = None
    if attention_mask is not None:
        mask = np.ascontiguousarray(torch.as_tensor(attention_mask).detach().cpu().numpy().astype(np.int64))
        if mask.shape != ids.shape:
            raise ValueError("attention_mask must have the shape of input_ids")
    grid = np.zeros((0, 3), dtype=np.int64) if image_grid_thw is None else np.ascontiguousarray(
        torch.as_tensor(image_grid_thw).detach().cpu().numpy().astype(np.int64).reshape(-1, 3))
    pos = np.zeros((3, B, L), dtype=np.int64)
    deltas = np.zeros((B,), dtype=np.int64)
    rc = _lib.load().kocr_mrope_position_ids_v2(ids.ctypes.data, mask.ctypes.data if mask is not None else None, B, L,
                                                grid.ctypes.data if len(grid) else None, len(grid), int(image_token_id),
                                                -1 if vision_start_token_id is None else int(vision_start_token_id),
                                                int(spatial_merge_size), 0 if semantics == "5.x" else 1, pos.ctypes.data,
                                                deltas.ctypes.data)
    _lib.check(rc)
    return torch.from_numpy(pos), torch.from_numpy(deltas).unsqueeze(1)


@torch.no_grad()
def scatter_image_features(inputs_embeds: torch.Tensor, input_ids, image_embeds: torch.Tensor,
                           image_token_id: int = IMAGE_TOKEN_ID) -> torch.Tensor:
    """In place: inputs_embeds[input_ids == image_token_id] = image_embeds (row order). inputs_embeds bf16 [B, L, H] on
    the GPU, image_embeds bf16 [n, H]. Raises ValueError like transformers when the counts disagree."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError("scatter_image_features runs on CUDA tensors only (no CPU fallback)")
    if inputs_embeds.dtype != torch.bfloat16 or not inputs_embeds.is_contiguous():
        raise ValueError("inputs_embeds must be a contiguous bfloat16 tensor")
    ids = np.ascontiguousarray(torch.as_tensor(input_ids).detach().cpu().numpy().astype(np.int64))
    B, L = ids.shape
    H = inputs_embeds.shape[-1]
    if inputs_embeds.numel() != B * L * H:
        raise ValueError("inputs_embeds does not match input_ids")
    dev = inputs_embeds.device
    src = image_embeds.to(device=dev, dtype=torch.bfloat16).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().kocr_scatter_image_embeds(_lib.context(dev.index), inputs_embeds.data_ptr(), src.data_ptr(), src.shape[0], H,
                                                   ids.ctypes.data, B, L, int(image_token_id),
                                                   torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc)
        src.record_stream(torch.cuda.current_stream(dev))
    return inputs_embeds
