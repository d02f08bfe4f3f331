"""vLLM tower adapter.

The production path of karanta-ocr is HTTP -> `vllm serve` (karanta/pipeline.py:707-742, bulk_processing/workers/
vllm_client.py:155-227); inside vLLM the page embeddings come from `self.visual(pixel_values, grid_thw=grid_thw)`
(vllm/model_executor/models/qwen2_vl.py:1376, qwen2_5_vl.py `_process_image_input`). `KarantaVllmVisual` is a module
with that call surface - forward(x, grid_thw: list[list[int]] | Tensor, *, encoder_metadata=None) -> Tensor[sum N/4, out] -
backed by the kernels of this repo; `replace_vllm_visual(model)` swaps it in after vLLM has loaded the weights:

    from karanta_ocr_b200.vllm_adapter import replace_vllm_visual
    replace_vllm_visual(llm_model)        # llm_model = the Qwen2VLForConditionalGeneration / Qwen2_5_VL... vLLM built

vLLM keeps HF parameter names for the tower; the only re-packing is Qwen2.5-VL's merged `mlp.gate_up_proj`
(rows [gate; up]) which is split back into gate_proj / up_proj. Tensor-parallel towers (tp_size > 1) are refused:
this tower runs replicated per GPU, which is also what vLLM's own `mm_encoder_tp_mode="data"` does.
"""
from __future__ import annotations

import torch

from .vision_tower import KarantaVisionTower


def _tower_config_from_vllm(visual) -> dict:
    blocks = visual.blocks
    embed = int(blocks[0].attn.qkv.weight.shape[1])
    is25 = hasattr(visual, "fullatt_block_indexes")
    if is25:
        mlp_hidden = int(blocks[0].mlp.down_proj.weight.shape[1])
        full = [int(i) for i in visual.fullatt_block_indexes]
        window = int(visual.window_size)
    else:
        mlp_hidden = int(blocks[0].mlp.fc1.weight.shape[0])
        full, window = [], 112
    return dict(arch="qwen2_5_vl" if is25 else "qwen2_vl", depth=len(blocks), embed_dim=embed, num_heads=int(visual.num_heads),
                mlp_hidden=mlp_hidden, out_hidden=int(visual.merger.mlp[2].weight.shape[0]), window_size=window,
                fullatt_block_indexes=full)


def hf_state_dict_from_vllm(visual) -> dict:
    """vLLM tower parameters under HF names (splits the merged gate_up_proj of Qwen2.5-VL)."""
    out = {}
    for k, v in visual.state_dict().items():
        if ".mlp.gate_up_proj." in k:
            half = v.shape[0] // 2
            out[k.replace("gate_up_proj", "gate_proj")] = v[:half]
            out[k.replace("gate_up_proj", "up_proj")] = v[half:]
        else:
            out[k] = v
    return out


class KarantaVllmVisual(torch.nn.Module):
    """Stands in for vLLM's Qwen2VisionTransformer / Qwen2_5_VisionTransformer instance (`model.visual`)."""

    def __init__(self, tower: KarantaVisionTower):
        super().__init__()
        self.tower = tower
        self.spatial_merge_size = 2
        self.out_hidden_size = tower.cfg["out_hidden"]
        self.num_heads = tower.cfg["num_heads"]
        self.embed_dim = self.hidden_size = tower.cfg["embed_dim"]

    @classmethod
    def from_vllm(cls, visual) -> "KarantaVllmVisual":
        if int(getattr(visual, "tp_size", 1)) != 1:
            raise RuntimeError("KarantaVllmVisual: the vision tower must not be tensor-parallel (use mm_encoder_tp_mode='data')")
        tower = KarantaVisionTower(_tower_config_from_vllm(visual), device=visual.device)
        tower.load_state_dict(hf_state_dict_from_vllm(visual))
        return cls(tower)

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    @property
    def device(self) -> torch.device:
        return self.tower.device

    @torch.no_grad()
    def forward(self, x: torch.Tensor, grid_thw, *, encoder_metadata=None) -> torch.Tensor:
        # encoder_metadata (vLLM's pre-computed rope / cu_seqlens for CUDA-graph capture) is not needed: the library plans
        # positions, sequence and window tables itself from grid_thw
        return self.tower(x, grid_thw=grid_thw)

    def load_weights(self, weights) -> set:
        raise RuntimeError("KarantaVllmVisual is built from an already-loaded vLLM tower (from_vllm); it does not load checkpoints")


def replace_vllm_visual(model):
    """Swap `model.visual` (vLLM Qwen2-VL / Qwen2.5-VL model object, weights loaded) for the B200 tower; returns the model."""
    model.visual = KarantaVllmVisual.from_vllm(model.visual)
    return model
