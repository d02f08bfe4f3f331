"""Host-side pieces of the vLLM adapter that need neither vLLM nor a GPU: reading a tower's
geometry off a module with vLLM's attribute layout and re-splitting Qwen2.5-VL's merged gate_up_proj."""
import types

import pytest
import torch

from karanta_ocr_b200 import vllm_adapter as va


class _Lin(torch.nn.Module):
    def __init__(self, o, i):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(o, i))
        self.bias = torch.nn.Parameter(torch.randn(o))


def _fake_visual(is25, depth=2, D=32, F=24, out=48):
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = torch.nn.Module()
            self.attn.qkv, self.attn.proj = _Lin(3 * D, D), _Lin(D, D)
            self.mlp = torch.nn.Module()
            if is25:
                self.mlp.gate_up_proj, self.mlp.down_proj = _Lin(2 * F, D), _Lin(D, F)
            else:
                self.mlp.fc1, self.mlp.fc2 = _Lin(F, D), _Lin(D, F)
    v = torch.nn.Module()
    v.blocks = torch.nn.ModuleList([Block() for _ in range(depth)])
    v.merger = torch.nn.Module()
    v.merger.mlp = torch.nn.ModuleList([_Lin(4 * D, 4 * D), torch.nn.GELU(), _Lin(out, 4 * D)])
    v.num_heads = 4
    if is25:
        v.fullatt_block_indexes, v.window_size = [1], 112
    return v


def test_config_from_qwen2_vl_layout():
    cfg = va._tower_config_from_vllm(_fake_visual(False))
    assert cfg == dict(arch="qwen2_vl", depth=2, embed_dim=32, num_heads=4, mlp_hidden=24, out_hidden=48, window_size=112,
                       fullatt_block_indexes=[])


def test_config_and_split_for_qwen2_5_vl_layout():
    v = _fake_visual(True)
    cfg = va._tower_config_from_vllm(v)
    assert cfg["arch"] == "qwen2_5_vl" and cfg["mlp_hidden"] == 24 and cfg["fullatt_block_indexes"] == [1]
    sd = va.hf_state_dict_from_vllm(v)
    assert not any("gate_up_proj" in k for k in sd)
    gu = v.blocks[0].mlp.gate_up_proj
    assert torch.equal(sd["blocks.0.mlp.gate_proj.weight"], gu.weight[:24]) and torch.equal(sd["blocks.0.mlp.up_proj.weight"], gu.weight[24:])
    assert torch.equal(sd["blocks.0.mlp.gate_proj.bias"], gu.bias[:24]) and torch.equal(sd["blocks.0.mlp.up_proj.bias"], gu.bias[24:])
    assert sd["blocks.1.mlp.down_proj.weight"].shape == (32, 24)


def test_tensor_parallel_tower_is_refused():
    v = _fake_visual(False)
    v.tp_size = 2
    v.device = torch.device("cpu")
    with pytest.raises(RuntimeError, match="tensor-parallel"):
        va.KarantaVllmVisual.from_vllm(v)
