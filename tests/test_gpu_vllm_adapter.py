"""the vLLM tower adapter against vLLM's own Qwen2VisionTransformer / Qwen2_5_VisionTransformer
(the modules the served path runs, vllm/model_executor/models/qwen2_vl.py, qwen2_5_vl.py), built stand-alone on the GPU
with seeded weights. Skipped where vLLM cannot be imported or initialised."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po
from tests.synth import synth_page

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture(scope="module")
def vllm_env():
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("VLLM_LOGGING_LEVEL", "ERROR")
    try:
        from vllm.config import VllmConfig, set_current_vllm_config
        from vllm.distributed import init_distributed_environment, initialize_model_parallel
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vllm not importable: {e}")
    try:
        ctx = set_current_vllm_config(VllmConfig())
        ctx.__enter__()
        if not torch.distributed.is_initialized():
            init_distributed_environment(world_size=1, rank=0, distributed_init_method="tcp://127.0.0.1:29547", local_rank=0, backend="nccl")
        initialize_model_parallel(1, 1)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vllm stand-alone initialisation failed: {e}")
    yield
    ctx.__exit__(None, None, None)


def _build(arch, depth):
    if arch == "qwen2_vl":
        from transformers.models.qwen2_vl.configuration_qwen2_vl import Qwen2VLVisionConfig
        from vllm.model_executor.models.qwen2_vl import Qwen2VisionTransformer as T
        cfg = Qwen2VLVisionConfig(depth=depth, embed_dim=1280, hidden_size=1536, mlp_ratio=4, num_heads=16)
    else:
        from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
        from vllm.model_executor.models.qwen2_5_vl import Qwen2_5_VisionTransformer as T
        cfg = Qwen2_5_VLVisionConfig(depth=depth, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=2048,
                                     fullatt_block_indexes=[1])
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            m = T(cfg)
    finally:
        torch.set_default_dtype(torch.float32)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.ndim == 1 and "norm" in k or k.endswith("ln_q.weight"):
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g)).to(p.dtype) if k.endswith("weight") else (0.1 * torch.randn(p.shape, generator=g)).to(p.dtype))
            elif p.ndim == 1:
                p.copy_((0.1 * torch.randn(p.shape, generator=g)).to(p.dtype))
            else:
                p.copy_((0.02 * torch.randn(p.shape, generator=g)).to(p.dtype))
    return m.eval()


@pytest.mark.parametrize("arch", ["qwen2_vl", "qwen2_5_vl"])
def test_adapter_matches_vllm_tower(vllm_env, arch):
    from karanta_ocr_b200.vllm_adapter import KarantaVllmVisual
    try:
        m = _build(arch, depth=3)
    except Exception as e:  # pragma: no cover - a vLLM build that cannot construct its own tower stand-alone
        pytest.skip(f"vllm tower construction failed: {type(e).__name__}: {e}")
    mine = KarantaVllmVisual.from_vllm(m)
    assert mine.out_hidden_size == m.out_hidden_size and mine.spatial_merge_size == m.spatial_merge_size
    pages = [synth_page(420, 336, 5), synth_page(252, 588, 6), synth_page(1288, 995, 7)]
    pv_np, grid = po.preprocess(pages, 3136, 12845056, po.RESIZE_ATEN)
    pv = torch.from_numpy(pv_np).cuda()
    glist = [[int(v) for v in g] for g in np.asarray(grid)]
    with torch.no_grad():
        out = mine(pv, grid_thw=glist).float()
        try:
            ref = m(pv.to(torch.bfloat16), grid_thw=glist).float()
        except Exception as e:  # pragma: no cover - vLLM's own attention back-end unusable on this box
            pytest.skip(f"vllm tower forward failed: {type(e).__name__}: {e}")
    assert out.shape == ref.shape
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    rel = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"vllm adapter {arch}: cosine {cos:.6f} rel-max {rel:.4f}")
    assert cos >= 0.999 and rel <= 0.03  # two bf16 towers against each other: each within tau = 1.5 x 0.0076 of fp32 at depth 3 (g5 golden), so <= 2 tau + margin


def test_replace_vllm_visual_swaps_module(vllm_env):
    from karanta_ocr_b200.vllm_adapter import KarantaVllmVisual, replace_vllm_visual

    class Holder(torch.nn.Module):
        def __init__(self, v):
            super().__init__()
            self.visual = v
    try:
        v = _build("qwen2_vl", depth=1)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vllm tower construction failed: {type(e).__name__}: {e}")
    h = replace_vllm_visual(Holder(v))
    assert isinstance(h.visual, KarantaVllmVisual)
    with pytest.raises(RuntimeError):
        h.visual.load_weights([])
