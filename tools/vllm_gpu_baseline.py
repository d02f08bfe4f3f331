"""Library baseline on the same GPU, second arm: vLLM's Qwen2VisionTransformer (the tower the
production path serves with, vllm/model_executor/models/qwen2_vl.py) built stand-alone in bf16 with its own ViT attention
back-end, same random weights and pages as tools/hf_gpu_baseline.py. Prints one JSON line. Evidence only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
import torch  # noqa: E402

import bench  # noqa: E402
from karanta_ocr_b200 import KarantaImageProcessor, KarantaVisionTower, presets  # noqa: E402

pages_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
pages = bench.make_pages(pages_n)
proc = KarantaImageProcessor(min_pixels=bench.MIN_PIXELS, max_pixels=bench.MAX_PIXELS)
feat = proc(images=[torch.from_numpy(p).to(dev) for p in pages], return_tensors="pt")
pv, grid = feat["pixel_values"].to(dev), feat["image_grid_thw"]
cfg = presets.preset("qwen2_vl_7b")
flops = presets.flops_per_batch(cfg, grid.numpy())["total"]
sd = presets.random_state_dict(cfg, seed=0)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


tower = KarantaVisionTower(cfg, device=dev)
tower.load_state_dict(sd)
with torch.no_grad():
    mine = tower(pv, grid_thw=grid).float()

try:
    import vllm
    from transformers.models.qwen2_vl.configuration_qwen2_vl import Qwen2VLVisionConfig
    from vllm.config import VllmConfig, set_current_vllm_config
    from vllm.distributed import init_distributed_environment, initialize_model_parallel
    from vllm.model_executor.models.qwen2_vl import Qwen2VisionTransformer

    vcfg = Qwen2VLVisionConfig(depth=32, embed_dim=1280, hidden_size=3584, mlp_ratio=4, num_heads=16)
    with set_current_vllm_config(VllmConfig()):
        init_distributed_environment(world_size=1, rank=0, distributed_init_method="tcp://127.0.0.1:29541", local_rank=0, backend="nccl")
        initialize_model_parallel(1, 1)
        torch.set_default_dtype(torch.bfloat16)
        with torch.device(dev):
            m = Qwen2VisionTransformer(vcfg)
        torch.set_default_dtype(torch.float32)
        m.load_weights((k, v.to(torch.bfloat16)) for k, v in sd.items())
        m = m.to(dev).eval()
        glist = grid.tolist()
        with torch.no_grad():
            out = m(pv.to(torch.bfloat16), grid_thw=glist)
            ms = timed(lambda: m(pv.to(torch.bfloat16), grid_thw=glist))
        cos = float(torch.nn.functional.cosine_similarity(out.float().flatten(), mine.flatten(), dim=0))
        print(json.dumps({"impl": f"vllm {vllm.__version__} Qwen2VisionTransformer, bf16, vit attention backend {m.attn_backend}", "pages": pages_n,
                          "ms_per_step": ms, "pages_per_s": pages_n / ms * 1e3, "model_tflops": flops / ms / 1e9, "cosine_vs_this_repo": cos}), flush=True)
except Exception as e:
    import traceback
    traceback.print_exc()
    print(json.dumps({"impl": "vllm Qwen2VisionTransformer", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}), flush=True)
