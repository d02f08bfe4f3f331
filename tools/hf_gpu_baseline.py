"""Library baseline on the same GPU (last row): the HF transformers Qwen2-VL vision tower in bf16 on
cuda with its stock attention back-ends (flash_attention_2 = flash-attn 2.8 library kernels, and sdpa), random weights,
same synthetic letter pages as bench.py's C2, pixel_values taken from this repo's preprocess so that only the tower is
compared. Prints one JSON line per back-end. Not part of the product or of bench.py; evidence for profiles/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
import torch  # noqa: E402

import bench  # noqa: E402
from karanta_ocr_b200 import KarantaImageProcessor, KarantaVisionTower, PageEncoder, presets  # noqa: E402

pages_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
pages = bench.make_pages(pages_n)
proc = KarantaImageProcessor(min_pixels=bench.MIN_PIXELS, max_pixels=bench.MAX_PIXELS)
feat = proc(images=[torch.from_numpy(p).to(dev) for p in pages], return_tensors="pt")
pv, grid = feat["pixel_values"].to(dev), feat["image_grid_thw"].to(dev)
cfg = presets.preset("qwen2_vl_7b")
flops = presets.flops_per_batch(cfg, grid.cpu().numpy())["total"]


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


from transformers.models.qwen2_vl.configuration_qwen2_vl import Qwen2VLVisionConfig  # noqa: E402
from transformers.models.qwen2_vl.modeling_qwen2_vl import Qwen2VisionTransformerPretrainedModel  # noqa: E402

sd = presets.random_state_dict(cfg, seed=0)
results = []
ref_out = None
for impl in ("flash_attention_2", "sdpa"):
    try:
        c = Qwen2VLVisionConfig(depth=32, embed_dim=1280, hidden_size=3584, mlp_ratio=4, num_heads=16)
        c._attn_implementation = impl
        m = Qwen2VisionTransformerPretrainedModel(c).eval()
        m.load_state_dict(sd, strict=False)
        m = m.to(dev, torch.bfloat16)
        with torch.no_grad():
            out = m(pv.to(torch.bfloat16), grid_thw=grid)
            out = getattr(out, "pooler_output", out)
            ms = timed(lambda: m(pv.to(torch.bfloat16), grid_thw=grid))
        ref_out = out.float()
        results.append({"impl": f"transformers {__import__('transformers').__version__} tower, bf16, attn={impl}", "pages": pages_n,
                        "ms_per_step": ms, "pages_per_s": pages_n / ms * 1e3, "model_tflops": flops / ms / 1e9})
        del m
        torch.cuda.empty_cache()
    except Exception as e:  # back-end not usable in this image
        results.append({"impl": f"transformers tower attn={impl}", "unavailable": f"{type(e).__name__}: {str(e)[:200]}"})

tower = KarantaVisionTower(cfg, device=dev)
tower.load_state_dict(sd)
with torch.no_grad():
    mine = tower(pv, grid_thw=grid)
    ms = timed(lambda: tower(pv, grid_thw=grid))
line = {"impl": "this repo: KarantaVisionTower.forward(pixel_values f32, grid_thw)", "pages": pages_n, "ms_per_step": ms,
        "pages_per_s": pages_n / ms * 1e3, "model_tflops": flops / ms / 1e9}
if ref_out is not None:
    a, b = mine.float().flatten(), ref_out.flatten()
    line["cosine_vs_hf_bf16"] = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
results.append(line)
for r in results:
    print(json.dumps(r), flush=True)
