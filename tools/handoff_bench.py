"""Scatter kernel bandwidth (LLM hand-off, row f3): 64 prompts x 1656 image tokens x 3584 bf16 = 0.76 GB read + 0.76 GB written."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
from karanta_ocr_b200 import scatter_image_features

IMG = 151655
B, L, H, T = 64, 1700, 3584, 1656
ids = torch.full((B, L), 11, dtype=torch.int64)
ids[:, 20:20 + T] = IMG
emb = torch.zeros(B, L, H, dtype=torch.bfloat16, device="cuda")
img = torch.randn(B * T, H, device="cuda").to(torch.bfloat16)
for _ in range(3):
    scatter_image_features(emb, ids, img, IMG)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    scatter_image_features(emb, ids, img, IMG)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byt = 2 * B * T * H * 2
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
print(json.dumps({"kernel": "scatter_rows_kernel", "ms_per_call_incl_host_planning": ms, "bytes": byt, "gbs": byt / ms / 1e6, "frac_of_measured_hbm_peak": byt / ms / 1e6 / peak}))
