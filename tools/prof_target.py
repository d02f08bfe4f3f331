"""Small profiling target: Qwen2-VL-7B widths at depth 1 (or Qwen2.5-VL-7B widths at depth 2: one windowed block, one full-attention
block), a few letter pages, two encodes (first = warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, presets
from tests.synth import synth_page

n_pages = int(sys.argv[1]) if len(sys.argv) > 1 else 8
arch = sys.argv[2] if len(sys.argv) > 2 else "qwen2_vl_7b"   # "qwen2_5_vl_7b": depth 2 = one windowed + one full-attention block
cfg = presets.preset(arch, depth=1) if arch == "qwen2_vl_7b" else presets.preset(arch, depth=2, fullatt_block_indexes=[1])
tower = KarantaVisionTower(cfg)
tower.load_state_dict(presets.random_state_dict(cfg, seed=0))
enc = PageEncoder(tower)
pages = [torch.from_numpy(synth_page(1288, 995, 1234 + i)).cuda() for i in range(n_pages)]
for _ in range(2):
    emb, grid = enc.encode(pages)
    torch.cuda.synchronize()
print("ok", tuple(emb.shape), float(emb.float().abs().mean()))
