"""Serving-latency experiment: one letter page (and 4) through the depth-32 Qwen2-VL-7B tower, launched kernel by kernel
(plan cached) against a CUDA-graph replay of the same forward (torch.cuda.graph around KarantaVisionTower.forward)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, presets
from tests.synth import synth_page

cfg = presets.preset("qwen2_vl_7b")
tower = KarantaVisionTower(cfg); tower.load_state_dict(presets.random_state_dict(cfg, seed=0))
enc = PageEncoder(tower)

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n in (1, 4):
    pages = [torch.from_numpy(synth_page(1288, 995, 1234 + i)).cuda() for i in range(n)]
    pv, grid = enc.processor.preprocess_device(pages, torch.bfloat16)
    t_eager = timed(lambda: tower(pv, grid_thw=grid))
    t_full = timed(lambda: enc.encode(pages))
    msg = f"{n} page(s): tower launched kernel by kernel {t_eager:.3f} ms, preprocess + tower {t_full:.3f} ms"
    try:
        gf = tower.capture(grid)
        ref = tower(pv, grid_thw=grid)
        out = gf.replay(pv); torch.cuda.synchronize()
        same = torch.equal(out, ref)
        t_graph = timed(lambda: gf.replay())
        t_graph_in = timed(lambda: gf.replay(pv))
        msg += f", CUDA-graph replay {t_graph:.3f} ms ({t_graph_in:.3f} ms with the input copy; bit-identical output: {same})"
    except Exception as e:
        msg += f", graph capture failed: {type(e).__name__}: {str(e)[:300]}"
    print(msg)
