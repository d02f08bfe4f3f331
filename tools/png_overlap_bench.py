"""How long does the GPU PNG decode of one 64-page batch take alone, and what does it do to a tower that runs beside it?
Letter pages as tests/synth.py makes them (noisy background) and the same pages with a clean background (what a PDF renderer
emits). CUDA events; Qwen2-VL-7B widths at depth 8 to keep the run short."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, presets
from karanta_ocr_b200.png_decode import decode_png_batch
from tests.synth import synth_page

def png_of(p):
    buf = io.BytesIO(); Image.fromarray(np.ascontiguousarray(p.transpose(1, 2, 0))).save(buf, format="PNG"); return buf.getvalue()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

noisy = [synth_page(1288, 995, 1234 + i) for i in range(8)]
clean = [np.where(p > 200, 250, p).astype(np.uint8) for p in noisy]
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = presets.preset("qwen2_vl_7b", depth=depth)
tower = KarantaVisionTower(cfg); tower.load_state_dict(presets.random_state_dict(cfg, seed=0))
enc = PageEncoder(tower)
dev_pages = [torch.from_numpy(noisy[i % 8]).cuda() for i in range(64)]
t_tower = timed(lambda: enc.encode(dev_pages))
print(f"tower alone (depth {depth}, 64 pages from device tensors): {t_tower:.1f} ms")
for name, pages in (("noisy", noisy), ("clean", clean)):
    files = [png_of(pages[i % 8]) for i in range(64)]
    mb = sum(len(f) for f in files) / 1e6
    t_dec = timed(lambda: decode_png_batch(files, check=False))
    print(f"{name}: 64 PNG files {mb:.1f} MB, decode alone {t_dec:.1f} ms = {64 / t_dec * 1e3:.0f} pages/s, {64 * 3.84 / t_dec:.2f} GB/s of pixels")
    # pipelined: encode_to_host_async over 6 batches of PNG files
    out = [torch.empty((64 * 1656, cfg["out_hidden"]), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    def job(n=6):
        pend = []
        for k in range(n):
            if len(pend) >= 2: pend.pop(0).synchronize()
            ev, _, _ = enc.encode_to_host_async(files, out[k % 2]); pend.append(ev)
        for ev in pend: ev.synchronize()
    t_job = timed(job, reps=1)
    print(f"{name}: 6 pipelined batches from PNG: {t_job:.1f} ms = {t_job / 6:.1f} ms per batch (tower alone {t_tower:.1f}, decode alone {t_dec:.1f})")
