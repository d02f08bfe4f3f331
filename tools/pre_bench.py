"""Preprocess-kernel micro-benchmark through the C ABI (KarantaImageProcessor.preprocess_device) on the C2 batch (64 letter pages)
and the C4 mixed batch; isolated launches (no tower around them), CUDA events. KOCR_PRE_TW=56|84|112 selects the tile width."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from karanta_ocr_b200 import KarantaImageProcessor

proc = KarantaImageProcessor(device="cuda")
for wl in ("c2", "c4"):
    pages = [torch.from_numpy(p).cuda() for p in bench.make_pages(64, workload=wl)]
    for _ in range(3):
        pv, grid = proc.preprocess_device(pages, torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        pv, grid = proc.preprocess_device(pages, torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = sum(p.numel() for p in pages) + pv.numel() * 2
    print(f"{wl} tw={os.environ.get('KOCR_PRE_TW', '84')}: {ms:.3f} ms per call (host planning + launch + kernel), {nbytes / ms / 1e6:.0f} GB/s")
