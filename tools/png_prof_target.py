"""Profiling target: GPU decode of 16 noisy letter-page PNG files (one warm-up call, one measured)."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from karanta_ocr_b200.png_decode import decode_png_batch
from tests.synth import synth_page
files = []
for i in range(4):
    p = synth_page(1288, 995, 1234 + i)
    buf = io.BytesIO(); Image.fromarray(np.ascontiguousarray(p.transpose(1, 2, 0))).save(buf, format="PNG"); files.append(buf.getvalue())
files = [files[i % 4] for i in range(16)]
for _ in range(2):
    out = decode_png_batch(files, check=True)
print("ok", out[0].shape)
