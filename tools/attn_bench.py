"""Attention-kernel micro-benchmark through the C ABI (kocr_op_attention) on C2-shaped inputs; run once per library
variant (KOCR_LIB=...). Prints ms per launch and TFLOP/s; checks one small case against torch first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, N = 16, 6624
g = torch.Generator().manual_seed(0)
# correctness on a small case
cu = [0, 300, 1000]
q, k, v = (torch.randn(cu[-1], 2, 80, generator=g).to(torch.bfloat16).cuda() for _ in range(3))
out = gu.op_attention(gu.pack_qkv(q, k, v), cu, 2).float()
ref = gu.attention_reference(q, k, v, cu).reshape(cu[-1], 160)
err = ((out - ref).abs().max() / ref.abs().max()).item()
S = pages * N
qkv = (torch.randn(S, H * 240, generator=g) * 0.5).to(torch.bfloat16).cuda()
cu = [i * N for i in range(pages + 1)]
for _ in range(3):
    gu.op_attention(qkv, cu, H)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
import numpy as np, ctypes as C
from karanta_ocr_b200 import _lib
o = torch.zeros((S, H * 80), dtype=torch.bfloat16, device="cuda")
cua = np.ascontiguousarray(np.asarray(cu, dtype=np.int32))
e0.record()
for _ in range(reps):
    _lib.check(_lib.load().kocr_op_attention(gu.ctx(), qkv.data_ptr(), o.data_ptr(), cua.ctypes.data, pages, H, 80, gu.stream()))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 4.0 * 1280 * pages * N * N
print(f"{os.path.basename(_lib.LIB_PATH):24s} pages {pages} rel-err {err:.4f}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")
