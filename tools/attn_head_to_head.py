"""Attention-only head-to-head on one GPU: this repo's attention kernel through the C ABI
(kocr_op_attention) against the sm_100 kernel the served path reaches today - vLLM's FA4 (CuTe-DSL flash_fwd_sm100,
vllm/v1/attention/backends/fa_utils.py:81-83) - and against flash-attn 2.8 (what HF flash_attention_2 calls), at the C2
shape: S = pages x 6624 rows, 16 heads, head_dim 80, one sequence per page, non-causal, bf16. CUDA events, warm, same
inputs; the outputs are compared with each other. Prints one JSON line per implementation. Evidence for profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from karanta_ocr_b200 import _lib  # noqa: E402
from tests import gpu_util as gu  # noqa: E402

pages = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, N, HD = 16, 6624, 80
S = pages * N
g = torch.Generator().manual_seed(0)
q, k, v = ((torch.randn(S, H, HD, generator=g) * (1.0 if i < 2 else 0.5)).to(torch.bfloat16).cuda() for i in range(3))
cu = [i * N for i in range(pages + 1)]
flops = 4.0 * H * HD * pages * N * N


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def line(name, ms, out, ref):
    d = {"impl": name, "pages": pages, "rows": S, "heads": H, "head_dim": HD, "ms_per_launch": ms, "tflops": flops / ms / 1e9}
    if ref is not None and out is not None:
        a, b = out.float().flatten(), ref.float().flatten()
        d["cosine_vs_this_repo"] = float(torch.nn.functional.cosine_similarity(a[:1 << 24].double(), b[:1 << 24].double(), dim=0))
        d["max_abs_diff"] = float((a[:1 << 24] - b[:1 << 24]).abs().max())
    print(json.dumps(d), flush=True)


# this repo (tower layout: per head q|k|v, q pre-scaled into the log2 domain as the QKV epilogue leaves it)
qkv = gu.pack_qkv(q, k, v)
mine = torch.zeros((S, H * HD), dtype=torch.bfloat16, device="cuda")
cua = np.ascontiguousarray(np.asarray(cu, dtype=np.int32))
lib, ctx = _lib.load(), gu.ctx()


def run_mine():
    _lib.check(lib.kocr_op_attention(ctx, qkv.data_ptr(), mine.data_ptr(), cua.ctypes.data, pages, H, HD, gu.stream()))


ms = timed(run_mine)
mine_v = mine.view(S, H, HD)
line(f"this repo: attention kernels via kocr_op_attention ({os.path.basename(_lib.LIB_PATH)})", ms, None, None)
del qkv

cu_t = torch.tensor(cu, dtype=torch.int32, device="cuda")
try:  # vLLM's FA4 (CuTe-DSL) - the kernel vLLM picks on compute capability 10.x
    from vllm.vllm_flash_attn.flash_attn_interface import flash_attn_varlen_func as vfa, is_fa_version_supported
    for ver in (4, 2):
        if not is_fa_version_supported(ver):
            print(json.dumps({"impl": f"vllm_flash_attn fa_version={ver}", "unavailable": "not supported in this build"}), flush=True)
            continue
        fn = lambda: vfa(q, k, v, max_seqlen_q=N, cu_seqlens_q=cu_t, max_seqlen_k=N, cu_seqlens_k=cu_t, softmax_scale=HD ** -0.5,
                         causal=False, fa_version=ver)
        out = fn()
        ms = timed(fn)
        line(f"vllm {__import__('vllm').__version__} vllm_flash_attn.flash_attn_varlen_func fa_version={ver}"
             + (" (CuTe-DSL flash_fwd_sm100: tcgen05/TMEM/TMA)" if ver == 4 else " (mma.sync)"), ms, out, mine_v)
        del out
except Exception as e:
    import traceback
    traceback.print_exc()
    print(json.dumps({"impl": "vllm_flash_attn", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}), flush=True)

try:  # flash-attn 2.8 (HF attn_implementation="flash_attention_2")
    from flash_attn import flash_attn_varlen_func as fa2
    fn = lambda: fa2(q, k, v, cu_t, cu_t, N, N, softmax_scale=HD ** -0.5, causal=False)
    out = fn()
    ms = timed(fn)
    line(f"flash_attn {__import__('flash_attn').__version__} flash_attn_varlen_func", ms, out, mine_v)
except Exception as e:
    print(json.dumps({"impl": "flash_attn", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}), flush=True)
