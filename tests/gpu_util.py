"""Thin torch <-> C-ABI helpers for the GPU parity tests (every call goes through libkocr.so's extern "C" surface)."""
import ctypes as C

import numpy as np
import torch

from karanta_ocr_b200 import _lib


def ctx(dev=0):
    return _lib.context(dev)


def stream():
    return torch.cuda.current_stream().cuda_stream


def op_gemm(A, B, bias=None, residual=None, epilogue=_lib.EPI_NONE, out_cols=None):
    M, K = A.shape
    N = B.shape[0]
    oc = out_cols if out_cols is not None else (N // 2 if epilogue == _lib.EPI_BIAS_SWIGLU else N)
    out = torch.zeros((M, oc), dtype=torch.bfloat16, device=A.device)
    rc = _lib.load().kocr_op_gemm(ctx(), A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                  bias.data_ptr() if bias is not None else None,
                                  residual.data_ptr() if residual is not None else None, out.data_ptr(), out.stride(0), M, N, K,
                                  epilogue, stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out


def op_norm(x, w, b=None, eps=1e-6, rms=False):
    y = torch.empty_like(x)
    rc = _lib.load().kocr_op_norm(ctx(), x.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None, y.data_ptr(),
                                  x.shape[0], x.shape[1], eps, int(rms), stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    return y


def op_attention(qkv, cu, heads):
    """qkv bf16 [S, heads*240] in the tower layout (per head: q|k|v, 80 each; q pre-scaled, log2 domain)."""
    S = qkv.shape[0]
    out = torch.zeros((S, heads * 80), dtype=torch.bfloat16, device=qkv.device)
    cu = np.ascontiguousarray(np.asarray(cu, dtype=np.int32))
    rc = _lib.load().kocr_op_attention(ctx(), qkv.data_ptr(), out.data_ptr(), cu.ctypes.data, len(cu) - 1, heads, 80, stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out


def attention_reference(q, k, v, cu):
    """f32 reference: q,k,v [S, H, 80] (q unscaled), softmax(q k^T / sqrt(80)) v per segment."""
    out = torch.empty_like(q, dtype=torch.float32)
    for a, b in zip(cu[:-1], cu[1:]):
        qs, ks, vs = (t[a:b].float().transpose(0, 1) for t in (q, k, v))
        s = torch.matmul(qs, ks.transpose(1, 2)) * (80 ** -0.5)
        out[a:b] = torch.matmul(torch.softmax(s, dim=-1), vs).transpose(0, 1)
    return out


def pack_qkv(q, k, v):
    """[S,H,80] x3 -> tower layout [S, H*240] bf16 with q scaled by 80^-0.5 * log2(e)."""
    scale = (80 ** -0.5) * 1.4426950408889634
    x = torch.stack([q.float() * scale, k.float(), v.float()], dim=2)  # [S,H,3,80]
    return x.reshape(q.shape[0], -1).to(torch.bfloat16).contiguous()
