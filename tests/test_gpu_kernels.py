"""GPU parity of the individual sm_100a kernels through the C ABI (kocr_op_*), against plain torch fp32.
Floating point: tolerances stated per test (bf16 outputs, f32 accumulation)."""
import numpy as np
import pytest
import torch

from karanta_ocr_b200 import _lib
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def _close(out, ref, tol):
    """bf16 output vs f32 reference: |d| <= tol * max|ref| everywhere and rel-rms <= tol/4."""
    d = (out.float() - ref.float()).abs()
    scale = ref.float().abs().max().item() + 1e-12
    rms = (d.pow(2).mean().sqrt() / (ref.float().pow(2).mean().sqrt() + 1e-12)).item()
    assert d.max().item() <= tol * scale and rms <= tol / 2, (d.max().item() / scale, rms)


def test_gemm_ones_exact():
    A = torch.ones(128, 64, dtype=torch.bfloat16, device="cuda")
    B = torch.ones(256, 64, dtype=torch.bfloat16, device="cuda")
    out = gu.op_gemm(A, B)
    assert torch.equal(out.float(), torch.full((128, 256), 64.0, device="cuda"))


def test_gemm_identity_picks_columns():
    """B = identity rows -> C[m, n] = A[m, n]: catches any swizzle / descriptor / tile-order mix-up exactly."""
    K = 256
    A = _rand((384, K), 1).to(torch.bfloat16).cuda()
    B = torch.eye(K, dtype=torch.bfloat16, device="cuda")
    out = gu.op_gemm(A, B)
    assert torch.equal(out, A)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 160, 1176), (1000, 1280, 1280), (257, 3584, 5120), (6624, 1280, 5120),
                                   (77, 32, 8)])
def test_gemm_plain_and_bias(M, N, K):
    A = _rand((M, K), 2).to(torch.bfloat16).cuda()
    B = _rand((N, K), 3, 0.05).to(torch.bfloat16).cuda()
    bias = _rand((N,), 4).cuda()
    ref = A.float() @ B.float().t()
    _close(gu.op_gemm(A, B), ref, 1e-2)
    _close(gu.op_gemm(A, B, bias, epilogue=_lib.EPI_BIAS), ref + bias, 1e-2)


@pytest.mark.parametrize("epi", ["quickgelu", "gelu", "residual", "swiglu"])
def test_gemm_epilogues(epi):
    M, N, K = 700, 1280, 640
    A = _rand((M, K), 5).to(torch.bfloat16).cuda()
    B = _rand((N, K), 6, 0.05).to(torch.bfloat16).cuda()
    bias = _rand((N,), 7).cuda()
    y = A.float() @ B.float().t() + bias
    if epi == "quickgelu":
        _close(gu.op_gemm(A, B, bias, epilogue=_lib.EPI_BIAS_QUICKGELU), y * torch.sigmoid(1.702 * y), 1e-2)
    elif epi == "gelu":
        _close(gu.op_gemm(A, B, bias, epilogue=_lib.EPI_BIAS_GELU), torch.nn.functional.gelu(y), 1e-2)
    elif epi == "residual":
        R = _rand((M, N), 8).to(torch.bfloat16).cuda()
        _close(gu.op_gemm(A, B, bias, residual=R, epilogue=_lib.EPI_BIAS_RESIDUAL), y + R.float(), 1e-2)
    else:
        out = gu.op_gemm(A, B, bias, epilogue=_lib.EPI_BIAS_SWIGLU)
        ref = torch.nn.functional.silu(y[:, 0::2]) * y[:, 1::2]
        assert out.shape == (M, N // 2)
        _close(out, ref, 1e-2)


def test_gemm_rejects_bad_shapes():
    A = torch.zeros(8, 60, dtype=torch.bfloat16, device="cuda")
    B = torch.zeros(32, 60, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        gu.op_gemm(A, B)


@pytest.mark.parametrize("dim,rms", [(1280, False), (1280, True), (160, False), (5120, False)])
def test_norm(dim, rms):
    x = (_rand((1000, dim), 9) * 3 + 0.5).to(torch.bfloat16).cuda()
    w = (1 + 0.1 * _rand((dim,), 10)).cuda()
    b = None if rms else (0.1 * _rand((dim,), 11)).cuda()
    y = gu.op_norm(x, w, b, 1e-6, rms)
    xf = x.float()
    if rms:
        ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16).float()
    else:
        ref = torch.nn.functional.layer_norm(xf, (dim,), w, b, 1e-6)
    _close(y, ref, 1e-2)


def _attn_case(cu, heads, seed):
    S = cu[-1]
    q, k, v = (_rand((S, heads, 80), seed + i).to(torch.bfloat16).cuda() for i in range(3))
    out = gu.op_attention(gu.pack_qkv(q, k, v), cu, heads)
    ref = gu.attention_reference(q, k, v, cu).reshape(S, heads * 80)
    return out, ref


def test_attention_uniform_probabilities():
    """q = 0 -> every row of the output is the mean of V over its own sequence (exercises P.V, not Q.K^T)."""
    cu = [0, 256, 640]
    S, H = cu[-1], 2
    v = _rand((S, H, 80), 21).to(torch.bfloat16).cuda()
    q = torch.zeros_like(v)
    k = _rand((S, H, 80), 22).to(torch.bfloat16).cuda()
    out = gu.op_attention(gu.pack_qkv(q, k, v), cu, H).float().reshape(S, H, 80)
    for a, b in zip(cu[:-1], cu[1:]):
        ref = v[a:b].float().mean(0, keepdim=True).expand(b - a, -1, -1)
        assert (out[a:b] - ref).abs().max().item() < 2e-2


def test_attention_key_order():
    """Scores depend on the key index only -> the same distribution for every query; a permuted P/V pairing shows up."""
    S, H = 384, 1
    f = torch.linspace(-4, 4, S)
    q = torch.zeros(S, H, 80)
    q[:, :, 0] = 3.0
    k = torch.zeros(S, H, 80)
    k[:, 0, 0] = f
    v = _rand((S, H, 80), 23)
    q, k, v = (t.to(torch.bfloat16).cuda() for t in (q, k, v))
    out = gu.op_attention(gu.pack_qkv(q, k, v), [0, S], H)
    ref = gu.attention_reference(q, k, v, [0, S]).reshape(S, 80)
    _close(out, ref, 2e-2)


@pytest.mark.parametrize("cu,heads", [([0, 128], 1), ([0, 256], 2), ([0, 300], 2), ([0, 64, 128, 200], 3), ([0, 1000, 1324], 16),
                                      ([0, 6624], 2), ([0, 4, 24, 1660, 1724], 2)])
def test_attention_random(cu, heads):
    out, ref = _attn_case(cu, heads, 30)
    _close(out, ref, 2e-2)


def test_attention_longest_sequence():
    """The longest sequence the path can produce: max_pixels = 12 845 056 caps one image at 65 536 patches (e.g. a
    3584 x 3584 page), i.e. 1024 score sub-steps per query block. The fp32 reference would need a 17 GB score matrix per
    head, so this compares against torch's flash SDPA in bf16 on the same device (an independent implementation), plus the
    exact property that identical keys give the mean of V."""
    S, H = 65536, 2
    g = torch.Generator().manual_seed(77)
    q, k, v = (torch.randn(S, H, 80, generator=g).to(torch.bfloat16).cuda() for _ in range(3))
    out = gu.op_attention(gu.pack_qkv(q, k, v), [0, S], H).float().reshape(S, H, 80)
    with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])
    ref = ref[0].transpose(0, 1).float()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    assert err <= 3e-2, err  # two bf16 flash implementations against each other
    kc = k[:1].expand(S, H, 80).contiguous()
    out2 = gu.op_attention(gu.pack_qkv(q, kc, v), [0, S], H).float().reshape(S, H, 80)
    mean_v = v.float().mean(0, keepdim=True).expand(S, H, 80)
    assert ((out2 - mean_v).abs().max() / mean_v.abs().max()).item() <= 2e-2


@pytest.mark.parametrize("gain", [6.0, 40.0, 400.0, 2000.0])
def test_attention_scores_that_keep_growing(gain):
    """The softmax reference is only moved when a score outgrows it (the common sub-step takes no row maximum at all and
    notices growth from the exponent bits of P): keys whose scores climb along the sequence - by a little per sub-step, by
    more than 2^9 per sub-step, and by more than the f32 exponent range per sub-step - force the redo path again and again,
    for some rows only (the query sign flips per row), and the last key block must dominate exactly as in the reference."""
    S, H = 1000, 2
    g = torch.Generator().manual_seed(5)
    q = torch.zeros(S, H, 80)
    q[:, :, 0] = torch.where(torch.arange(S) % 3 == 0, -1.0, 1.0)[:, None] * (80 ** 0.5) / 1.4426950408889634  # score = +-k[.,0] in log2 units
    q[:, :, 1:] = torch.randn(S, H, 79, generator=g) * 0.05
    k = torch.randn(S, H, 80, generator=g) * 0.05
    k[:, :, 0] = (torch.arange(S) // 80).float()[:, None] * gain / 12.0 + torch.randn(S, H, generator=g) * 0.3   # climbs by `gain / 12` per 80-key tile
    v = torch.randn(S, H, 80, generator=g)
    q, k, v = (t.to(torch.bfloat16).cuda() for t in (q, k, v))
    out = gu.op_attention(gu.pack_qkv(q, k, v), [0, S], H)
    ref = gu.attention_reference(q, k, v, [0, S]).reshape(S, H * 80)
    assert torch.isfinite(out.float()).all()
    _close(out, ref, 2e-2)
