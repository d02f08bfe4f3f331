"""CPU oracle for the page-image preprocess path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the algorithm the reference reaches through its
third-party dependency `transformers` (pinned 4.53.3 in /root/reference/uv.lock:2168-2169,
call sites /root/reference/karanta/training/pipeline_steps.py:289-294 and
/root/reference/karanta/training/test_trained_model.py:82-87).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  The product path (`karanta_ocr_b200`) never does.

Restated functions and the upstream lines they follow (HF = site-packages/transformers 5.5.0,
TV = torchvision 0.26.0, PIL = Pillow 12.2 `src/libImaging/Resample.c`, ATen =
`aten/src/ATen/native/cpu/UpSampleKernel.cpp` of torch 2.11):

  smart_resize            HF models/qwen2_vl/image_processing_qwen2_vl.py:62-88
  resample_coeffs(PIL)    PIL Resample.c precompute_coeffs + normalize_coeffs_8bpc
  resample_coeffs(ATEN)   ATen HelperInterpBase::_compute_indices_min_size_weights_aa and
                          HelperInterpCubic::compute_index_ranges_int16_weights
  resize_u8               PIL ImagingResampleHorizontal_8bpc / Vertical_8bpc; ATen
                          basic_loop_aa_horizontal<uint8_t> / vertical (same integer arithmetic,
                          horizontal pass first, uint8 intermediate)
  normalize               HF image_processing_backends.py:291-331 (fused mean*255, std*255, f32 sub/div)
  patchify                HF models/qwen2_vl/image_processing_qwen2_vl.py:194-220
  preprocess              HF models/qwen2_vl/image_processing_qwen2_vl.py:148-232

Parity status: the reference's own tests hold no vector for this path, so
the oracle is pinned against outputs of the third-party implementation itself, generated in the
build container by tests/golden/make_golden.py and committed under tests/golden/.
"""
from __future__ import annotations

import math

import numpy as np

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

RESIZE_PIL = 0    # Pillow 8bpc fixed point, 22-bit int32 coefficients (transformers "pil" backend; 4.53.3 slow processor)
RESIZE_ATEN = 1   # ATen uint8 fixed point, int16 coefficients with dynamic precision (transformers 5.x torchvision backend on CPU)

PIL_PRECISION_BITS = 32 - 8 - 2


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                 max_pixels: int = 14 * 14 * 4 * 1280) -> tuple[int, int]:
    """HF image_processing_qwen2_vl.py:62-88. Python round() is round-half-even on a float64 quotient."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(
            f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}"
        )
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def _bicubic_pil(x: float) -> float:
    # PIL Resample.c bicubic_filter, a = -0.5
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bicubic_aten(x: float) -> float:
    # ATen UpSample.h cubic_convolution1 / cubic_convolution2 via HelperInterpCubic::aa_filter, a = -0.5
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return 0.0


def resample_ksize(in_size: int, out_size: int) -> int:
    scale = in_size / out_size
    support = 2.0 * max(scale, 1.0)
    return int(math.ceil(support)) * 2 + 1


def resample_coeffs(in_size: int, out_size: int, mode: int):
    """Returns (bounds[out,2] int32 = (xmin, xcount), coeffs[out,ksize] int32, precision_bits).

    Both upstream implementations compute the float64 weights the same way (window centred at
    scale*(i+0.5), support 2*max(scale,1), weights normalised to sum 1); they differ in the cubic's
    algebraic form and in the fixed-point conversion.
    """
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    filt = _bicubic_pil if mode == RESIZE_PIL else _bicubic_aten
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        if mode == RESIZE_PIL:
            center = 0 + (xx + 0.5) * scale
        else:
            center = scale * (xx + 0.5)
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xcount = xmax - xmin
        if mode == RESIZE_ATEN:
            xcount = min(max(xcount, 0), ksize)
        ww = 0.0
        ws = []
        for x in range(xcount):
            w = filt((x + xmin - center + 0.5) * inv)
            ws.append(w)
            ww += w
        for x in range(xcount):
            kk[xx, x] = ws[x] / ww if ww != 0.0 else ws[x]
        bounds[xx, 0] = xmin
        bounds[xx, 1] = xcount
    if mode == RESIZE_PIL:
        prec = PIL_PRECISION_BITS
    else:
        wt_max = float(kk.max()) if kk.size else 0.0
        prec = 0
        while prec < 22:
            next_value = int(0.5 + wt_max * (1 << (prec + 1)))
            if next_value >= (1 << 15):
                break
            prec += 1
    scaled = kk * float(1 << prec)
    coeffs = np.where(scaled < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int64)
    if mode == RESIZE_ATEN:
        coeffs = coeffs.astype(np.int16).astype(np.int64)
    return bounds, coeffs.astype(np.int32), prec


def _resample_last_axis(img: np.ndarray, bounds: np.ndarray, coeffs: np.ndarray, prec: int) -> np.ndarray:
    """uint8 [..., in] -> uint8 [..., out]: ss = 2^(prec-1) + sum(pix*coef); out = clip(ss >> prec, 0, 255)."""
    out_size, ksize = coeffs.shape
    in_size = img.shape[-1]
    idx = bounds[:, 0:1].astype(np.int64) + np.arange(ksize, dtype=np.int64)[None, :]
    valid = np.arange(ksize)[None, :] < bounds[:, 1:2]
    idx = np.where(valid, idx, 0).clip(0, in_size - 1)
    w = np.where(valid, coeffs, 0).astype(np.int64)
    acc = np.full(img.shape[:-1] + (out_size,), 1 << (prec - 1), dtype=np.int64)
    src = img.astype(np.int64)
    for k in range(ksize):
        acc += src[..., idx[:, k]] * w[:, k]
    return np.clip(acc >> prec, 0, 255).astype(np.uint8)


def resize_u8(img_chw: np.ndarray, out_h: int, out_w: int, mode: int = RESIZE_PIL) -> np.ndarray:
    """Bicubic antialiased resize of a uint8 CHW image; horizontal pass first, uint8 intermediate.
    A pass whose size does not change is skipped (PIL need_horizontal/need_vertical; ATen same)."""
    assert img_chw.dtype == np.uint8 and img_chw.ndim == 3
    _, in_h, in_w = img_chw.shape
    x = img_chw
    if out_w != in_w:
        b, c, p = resample_coeffs(in_w, out_w, mode)
        x = _resample_last_axis(x, b, c, p)
    if out_h != in_h:
        b, c, p = resample_coeffs(in_h, out_h, mode)
        x = np.ascontiguousarray(
            _resample_last_axis(np.ascontiguousarray(x.transpose(0, 2, 1)), b, c, p).transpose(0, 2, 1))
    return x


def normalize_lut(mode: int = RESIZE_ATEN) -> np.ndarray:
    """The 3x256 float32 values a uint8 level maps to.

    RESIZE_ATEN (torchvision backend, HF image_processing_backends.py:291-331): mean/std f32 tensors times 255.0,
    then one f32 subtract and one f32 divide.
    RESIZE_PIL (pil backend, HF image_transforms.py rescale:118-122 then normalize:417-439): f32(f64(x) * (1/255)),
    then f32 subtract of the f32 mean and f32 divide by the f32 std.
    """
    lv = np.arange(256)
    mean = np.asarray(OPENAI_CLIP_MEAN, dtype=np.float32)
    std = np.asarray(OPENAI_CLIP_STD, dtype=np.float32)
    if mode == RESIZE_ATEN:
        m = (mean * np.float32(1.0 / (1 / 255))).astype(np.float32)
        s = (std * np.float32(1.0 / (1 / 255))).astype(np.float32)
        x = lv.astype(np.float32)[None, :]
    else:
        m, s = mean, std
        x = (lv.astype(np.float64) * (1 / 255)).astype(np.float32)[None, :]
    return ((x - m[:, None]) / s[:, None]).astype(np.float32)


def normalize_f32(img_u8_chw: np.ndarray, mode: int = RESIZE_ATEN) -> np.ndarray:
    lut = normalize_lut(mode)
    return np.stack([lut[c][img_u8_chw[c]] for c in range(3)], axis=0)


def patchify(img_f32_chw: np.ndarray, patch: int = 14, tps: int = 2, merge: int = 2):
    """HF image_processing_qwen2_vl.py:194-220 for one still image -> ([N, C*tps*patch*patch], (1, gh, gw))."""
    c, h, w = img_f32_chw.shape
    gh, gw = h // patch, w // patch
    x = np.broadcast_to(img_f32_chw[None], (tps, c, h, w))  # temporal pad: repeat the last frame
    x = x.reshape(1, tps, c, gh // merge, merge, patch, gw // merge, merge, patch)
    # axes now: (gt, tps, c, hb, mh, py, wb, mw, px) -> (gt, hb, wb, mh, mw, c, tps, py, px)
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return np.ascontiguousarray(x.reshape(gh * gw, c * tps * patch * patch)), (1, gh, gw)


def to_chw_u8(image) -> np.ndarray:
    """do_convert_rgb + channels-first: accepts PIL, HWC/CHW/HW uint8 arrays."""
    if hasattr(image, "convert"):
        image = np.asarray(image.convert("RGB"))
    a = np.asarray(image)
    if a.ndim == 2:
        a = np.stack([a, a, a], axis=0)
    elif a.shape[-1] in (1, 3, 4) and a.shape[0] not in (1, 3):
        a = a.transpose(2, 0, 1)
    if a.shape[0] == 1:
        a = np.repeat(a, 3, axis=0)
    if a.shape[0] == 4:
        a = a[:3]
    assert a.dtype == np.uint8 and a.shape[0] == 3
    return np.ascontiguousarray(a)


def preprocess(images, min_pixels: int = 56 * 56, max_pixels: int = 28 * 28 * 1280, mode: int = RESIZE_PIL,
               patch: int = 14, tps: int = 2, merge: int = 2):
    """HF Qwen2VLImageProcessor._preprocess: list of pages -> (pixel_values f32 [sumN,1176], grid_thw i64 [n,3])."""
    if not isinstance(images, (list, tuple)):
        images = [images]
    pv, grids = [], []
    for im in images:
        a = to_chw_u8(im)
        _, h, w = a.shape
        rh, rw = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
        r = resize_u8(a, rh, rw, mode)
        p, g = patchify(normalize_f32(r, mode), patch, tps, merge)
        pv.append(p)
        grids.append(g)
    return np.concatenate(pv, axis=0), np.asarray(grids, dtype=np.int64)
