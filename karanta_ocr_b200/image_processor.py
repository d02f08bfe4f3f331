"""KarantaImageProcessor: the Qwen2-VL image-processor call surface on one fused sm_100a kernel.

Mirrors `Qwen2VLImageProcessor` (transformers models/qwen2_vl/image_processing_qwen2_vl.py:92-261), the object
karanta-ocr reaches through `AutoProcessor` at karanta/training/pipeline_steps.py:289-294,
karanta/training/data.py:188,208 and karanta/training/test_trained_model.py:25-31,82-87:

    proc = KarantaImageProcessor(min_pixels=3136, max_pixels=12845056)
    out = proc(images=[pil_page, ...], return_tensors="pt")
    out["pixel_values"]    # float32 [sum N, 1176]
    out["image_grid_thw"]  # int64   [n, 3]

Same names, argument meaning and errors (ValueError when the aspect ratio exceeds 200). All arithmetic runs in
libkocr.so on the GPU; `resize_backend` picks which upstream fixed-point resize to reproduce bit-for-bit:
"torchvision" (transformers 5.x default backend) or "pil" (the 4.53.3 slow processor the reference pins).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

OPENAI_CLIP_MEAN = [0.48145466, 0.4578275, 0.40821073]
OPENAI_CLIP_STD = [0.26862954, 0.26130258, 0.27577711]


class _BatchFeature(dict):
    """Minimal stand-in used when transformers cannot be imported (same item / attribute access)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *a, **kw):
        return _BatchFeature({k: v.to(*a, **kw) if hasattr(v, "to") else v for k, v in self.items()})


def _batch_feature(data):
    try:
        from transformers.feature_extraction_utils import BatchFeature
        return BatchFeature(data=data)
    except Exception:  # transformers absent or broken: keep the mapping interface
        return _BatchFeature(data)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                 max_pixels: int = 14 * 14 * 4 * 1280):
    """transformers image_processing_qwen2_vl.py:62-88, evaluated by libkocr's host planner."""
    h, w = C.c_int(), C.c_int()
    _lib.check(_lib.load().kocr_smart_resize(int(height), int(width), int(factor), int(min_pixels), int(max_pixels),
                                             C.byref(h), C.byref(w)))
    return h.value, w.value


def _as_u8_page(image):
    """-> (uint8 tensor, height, width, layout). do_convert_rgb semantics: gray pages become 3 equal channels
    (inside the kernel), alpha is dropped."""
    if hasattr(image, "convert") and hasattr(image, "mode"):  # PIL
        if image.mode == "L":
            a = np.array(image)  # a writable copy: torch refuses to wrap PIL's read-only buffer without a warning
            return torch.from_numpy(a), a.shape[0], a.shape[1], _lib.LAYOUT_GRAY
        a = np.array(image.convert("RGB"))
        return torch.from_numpy(a), a.shape[0], a.shape[1], _lib.LAYOUT_HWC
    t = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(image)))
    if t.dtype != torch.uint8:
        raise ValueError(f"page images must be uint8 (got {t.dtype}); rescale before the processor is not supported")
    if t.ndim == 2:
        return t.contiguous(), t.shape[0], t.shape[1], _lib.LAYOUT_GRAY
    if t.ndim != 3:
        raise ValueError(f"expected a 2-D or 3-D image, got shape {tuple(t.shape)}")
    if t.shape[0] in (1, 3, 4) and t.shape[-1] not in (1, 3, 4):
        if t.shape[0] == 1:
            return t[0].contiguous(), t.shape[1], t.shape[2], _lib.LAYOUT_GRAY
        return t[:3].contiguous(), t.shape[1], t.shape[2], _lib.LAYOUT_CHW
    if t.shape[-1] in (1, 3, 4):
        if t.shape[-1] == 1:
            return t[..., 0].contiguous(), t.shape[0], t.shape[1], _lib.LAYOUT_GRAY
        return t[..., :3].contiguous(), t.shape[0], t.shape[1], _lib.LAYOUT_HWC
    raise ValueError(f"cannot infer the channel dimension of an image of shape {tuple(t.shape)}")


class KarantaImageProcessor:
    model_input_names = ["pixel_values", "image_grid_thw"]
    do_resize = True
    do_rescale = True
    do_normalize = True
    do_convert_rgb = True
    image_mean = OPENAI_CLIP_MEAN
    image_std = OPENAI_CLIP_STD
    patch_size = 14
    temporal_patch_size = 2
    merge_size = 2

    def __init__(self, min_pixels: int | None = None, max_pixels: int | None = None, size: dict | None = None,
                 resize_backend: str = "torchvision", device: str | torch.device | None = None, **kwargs):
        size = dict(size) if size is not None else {"shortest_edge": 56 * 56, "longest_edge": 28 * 28 * 1280}
        if min_pixels is not None:
            size["shortest_edge"] = min_pixels
        if max_pixels is not None:
            size["longest_edge"] = max_pixels
        if "shortest_edge" not in size or "longest_edge" not in size:
            raise ValueError("size must contain 'shortest_edge' and 'longest_edge' keys.")
        self._check_fixed_kwargs(kwargs)
        if resize_backend not in ("torchvision", "pil"):
            raise ValueError("resize_backend must be 'torchvision' or 'pil'")
        self.size = size
        self.resize_mode = _lib.RESIZE_ATEN if resize_backend == "torchvision" else _lib.RESIZE_PIL
        self.resize_backend = resize_backend
        self.device = torch.device(device) if device is not None else None
        self._pinned = None
        self._pinned_ev = None

    @classmethod
    def _check_fixed_kwargs(cls, kwargs):
        """Options the kernel is built around: a different value is an error, never silently ignored."""
        for k in ("patch_size", "temporal_patch_size", "merge_size"):
            if kwargs.get(k) is not None and kwargs[k] != getattr(cls, k):
                raise ValueError(f"{k}={kwargs[k]} is not supported: the kernel is built for 14 / 2 / 2")
        for k, ref in (("image_mean", OPENAI_CLIP_MEAN), ("image_std", OPENAI_CLIP_STD)):
            if kwargs.get(k) is not None and [float(v) for v in kwargs[k]] != ref:
                raise ValueError(f"{k} other than the OPENAI_CLIP constants is not supported")
        if kwargs.get("resample") is not None and int(kwargs["resample"]) != 3:
            raise ValueError("resample other than BICUBIC (3) is not supported")
        if kwargs.get("rescale_factor") is not None and float(kwargs["rescale_factor"]) != 1 / 255:
            raise ValueError("rescale_factor other than 1/255 is not supported")

    @property
    def min_pixels(self):
        return self.size["shortest_edge"]

    @property
    def max_pixels(self):
        return self.size["longest_edge"]

    def get_number_of_image_patches(self, height: int, width: int, images_kwargs=None):
        """transformers image_processing_qwen2_vl.py:234-261 (vLLM sizes its placeholders with it)."""
        kw = images_kwargs or {}
        n = _lib.load().kocr_num_patches(int(height), int(width), kw.get("patch_size", self.patch_size),
                                         kw.get("merge_size", self.merge_size), int(kw.get("min_pixels", self.min_pixels)),
                                         int(kw.get("max_pixels", self.max_pixels)))
        return _lib.check(n)

    def __call__(self, images, **kwargs):
        return self.preprocess(images, **kwargs)

    # ------------------------------------------------------------------ device path
    def _cuda_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("KarantaImageProcessor needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = self.device if self.device is not None and self.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())

    def _upload(self, pages, dev):
        """Pack the host pages into one pinned buffer -> one H2D copy. Returns per-page device pointers + keep-alives."""
        host = [(i, p) for i, p in enumerate(pages) if p[0].device.type == "cpu"]
        ptrs = [None] * len(pages)
        keep = []
        h2d = 0
        if host:
            offs, total = [], 0
            for _, p in host:
                offs.append(total)
                total += (p[0].numel() + 255) // 256 * 256
            dbuf = torch.empty(total, dtype=torch.uint8, device=dev)
            if all(p[0].is_pinned() for _, p in host):
                # already page-locked (bulk encode keeps its pages pinned): copy each page straight into the packed buffer
                for (i, p), o in zip(host, offs):
                    dbuf[o:o + p[0].numel()].copy_(p[0].reshape(-1), non_blocking=True)
            else:
                if self._pinned is None or self._pinned.numel() < total:
                    self._pinned = torch.empty(max(total, 1 << 20), dtype=torch.uint8, pin_memory=True)
                    self._pinned_ev = None
                if self._pinned_ev is not None:
                    self._pinned_ev.synchronize()
                for (i, p), o in zip(host, offs):
                    self._pinned[o:o + p[0].numel()].copy_(p[0].reshape(-1))
                dbuf.copy_(self._pinned[:total], non_blocking=True)
                self._pinned_ev = torch.cuda.Event()
                self._pinned_ev.record(torch.cuda.current_stream(dev))
            keep.append(dbuf)
            for (i, p), o in zip(host, offs):
                ptrs[i] = dbuf.data_ptr() + o
            h2d = total
        for i, p in enumerate(pages):
            if ptrs[i] is None:
                t = p[0].to(dev) if p[0].device != dev else p[0]
                keep.append(t)
                ptrs[i] = t.data_ptr()
        return ptrs, keep, h2d

    def preprocess_device(self, images, out_dtype=torch.float32, min_pixels=None, max_pixels=None, out=None):
        """pages -> (pixel_values on the GPU [sum N, 1176] in out_dtype, image_grid_thw int64 CPU [n, 3])."""
        if not isinstance(images, (list, tuple)):
            images = [images]
        if len(images) == 0:
            raise ValueError("images is empty")
        minp = int(self.min_pixels if min_pixels is None else min_pixels)
        maxp = int(self.max_pixels if max_pixels is None else max_pixels)
        dev = self._cuda_device()
        pages = [_as_u8_page(im) for im in images]
        n_rows = 0
        for _, h, w, _ in pages:
            rh, rw = smart_resize(h, w, self.patch_size * self.merge_size, minp, maxp)  # raises ValueError on aspect > 200
            n_rows += (rh // self.patch_size) * (rw // self.patch_size)
        with torch.cuda.device(dev):
            ptrs, keep, h2d = self._upload(pages, dev)
            arr = (_lib.KocrImage * len(pages))()
            for i, (_, h, w, layout) in enumerate(pages):
                arr[i].data, arr[i].height, arr[i].width, arr[i].layout = ptrs[i], h, w, layout
            patch_dim = 3 * self.temporal_patch_size * self.patch_size * self.patch_size
            if out is None:
                out = torch.empty((n_rows, patch_dim), dtype=out_dtype, device=dev)
            elif out.shape[0] < n_rows or out.shape[1] != patch_dim or out.dtype != out_dtype or not out.is_contiguous():
                raise ValueError("out buffer has the wrong shape / dtype")
            dt = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}.get(out_dtype)
            if dt is None:
                raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
            grid = np.zeros((len(pages), 3), dtype=np.int64)
            rc = _lib.load().kocr_preprocess(_lib.context(dev.index), arr, len(pages), minp, maxp, self.resize_mode, dt,
                                             out.data_ptr(), out.shape[0], grid.ctypes.data,
                                             torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc)
            for k in keep:  # inputs must outlive the kernel that is still queued on the stream
                k.record_stream(torch.cuda.current_stream(dev))
        self.last_h2d_bytes = h2d
        return out[:n_rows], torch.from_numpy(grid)

    def preprocess(self, images, return_tensors="pt", min_pixels=None, max_pixels=None, device=None, **kwargs):
        """Qwen2VLImageProcessor.preprocess: returns a BatchFeature with float32 pixel_values and int64 image_grid_thw.
        Tensors land on the CPU like the transformers processor's do, unless `device` (or the constructor's) says cuda."""
        for k in ("do_resize", "do_rescale", "do_normalize", "do_convert_rgb"):
            if kwargs.get(k, True) is False:
                raise ValueError(f"{k}=False is not supported by the fused kernel")
        if kwargs.get("videos") is not None:
            raise ValueError("video input is not supported on this path (karanta-ocr sends still pages)")
        self._check_fixed_kwargs(kwargs)
        size = kwargs.get("size")
        if size is not None:  # call-time size= has the constructor's meaning (HF image_processing_qwen2_vl.py:148-166)
            if "shortest_edge" not in size or "longest_edge" not in size:
                raise ValueError("size must contain 'shortest_edge' and 'longest_edge' keys.")
            min_pixels = size["shortest_edge"] if min_pixels is None else min_pixels
            max_pixels = size["longest_edge"] if max_pixels is None else max_pixels
        pv, grid = self.preprocess_device(images, torch.float32, min_pixels, max_pixels)
        target = torch.device(device) if device is not None else (self.device or torch.device("cpu"))
        if target.type == "cpu":
            pv = pv.cpu()
        else:
            grid = grid.to(target)
        if return_tensors in ("np", "numpy"):
            return _batch_feature({"pixel_values": pv.cpu().numpy(), "image_grid_thw": grid.cpu().numpy()})
        if return_tensors not in ("pt", None):
            raise ValueError("return_tensors must be 'pt' or 'np'")
        return _batch_feature({"pixel_values": pv, "image_grid_thw": grid})
