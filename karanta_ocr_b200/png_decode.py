"""Page decode on the GPU: PNG file bytes -> uint8 pixels in HBM, ready for the preprocess kernel.

Stands in for the host decode the reference does in front of the path - `Image.open(BytesIO(base64.b64decode(...)))` in
karanta/data/utils.py:186-225 (base64_to_grayscale), :228-251 (prepare_image_and_text) and on the serving side of
karanta/pipeline.py:131-142 - for the PNG flavours the reference produces (8-bit gray 'L' and RGB, also with alpha, not
interlaced). The container is walked on the host, DEFLATE and the scan-line filters run in libkocr.so
(csrc/kocr_png.cu). Pages come back as device tensors [H, W] (gray) or [H, W, 3] that KarantaImageProcessor /
PageEncoder accept as they are, so a page goes from its base64 payload to embeddings without its pixels touching the host.
"""
from __future__ import annotations

import base64
import ctypes as C

import torch

from . import _lib

_STATUS = {1: "bad zlib / deflate block header", 2: "invalid Huffman code", 3: "match distance reaches before the start of the image",
           4: "more pixel data than the image holds", 5: "compressed data ends early", 6: "unknown scan-line filter type"}


class PngError(ValueError):
    """A page could not be decoded; `.index` is its position in the batch."""

    def __init__(self, index, msg):
        super().__init__(f"page {index}: {msg}")
        self.index = index


def payload_bytes(url_or_b64) -> bytes:
    """`data:image/png;base64,<payload>`, a bare base64 string, or raw bytes -> file bytes."""
    if isinstance(url_or_b64, (bytes, bytearray, memoryview)):
        return bytes(url_or_b64)
    s = url_or_b64.split(",", 1)[1] if url_or_b64.startswith("data:") else url_or_b64
    return base64.b64decode(s)


def png_info(data: bytes) -> _lib.KocrPngInfo:
    """Header of a PNG file (host only). ValueError for a damaged file, RuntimeError for flavours the GPU path does not decode."""
    info = _lib.KocrPngInfo()
    _lib.check(_lib.load().kocr_png_info(data, len(data), C.byref(info)))
    return info


def is_gpu_decodable(data: bytes) -> bool:
    try:
        png_info(data)
        return True
    except (ValueError, RuntimeError):
        return False


@torch.no_grad()
def decode_png_batch(files, device=None, check: bool = True):
    """files: list of PNG file bytes -> list of uint8 CUDA tensors, [H, W] for gray pages and [H, W, 3] for colour ones.
    With `check` (default) the stream is synchronised and a page whose compressed data is damaged raises PngError; pass
    check=False to keep the call asynchronous and read `decode_png_batch.last_status` yourself."""
    if not torch.cuda.is_available():
        raise RuntimeError("decode_png_batch needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    files = [bytes(f) for f in files]
    n = len(files)
    if n == 0:
        return []
    lib = _lib.load()
    infos = (_lib.KocrPngInfo * n)()
    for i, f in enumerate(files):
        try:
            _lib.check(lib.kocr_png_info(f, len(f), C.byref(infos[i])))
        except (ValueError, RuntimeError) as e:
            raise PngError(i, str(e)) from e
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        outs = [torch.empty((inf.height, inf.width) if inf.channels == 1 else (inf.height, inf.width, 3), dtype=torch.uint8, device=dev)
                for inf in infos]
        scratch = torch.empty(_lib.check(lib.kocr_png_scratch_bytes(infos, n)), dtype=torch.uint8, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        fptr = (C.c_char_p * n)(*files)
        sizes = (C.c_int64 * n)(*[len(f) for f in files])
        optr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        _lib.check(lib.kocr_png_decode(_lib.context(dev.index), fptr, sizes, n, optr, scratch.data_ptr(), scratch.numel(),
                                       status.data_ptr(), stream.cuda_stream))
        scratch.record_stream(stream)
        decode_png_batch.last_status = status
        if check:
            st = status.cpu().tolist()  # synchronises the stream
            for i, v in enumerate(st):
                if v:
                    raise PngError(i, _STATUS.get(v, f"decode status {v}"))
    return outs
