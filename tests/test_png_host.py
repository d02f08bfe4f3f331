"""SURVEY.md section 8 row f2, host half: the DEFLATE state machine and the PNG scan-line reconstruction the CUDA decode
kernels execute (karanta_ocr_b200/csrc/kocr_inflate_core.h is compiled for host and device alike) run here on the CPU,
bit-exact against zlib and Pillow - the decoder inside the reference's own page decode (karanta/data/utils.py:186-225).
Also the host-side PNG container parsing of libkocr.so (kocr_png_info). The GPU kernels are tested in test_gpu_png.py."""
import ctypes as C
import io
import os
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

from tests.synth import synth_page

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "inflate_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "inflate_host.cpp")])
    lib = C.CDLL(so)
    lib.kocr_test_inflate.restype = C.c_int
    lib.kocr_test_inflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.kocr_test_unfilter.restype = C.c_int
    lib.kocr_test_unfilter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def _inflate(core, comp: bytes, n: int):
    src = np.frombuffer(comp, dtype=np.uint8).copy()
    out = np.zeros(max(n, 1), dtype=np.uint8)
    got = C.c_int64()
    st = core.kocr_test_inflate(src.ctypes.data, len(src), out.ctypes.data, n, C.byref(got))
    return st, out[:n], got.value


def _streams():
    rng = np.random.default_rng(5)
    page = synth_page(300, 420, 3).transpose(1, 2, 0).tobytes()
    yield "empty", b""
    yield "one byte", b"x"
    yield "zeros (long matches, distance 1)", bytes(70000)
    yield "noise (stored / literal only)", rng.integers(0, 256, 50000, dtype=np.uint8).tobytes()
    yield "page", page
    yield "text", (b"the quick brown fox jumps over the lazy dog. " * 3000)[:100001]
    yield "low entropy", rng.integers(0, 4, 200000, dtype=np.uint8).tobytes()
    yield "skewed (codes longer than the fast table)", rng.choice(256, 120000, p=np.r_[[0.5], np.full(255, 0.5 / 255)]).astype(np.uint8).tobytes()
    yield "period 3", bytes([1, 2, 3]) * 30000


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_core_equals_zlib(core, level):
    for name, data in _streams():
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
            comp = co.compress(data) + co.flush()
            st, out, got = _inflate(core, comp, len(data))
            assert st == 0 and got == len(data), (name, level, strategy, st, got)
            assert out.tobytes() == data, (name, level, strategy)


def test_inflate_core_multi_block_and_flush_points(core):
    data = synth_page(200, 600, 9).tobytes()
    co = zlib.compressobj(6)
    comp = b""
    for i in range(0, len(data), 7001):     # sync flushes insert empty stored blocks; full flushes reset the dictionary
        comp += co.compress(data[i:i + 7001]) + co.flush(zlib.Z_SYNC_FLUSH if (i // 7001) % 3 else zlib.Z_FULL_FLUSH)
    comp += co.flush()
    st, out, got = _inflate(core, comp, len(data))
    assert st == 0 and out.tobytes() == data


def test_inflate_core_rejects_damage(core):
    data = synth_page(120, 200, 2).tobytes()
    comp = bytearray(zlib.compress(data, 6))
    st, _, _ = _inflate(core, bytes(comp[:len(comp) // 2]), len(data))       # truncated
    assert st != 0
    st, _, _ = _inflate(core, bytes(comp), len(data) - 10)                    # more output than the image holds
    assert st != 0
    st, _, _ = _inflate(core, bytes(comp), len(data) + 10)                    # less
    assert st != 0
    bad = bytearray(comp)
    bad[0] = 0x79                                                             # not deflate
    assert _inflate(core, bytes(bad), len(data))[0] != 0
    rng = np.random.default_rng(0)
    for _ in range(200):                                                      # random corruption never crashes or overruns
        bad = bytearray(comp)
        for k in rng.integers(2, len(bad), 3):
            bad[k] ^= 1 << int(rng.integers(0, 8))
        st, out, got = _inflate(core, bytes(bad), len(data))
        assert 0 <= got <= len(data)


def _png_raw(img: Image.Image, **save_kw):
    """-> (IDAT payload inflated by zlib = filtered scan lines, decoded pixels by Pillow)."""
    buf = io.BytesIO()
    img.save(buf, format="PNG", **save_kw)
    b = buf.getvalue()
    pos, idat = 8, b""
    while pos < len(b):
        n = int.from_bytes(b[pos:pos + 4], "big")
        if b[pos + 4:pos + 8] == b"IDAT":
            idat += b[pos + 8:pos + 8 + n]
        pos += 12 + n
    return b, zlib.decompress(idat), np.asarray(Image.open(io.BytesIO(b)))


@pytest.mark.parametrize("mode,bpp,out_ch", [("L", 1, 1), ("RGB", 3, 3), ("LA", 2, 1), ("RGBA", 4, 3)])
def test_unfilter_core_equals_pillow(core, mode, bpp, out_ch):
    rng = np.random.default_rng(11)
    base = synth_page(97, 131, 4).transpose(1, 2, 0)
    photo = np.asarray(Image.open(os.path.join(HERE, "golden", "sample_760x1024.png")).convert("RGB").crop((100, 200, 331, 297)))
    for arr in (base, photo, rng.integers(0, 256, (40, 33, 3), dtype=np.uint8)):
        img = Image.fromarray(arr).convert(mode)
        for kw in ({}, {"optimize": True}, {"compress_level": 1}):
            _, raw, want = _png_raw(img, **kw)
            h, w = want.shape[:2]
            rawa = np.frombuffer(raw, dtype=np.uint8).copy()
            assert len(rawa) == h * (1 + w * bpp)
            out = np.zeros((h, w, out_ch), dtype=np.uint8)
            assert core.kocr_test_unfilter(rawa.ctypes.data, h, w, bpp, out_ch, out.ctypes.data) == 0
            ref = want.reshape(h, w, -1)[:, :, :out_ch]
            assert np.array_equal(out, ref), (mode, kw)
    # every filter type really occurred in these fixtures
    _, raw, want = _png_raw(Image.fromarray(photo))
    types = set(np.frombuffer(raw, dtype=np.uint8).reshape(want.shape[0], -1)[:, 0].tolist())
    assert types >= {1, 2, 4} or len(types) >= 3, types
