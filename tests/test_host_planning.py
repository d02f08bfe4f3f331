"""Host planning behind the C ABI (no GPU): libkocr.so loads, exports every symbol include/kocr.h declares, and its
integer / index work equals the oracle and the golden vectors bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from karanta_ocr_b200 import _lib, shard_pages, smart_resize
from karanta_ocr_b200.image_processor import KarantaImageProcessor
from oracle import preprocess_oracle as po
from oracle import vision_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kocr.h")).read()
    declared = sorted(set(re.findall(r"KOCR_API [a-z_0-9\* ]+?\b(kocr_[a-z_0-9]+)\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == declared
    assert b"sm_100a" in _lib.load().kocr_version()


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = _lib.load().kocr_create(0, C.byref(h))
    assert rc == _lib.ERR_CUDA and "no CPU fallback" in _lib.last_error()
    with pytest.raises(RuntimeError):
        KarantaImageProcessor()(images=[np.zeros((3, 56, 56), dtype=np.uint8)])


def test_smart_resize_golden_table():
    t = np.load(os.path.join(G, "g1_smart_resize.npz"))["table"]
    for h, w, mn, mx, hb, wb in t.tolist():
        if hb < 0:
            with pytest.raises(ValueError, match="absolute aspect ratio must be smaller than 200"):
                smart_resize(h, w, 28, mn, mx)
        else:
            assert smart_resize(h, w, 28, mn, mx) == (hb, wb), (h, w, mx)


@pytest.mark.parametrize("mode", [po.RESIZE_PIL, po.RESIZE_ATEN])
@pytest.mark.parametrize("in_size,out_size", [(995, 1008), (1288, 1288 - 28), (910, 896), (760, 756), (1024, 1036),
                                              (2048, 2044), (3000, 1484), (37, 28), (5000, 28), (28, 56), (100, 112)])
def test_filter_bank_equals_oracle(mode, in_size, out_size):
    lib = _lib.load()
    k = lib.kocr_resample_ksize(in_size, out_size)
    assert k == po.resample_ksize(in_size, out_size)
    b = np.zeros((out_size, 2), dtype=np.int32)
    c = np.zeros((out_size, k), dtype=np.int32)
    prec = C.c_int()
    assert lib.kocr_resample_coeffs(in_size, out_size, mode, b.ctypes.data, c.ctypes.data, C.byref(prec)) == 0
    ob, oc, oprec = po.resample_coeffs(in_size, out_size, mode)
    assert prec.value == oprec and np.array_equal(b, ob) and np.array_equal(c, oc)


@pytest.mark.parametrize("mode", [po.RESIZE_PIL, po.RESIZE_ATEN])
def test_normalize_lut_bit_exact(mode):
    lut = np.zeros((3, 256), dtype=np.float32)
    assert _lib.load().kocr_normalize_lut(mode, lut.ctypes.data) == 0
    assert np.array_equal(lut, po.normalize_lut(mode))


def test_index_work_golden():
    z = np.load(os.path.join(G, "g3_index_work.npz"))
    lib = _lib.load()
    for n in sorted({k.split(".")[0] for k in z.files}):
        g = np.ascontiguousarray(z[f"{n}.grid"].astype(np.int64))
        S = int(g.prod(-1).sum())
        pos = np.zeros((S, 2), dtype=np.int32)
        assert lib.kocr_pos_ids(g.ctypes.data, len(g), 2, pos.ctypes.data) == 0
        assert np.array_equal(pos, z[f"{n}.pos_ids"]), n
        cu = np.zeros(int(g[:, 0].sum()) + 1, dtype=np.int32)
        ncu = C.c_int()
        assert lib.kocr_cu_seqlens(g.ctypes.data, len(g), cu.ctypes.data, C.byref(ncu)) == 0
        assert np.array_equal(cu[: ncu.value], z[f"{n}.cu_seqlens"]), n
        wi = np.zeros(S // 4, dtype=np.int32)
        cuw = np.zeros(S // 4 + 2, dtype=np.int32)
        ncw = C.c_int()
        assert lib.kocr_window_index(g.ctypes.data, len(g), 112, 2, 14, wi.ctypes.data, cuw.ctypes.data, C.byref(ncw)) == 0
        assert np.array_equal(wi, z[f"{n}.window_index"]), n
        assert np.array_equal(cuw[: ncw.value], z[f"{n}.cu_window_seqlens"]), n


def test_index_work_random_grids_vs_oracle():
    rng = np.random.default_rng(3)
    lib = _lib.load()
    for _ in range(20):
        n = int(rng.integers(1, 6))
        g = np.stack([rng.integers(1, 3, n), rng.integers(1, 50, n) * 2, rng.integers(1, 50, n) * 2], -1).astype(np.int64)
        g = np.ascontiguousarray(g)
        S = int(g.prod(-1).sum())
        pos = np.zeros((S, 2), dtype=np.int32)
        lib.kocr_pos_ids(g.ctypes.data, n, 2, pos.ctypes.data)
        assert np.array_equal(pos, vo.pos_ids(g))
        wi = np.zeros(S // 4, dtype=np.int32)
        cuw = np.zeros(S // 4 + 2, dtype=np.int32)
        ncw = C.c_int()
        lib.kocr_window_index(g.ctypes.data, n, 112, 2, 14, wi.ctypes.data, cuw.ctypes.data, C.byref(ncw))
        owi, ocu = vo.window_index(g)
        assert np.array_equal(wi, owi) and np.array_equal(cuw[: ncw.value], ocu)


def test_processor_surface_without_gpu():
    p = KarantaImageProcessor(min_pixels=3136, max_pixels=12845056)
    assert p.get_number_of_image_patches(1288, 995) == 6624
    assert p.get_number_of_image_patches(1288, 995, {"max_pixels": 1003520}) == 4960
    assert p.model_input_names == ["pixel_values", "image_grid_thw"]
    assert (p.patch_size, p.temporal_patch_size, p.merge_size) == (14, 2, 2)
    with pytest.raises(ValueError):
        KarantaImageProcessor(size={"height": 10})
    with pytest.raises(ValueError):
        KarantaImageProcessor(patch_size=16)


def test_shard_pages_lpt():
    costs = [10, 1, 1, 1, 9, 2, 8, 3]
    shards = shard_pages(costs, 3)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 3
    assert shard_pages([5] * 64, 8) == [[r + 8 * k for k in range(8)] for r in range(8)]
    assert all(s == sorted(s) for s in shards)


@pytest.mark.parametrize("name,ocfg", [("qwen2_vl_7b", vo.qwen2_vl_7b()), ("qwen2_vl_2b", vo.qwen2_vl_2b()), ("qwen2_5_vl_7b", vo.qwen2_5_vl_7b())])
def test_presets_match_oracle_accounting(name, ocfg):
    """The product-side presets (what bench.py's GPU arm uses) agree with the oracle's configs, state-dict layout and
    FLOP formula on uniform and mixed batches."""
    from karanta_ocr_b200 import presets
    cfg = presets.preset(name)
    assert (cfg["arch"], cfg["depth"], cfg["embed_dim"], cfg["num_heads"], cfg["mlp_hidden"], cfg["out_hidden"]) == \
           (ocfg.arch, ocfg.depth, ocfg.embed_dim, ocfg.num_heads, ocfg.mlp_hidden, ocfg.out_hidden)
    small = presets.preset(name, depth=2)
    osmall = vo.TowerConfig(ocfg.arch, 2, ocfg.embed_dim, ocfg.num_heads, ocfg.mlp_hidden, ocfg.out_hidden)
    a, b = presets.random_state_dict(small, seed=1), vo.init_weights(osmall, seed=1)
    assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
    for grid in ([[1, 92, 72]] * 3, [[1, 92, 72], [1, 20, 18], [1, 46, 36], [1, 92, 64]]):
        got = presets.flops_per_batch(cfg, grid)
        assert got["total"] == pytest.approx(vo.flops_per_batch(ocfg, grid), rel=1e-12)
    if name == "qwen2_vl_7b":
        assert presets.flops_per_batch(cfg, [[1, 92, 72]])["total"] == pytest.approx(15.69e12, rel=2e-3)
