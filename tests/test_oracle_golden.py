"""The CPU oracle (oracle/) against the golden vectors minted from the third-party implementation
(tests/golden/make_golden.py). CPU only."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po
from oracle import vision_oracle as vo
from tests.synth import synth_page

G = os.path.join(os.path.dirname(__file__), "golden")
MODES = {"aten": po.RESIZE_ATEN, "pil": po.RESIZE_PIL}
CKPT_MAX = 12845056


def test_smart_resize_table():
    t = np.load(os.path.join(G, "g1_smart_resize.npz"))["table"]
    for h, w, mn, mx, hb, wb in t.tolist():
        try:
            got = po.smart_resize(h, w, 28, mn, mx)
        except ValueError:
            got = (-1, -1)
        assert got == (hb, wb), (h, w, mx)


def test_smart_resize_worked_examples():
    # SURVEY.md appendix A1
    assert po.smart_resize(1288, 995, 28, 3136, CKPT_MAX) == (1288, 1008)
    assert po.smart_resize(1288, 910, 28, 3136, CKPT_MAX) == (1288, 896)   # 32.5 -> 32 (half-even)
    assert po.smart_resize(1024, 760, 28, 3136, CKPT_MAX) == (1036, 756)
    assert po.smart_resize(1288, 995, 28, 3136, 1003520) == (1120, 868)
    with pytest.raises(ValueError):
        po.smart_resize(10, 2001)


@pytest.mark.parametrize("backend", ["aten", "pil"])
def test_pixel_values_small_exact(backend):
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    for name in ("coord_56x84", "noise_100x37", "noise_61x230", "noise_300x200", "page_256x256"):
        pv, grid = po.preprocess([z[f"{name}.image"]], 3136, CKPT_MAX, MODES[backend])
        assert (grid == z[f"{name}.{backend}.grid"]).all()
        assert pv.dtype == np.float32 and np.array_equal(pv, z[f"{name}.{backend}.pixel_values"]), name


def test_patch_order_map():
    """G2: 56x84 needs no resize, so pixel_values must be exactly lut[image] gathered in (A2) order."""
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    img = z["coord_56x84.image"]
    pv = z["coord_56x84.aten.pixel_values"]
    lut = po.normalize_lut(po.RESIZE_ATEN)
    gh, gw = 4, 6
    for n in range(gh * gw):
        cell, mh, mw = n // 4, (n % 4) // 2, n % 2
        r, c = 2 * (cell // (gw // 2)) + mh, 2 * (cell % (gw // 2)) + mw
        for f in (0, 13, 14, 195, 196, 392, 587, 588, 1175):
            ch, tp, py, px = f // 392, (f // 196) % 2, (f // 14) % 14, f % 14
            assert pv[n, f] == lut[ch][img[ch, r * 14 + py, c * 14 + px]]


@pytest.mark.parametrize("backend", ["aten", "pil"])
@pytest.mark.parametrize("name,shape,seed,maxp", [
    ("letter_1288x995", (1288, 995), 1234, CKPT_MAX),
    ("letter_1288x995_classmax", (1288, 995), 1234, 1003520),
    ("a4_1288x910", (1288, 910), 1235, CKPT_MAX),
    ("landscape_995x1288", (995, 1288), 1236, CKPT_MAX),
    ("column_1288x420", (1288, 420), 1237, CKPT_MAX),
])
def test_pixel_values_pages_crc(backend, name, shape, seed, maxp):
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    pv, grid = po.preprocess([synth_page(*shape, seed)], 3136, maxp, MODES[backend])
    assert (grid == z[f"{name}.{backend}.grid"]).all()
    assert zlib.crc32(pv.tobytes()) == int(z[f"{name}.{backend}.crc"])
    assert np.array_equal(pv.reshape(-1)[::1009], z[f"{name}.{backend}.sub"])


def test_mixed_batch_order():
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    pages = [synth_page(256, 256, 21), synth_page(640, 880, 22), synth_page(256, 256, 23), synth_page(308, 196, 24)]
    pv, grid = po.preprocess(pages, 3136, CKPT_MAX, po.RESIZE_ATEN)
    assert (grid == z["mixed.grid"]).all()
    assert zlib.crc32(pv.tobytes()) == int(z["mixed.crc"])


def test_index_work():
    z = np.load(os.path.join(G, "g3_index_work.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) == 8
    for n in names:
        g = z[f"{n}.grid"]
        assert np.array_equal(vo.pos_ids(g), z[f"{n}.pos_ids"]), n
        assert np.array_equal(vo.cu_seqlens(g), z[f"{n}.cu_seqlens"]) and vo.cu_seqlens(g).dtype == np.int32
        wi, cuw = vo.window_index(g)
        assert np.array_equal(wi, z[f"{n}.window_index"]), n
        assert np.array_equal(cuw, z[f"{n}.cu_window_seqlens"]), n
        cos, sin = vo.rope_cos_sin(g, 80)
        rot = z[f"{n}.rotary_f32"]
        stride = max(1, cos.shape[0] // 64)
        ref = torch.from_numpy(rot)
        assert torch.equal(cos[::stride, :40], ref.cos()) and torch.equal(sin[::stride, 40:], ref.sin())


@pytest.mark.parametrize("name,cfg", [
    ("tiny_q2", vo.TowerConfig("qwen2_vl", 2, 160, 2, 640, 256)),
    ("tiny_q25", vo.TowerConfig("qwen2_5_vl", 3, 160, 2, 428, 256, fullatt_block_indexes=(1,))),
    ("mid_q2_d2", vo.TowerConfig("qwen2_vl", 2, 1280, 16, 5120, 1536)),
    ("mid_q25_d2", vo.TowerConfig("qwen2_5_vl", 2, 1280, 16, 3420, 2048, fullatt_block_indexes=(1,))),
])
def test_tower_fp32_vs_hf(name, cfg):
    z = np.load(os.path.join(G, "g5_embeddings.npz"))
    pages = [z[f"{name}.page{i}"] for i in range(8) if f"{name}.page{i}" in z.files]
    pv, grid = po.preprocess(pages, 3136, CKPT_MAX, po.RESIZE_ATEN)
    assert (grid == z[f"{name}.grid"]).all()
    out = vo.tower_forward(cfg, vo.init_weights(cfg, seed=100), torch.from_numpy(pv), grid).numpy()
    ref = z[f"{name}.emb"]
    assert out.shape == ref.shape
    scale = float(z[f"{name}.emb_absmax"])
    assert np.abs(out - ref).max() <= 2e-5 * scale, np.abs(out - ref).max() / scale


def test_flops_formula_matches_survey():
    # 15.69 TFLOP per letter page (Qwen2-VL-7B), 9.47 (Qwen2.5-VL-7B)
    assert abs(vo.flops_per_batch(vo.qwen2_vl_7b(), [[1, 92, 72]]) / 1e12 - 15.69) < 0.02
    assert abs(vo.flops_per_batch(vo.qwen2_5_vl_7b(), [[1, 92, 72]]) / 1e12 - 9.47) < 0.05
