"""GPU parity: the fused preprocess kernel (through the C ABI / KarantaImageProcessor) against the oracle and the
golden vectors. Integer/byte work: bit-exact."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po
from tests.synth import synth_page

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CKPT_MAX = 12845056
MODES = {"torchvision": po.RESIZE_ATEN, "pil": po.RESIZE_PIL}
GOLD = {"torchvision": "aten", "pil": "pil"}


def _proc(backend, maxp=CKPT_MAX):
    from karanta_ocr_b200 import KarantaImageProcessor
    return KarantaImageProcessor(min_pixels=3136, max_pixels=maxp, resize_backend=backend)


@pytest.mark.parametrize("backend", ["torchvision", "pil"])
def test_small_images_equal_golden(backend):
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    p = _proc(backend)
    for name in ("coord_56x84", "noise_100x37", "noise_61x230", "noise_300x200", "page_256x256"):
        out = p(images=[z[f"{name}.image"]], return_tensors="pt")
        assert out["pixel_values"].dtype == torch.float32 and out["pixel_values"].device.type == "cpu"
        assert out["image_grid_thw"].dtype == torch.int64
        assert np.array_equal(out["image_grid_thw"].numpy(), z[f"{name}.{GOLD[backend]}.grid"])
        assert np.array_equal(out["pixel_values"].numpy(), z[f"{name}.{GOLD[backend]}.pixel_values"]), name


@pytest.mark.parametrize("backend", ["torchvision", "pil"])
@pytest.mark.parametrize("name,shape,seed,maxp", [
    ("letter_1288x995", (1288, 995), 1234, CKPT_MAX),
    ("letter_1288x995_classmax", (1288, 995), 1234, 1003520),
    ("a4_1288x910", (1288, 910), 1235, CKPT_MAX),
    ("landscape_995x1288", (995, 1288), 1236, CKPT_MAX),
    ("column_1288x420", (1288, 420), 1237, CKPT_MAX),
    ("datagen_2048x1583", (2048, 1583), 1238, CKPT_MAX),
])
def test_pages_equal_golden_crc(backend, name, shape, seed, maxp):
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    out = _proc(backend, maxp)(images=[synth_page(*shape, seed)], return_tensors="pt")
    pv = out["pixel_values"].numpy()
    assert np.array_equal(out["image_grid_thw"].numpy(), z[f"{name}.{GOLD[backend]}.grid"])
    assert np.array_equal(pv.reshape(-1)[::1009], z[f"{name}.{GOLD[backend]}.sub"])
    assert zlib.crc32(pv.tobytes()) == int(z[f"{name}.{GOLD[backend]}.crc"])


def test_mixed_batch_order_and_layouts():
    """One call, mixed shapes (grouping / reorder in HF): input order is preserved; HWC, CHW, PIL and gray inputs agree."""
    from PIL import Image
    z = np.load(os.path.join(G, "g2_g4_pixel_values.npz"))
    pages = [synth_page(256, 256, 21), synth_page(640, 880, 22), synth_page(256, 256, 23), synth_page(308, 196, 24)]
    p = _proc("torchvision")
    out = p(images=[torch.from_numpy(x) for x in pages], return_tensors="pt")
    assert np.array_equal(out["image_grid_thw"].numpy(), z["mixed.grid"])
    assert zlib.crc32(out["pixel_values"].numpy().tobytes()) == int(z["mixed.crc"])
    hwc = p(images=[x.transpose(1, 2, 0).copy() for x in pages])
    pil = p(images=[Image.fromarray(x.transpose(1, 2, 0)) for x in pages])
    assert torch.equal(hwc["pixel_values"], out["pixel_values"]) and torch.equal(pil["pixel_values"], out["pixel_values"])
    gray = synth_page(300, 260, 5, gray=True)
    a = p(images=[gray])["pixel_values"]
    b = p(images=[Image.fromarray(gray[0], mode="L")])["pixel_values"]
    c = p(images=[gray[0]])["pixel_values"]
    assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("backend", ["torchvision", "pil"])
def test_random_sizes_vs_oracle(backend):
    rng = np.random.default_rng(11)
    p = _proc(backend)
    shapes = [(28, 28), (29, 5600), (57, 31), (700, 1300), (1500, 333), (3100, 2300), (420, 1288), (1288, 1288)]
    pages = [rng.integers(0, 256, (3, h, w), dtype=np.uint8) for h, w in shapes]
    out = p(images=pages)
    ref_pv, ref_grid = po.preprocess(pages, 3136, CKPT_MAX, MODES[backend])
    assert np.array_equal(out["image_grid_thw"].numpy(), ref_grid)
    assert np.array_equal(out["pixel_values"].numpy(), ref_pv)


def test_downscale_class_default_max_pixels_vs_oracle():
    page = synth_page(2048, 1583, 77)
    out = _proc("torchvision", 1003520)(images=[page])
    ref_pv, ref_grid = po.preprocess([page], 3136, 1003520, po.RESIZE_ATEN)
    assert np.array_equal(out["image_grid_thw"].numpy(), ref_grid) and np.array_equal(out["pixel_values"].numpy(), ref_pv)


def test_bf16_output_is_rounded_f32():
    page = synth_page(1288, 995, 1234)
    p = _proc("torchvision")
    f32, g1 = p.preprocess_device([page], torch.float32)
    b16, g2 = p.preprocess_device([page], torch.bfloat16)
    assert torch.equal(g1, g2) and torch.equal(f32.to(torch.bfloat16), b16)


def test_errors_match_transformers():
    p = _proc("torchvision")
    with pytest.raises(ValueError, match="absolute aspect ratio must be smaller than 200"):
        p(images=[np.zeros((3, 10, 2001), dtype=np.uint8)])
    with pytest.raises(ValueError):
        p(images=[np.zeros((3, 56, 56), dtype=np.float32)])
    with pytest.raises(ValueError):
        p(images=[])


def test_full_batch_idempotent_and_page_independent():
    """C2 shape: 64 letter pages in one call; each page's rows equal the single-page result (size-independent property)."""
    p = _proc("torchvision")
    pages = [synth_page(1288, 995, 1234 + i) for i in range(8)]
    batch, grid = p.preprocess_device(pages * 8, torch.float32)
    assert batch.shape == (64 * 6624, 1176) and (grid.numpy() == [1, 92, 72]).all()
    again, _ = p.preprocess_device(pages * 8, torch.float32)
    assert torch.equal(batch, again)
    for i in (0, 5, 63):
        one, _ = p.preprocess_device([pages[i % 8]], torch.float32)
        assert torch.equal(batch[i * 6624:(i + 1) * 6624], one)


@pytest.mark.parametrize("backend,shape", [("torchvision", (4200, 4100)), ("pil", (3700, 3650))])
def test_largest_page_vs_oracle(backend, shape):
    """A page above max_pixels: resized down to the largest grid the path can produce (~65 536 patches, the longest
    filter supports and the widest row tables), still bit-exact."""
    rng = np.random.default_rng(5)
    page = rng.integers(0, 256, (3,) + shape, dtype=np.uint8)
    out = _proc(backend)(images=[page])
    ref_pv, ref_grid = po.preprocess([page], 3136, CKPT_MAX, MODES[backend])
    g = out["image_grid_thw"].numpy()
    assert np.array_equal(g, ref_grid) and 60000 < int(g[0, 1] * g[0, 2]) <= 65536
    assert np.array_equal(out["pixel_values"].numpy(), ref_pv)


@pytest.mark.parametrize("backend", ["torchvision", "pil"])
@pytest.mark.parametrize("layout", ["chw", "hwc", "gray"])
@pytest.mark.parametrize("h,w,maxp,what", [
    (308, 560, CKPT_MAX, "no resize at all: copy path, no vertical pass"),
    (300, 560, CKPT_MAX, "vertical taps only (copy path + pass 2)"),
    (308, 555, CKPT_MAX, "horizontal up by 1 %: four columns per thread"),
    (308, 570, CKPT_MAX, "horizontal down by 2 %: four columns per thread, window starts up to 4 apart"),
    (600, 800, 213000, "both axes down by ~1.5: two columns per thread"),
    (600, 800, 100000, "both axes down by ~2.2: generic tap loops"),
    (90, 3000, CKPT_MAX, "wide strip: many column tiles, last tile narrower"),
])
def test_every_horizontal_pass_variant_vs_oracle(backend, layout, h, w, maxp, what):
    """The horizontal pass has four shapes (copy, 4 / 2 adjacent columns per thread, generic loops) chosen per page from its
    tap table; each is held bit-exact to the oracle for both resize arithmetics and the three input layouts."""
    rng = np.random.default_rng(h * 7919 + w)
    chw = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    if layout == "gray":
        chw = np.repeat(chw[:1], 3, axis=0)
        img = chw[0]
    elif layout == "hwc":
        img = np.ascontiguousarray(chw.transpose(1, 2, 0))
    else:
        img = chw
    out = _proc(backend, maxp)(images=[img])
    ref_pv, ref_grid = po.preprocess([chw], 3136, maxp, MODES[backend])
    assert np.array_equal(out["image_grid_thw"].numpy(), ref_grid), what
    assert np.array_equal(out["pixel_values"].numpy(), ref_pv), what
