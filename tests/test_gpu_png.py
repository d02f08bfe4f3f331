"""PNG page decode on the GPU (csrc/kocr_png.cu through the C ABI) is byte-equal to Pillow, the
decoder the reference uses in front of the path (karanta/data/utils.py:186-251, karanta/data/process_pdf_utils.py:50-75),
and the fused PNG -> embeddings call equals decoding on the host first."""
import base64
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import vision_oracle as vo
from tests.synth import synth_page

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _png(img, **kw):
    buf = io.BytesIO()
    img.save(buf, format="PNG", **kw)
    return buf.getvalue()


def _cases():
    rng = np.random.default_rng(3)
    letter = Image.fromarray(synth_page(1288, 995, 1234).transpose(1, 2, 0))
    photo = Image.open(os.path.join(G, "sample_760x1024.png"))
    yield "letter RGB (pdftoppm -png)", letter, {}
    yield "letter gray L (base64_to_grayscale)", letter.convert("L"), {}
    yield "letter RGB optimize", letter, {"optimize": True}
    yield "letter RGB level 1", letter, {"compress_level": 1}
    yield "photo RGB (adaptive filters, Paeth)", photo, {}
    yield "photo gray", photo.convert("L"), {"optimize": True}
    yield "photo RGBA", photo.convert("RGBA"), {}
    yield "photo LA", photo.convert("LA"), {}
    yield "noise (stored blocks)", Image.fromarray(rng.integers(0, 256, (300, 211, 3), dtype=np.uint8)), {}
    yield "noise level 0", Image.fromarray(rng.integers(0, 256, (64, 70), dtype=np.uint8)), {"compress_level": 0}
    yield "1x1", Image.fromarray(np.full((1, 1, 3), 7, dtype=np.uint8)), {}
    yield "1 x 3000", Image.fromarray(rng.integers(0, 3, (1, 3000), dtype=np.uint8) * 100), {}
    yield "3000 x 1", Image.fromarray(rng.integers(0, 3, (3000, 1), dtype=np.uint8) * 100), {}
    yield "white page", Image.fromarray(np.full((900, 700), 255, dtype=np.uint8)), {}
    yield "33 rows (band boundary)", Image.fromarray(synth_page(33, 257, 5).transpose(1, 2, 0)), {}


def test_decode_equals_pillow():
    from karanta_ocr_b200 import decode_png_batch
    names, files, want = [], [], []
    for name, img, kw in _cases():
        data = _png(img, **kw)
        ref = np.asarray(Image.open(io.BytesIO(data)))
        if ref.ndim == 3 and ref.shape[2] in (2, 4):
            ref = ref[:, :, :ref.shape[2] - 1]          # alpha is dropped, like .convert("RGB") / .convert("L")
            if ref.shape[2] == 1:
                ref = ref[:, :, 0]
        names.append(name), files.append(data), want.append(ref)
    got = decode_png_batch(files)                         # one call: many pages in flight
    for name, g, w in zip(names, got, want):
        assert g.dtype == torch.uint8 and g.is_cuda and tuple(g.shape) == w.shape, (name, g.shape, w.shape)
        assert np.array_equal(g.cpu().numpy(), w), name
    one = decode_png_batch(files[:1])[0]
    assert torch.equal(one, got[0])


def test_damaged_pages_are_reported_by_index():
    from karanta_ocr_b200 import PngError, decode_png_batch
    good = _png(Image.fromarray(synth_page(120, 200, 2).transpose(1, 2, 0)))
    # flip a bit inside the compressed data and repair the chunk CRC so that only the stream itself is wrong
    import zlib
    pos = good.index(b"IDAT")
    n = int.from_bytes(good[pos - 4:pos], "big")
    bad = bytearray(good)
    bad[pos + 4 + n // 2] ^= 0x10
    bad[pos + 4 + n:pos + 8 + n] = zlib.crc32(bytes(bad[pos:pos + 4 + n])).to_bytes(4, "big")
    try:
        out = decode_png_batch([good, bytes(bad), good])
    except PngError as e:
        assert e.index == 1
    else:  # a flipped bit may still inflate to the right length: then the pixels differ, which Pillow would show as well
        ref = np.asarray(Image.open(io.BytesIO(good)))
        assert np.array_equal(out[0].cpu().numpy(), ref) and np.array_equal(out[2].cpu().numpy(), ref)
    with pytest.raises(PngError) as ei:
        decode_png_batch([good, good[:len(good) // 2]])
    assert ei.value.index == 1
    with pytest.raises(PngError):
        decode_png_batch([_png(Image.fromarray(synth_page(50, 50, 1)[0]).convert("P"))])   # palette: not a GPU flavour
    pal = _png(Image.fromarray(synth_page(50, 50, 1)[0]).convert("P"))
    from karanta_ocr_b200.png_decode import is_gpu_decodable
    assert not is_gpu_decodable(pal) and is_gpu_decodable(good) and not is_gpu_decodable(b"\xff\xd8\xff\xe0 jpeg")


def test_png_to_embeddings_equals_host_decoded_pages():
    """PageEncoder.encode on PNG files / data URIs = PageEncoder.encode on the Pillow-decoded pages, bit for bit; mixed lists
    (PNG bytes next to already decoded pages) keep their order."""
    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder
    cfg = vo.TowerConfig("qwen2_vl", 2, 1280, 16, 5120, 1536)
    tower = KarantaVisionTower(dict(arch="qwen2_vl", depth=2, embed_dim=1280, num_heads=16, mlp_hidden=5120, out_hidden=1536))
    tower.load_state_dict(vo.init_weights(cfg, seed=100))
    enc = PageEncoder(tower)
    imgs = [Image.fromarray(synth_page(420, 322, 7).transpose(1, 2, 0)), Image.fromarray(synth_page(300, 500, 8).transpose(1, 2, 0)).convert("L"),
            Image.fromarray(synth_page(644, 455, 9).transpose(1, 2, 0))]
    files = [_png(im) for im in imgs]
    ref, grid_ref = enc.encode(imgs)
    got, grid = enc.encode_png(files)
    assert torch.equal(grid, grid_ref) and torch.equal(got, ref)
    uris = ["data:image/png;base64," + base64.b64encode(f).decode() for f in files]
    got2, _ = enc.encode([uris[0], imgs[1], files[2]])
    assert torch.equal(got2, ref)
    out_host = torch.empty(ref.shape, dtype=torch.bfloat16).pin_memory()
    ev, grid3, rows = enc.encode_to_host_async(files, out_host)
    ev.synchronize()
    idx, status = enc.last_png_status
    assert idx == [0, 1, 2] and status.cpu().tolist() == [0, 0, 0]
    assert rows == ref.shape[0] and torch.equal(out_host, ref.cpu())


def test_bulk_job_decodes_png_on_the_gpu(tmp_path):
    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, bulk
    from tests.test_bulk_formats import make_requests
    cfg = vo.TowerConfig("qwen2_vl", 1, 160, 2, 640, 256)
    tower = KarantaVisionTower(dict(arch="qwen2_vl", depth=1, embed_dim=160, num_heads=2, mlp_hidden=640, out_hidden=256))
    tower.load_state_dict(vo.init_weights(cfg, seed=3))
    enc = PageEncoder(tower)
    pages = [synth_page(140, 112, 1), synth_page(84, 196, 2), synth_page(56, 56, 3), synth_page(280, 280, 4)]
    path = make_requests(tmp_path, pages, gray={2})
    s_gpu = bulk.run_encode_job(path, str(tmp_path / "gpu"), enc, batch_pages=3, decode="gpu")
    s_host = bulk.run_encode_job(path, str(tmp_path / "host"), enc, batch_pages=3, decode="host")
    assert s_gpu["completed"] == s_host["completed"] == 4 and s_gpu["gpu_decoded_pages"] == 4 and s_host["gpu_decoded_pages"] == 0
    for i in range(4):
        tid = f"doc{i}.pdf-{i + 1}"
        assert torch.equal(bulk.load_embedding(str(tmp_path / "gpu"), tid), bulk.load_embedding(str(tmp_path / "host"), tid))
