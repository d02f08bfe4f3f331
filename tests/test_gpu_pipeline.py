"""GPU: the public end-to-end calls (host pages in, host embeddings out), synchronous and pipelined, agree bit for bit."""
import pytest
import torch

from oracle import vision_oracle as vo
from tests.synth import synth_page

pytestmark = pytest.mark.gpu


def test_async_pipeline_equals_sync():
    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder
    cfg = vo.TowerConfig("qwen2_vl", 2, 1280, 16, 5120, 1536)
    tower = KarantaVisionTower(dict(arch="qwen2_vl", depth=2, embed_dim=1280, num_heads=16, mlp_hidden=5120, out_hidden=1536))
    tower.load_state_dict(vo.init_weights(cfg, seed=100))
    enc = PageEncoder(tower)
    batches = [[torch.from_numpy(synth_page(420, 322, 70 + 3 * b + i)).pin_memory() for i in range(3)] for b in range(4)]
    ref = [enc.encode_to_host(b)[0].clone() for b in batches]
    outs = [torch.empty((ref[0].shape[0] + 64, 1536), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    got, pending = [], []
    for k, b in enumerate(batches):
        if len(pending) >= 2:
            ev, buf, n = pending.pop(0)
            ev.synchronize()
            got.append(buf[:n].clone())
        ev, grid, n = enc.encode_to_host_async(b, outs[k % 2])
        pending.append((ev, outs[k % 2], n))
    for ev, buf, n in pending:
        ev.synchronize()
        got.append(buf[:n].clone())
    assert len(got) == len(ref) and all(torch.equal(a, b) for a, b in zip(got, ref))


def test_sharded_encode_single_rank_order():
    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, gather_pages
    cfg = vo.TowerConfig("qwen2_vl", 1, 160, 2, 640, 256)
    tower = KarantaVisionTower(dict(arch="qwen2_vl", depth=1, embed_dim=160, num_heads=2, mlp_hidden=640, out_hidden=256))
    tower.load_state_dict(vo.init_weights(cfg, seed=3))
    enc = PageEncoder(tower)
    pages = [synth_page(56 * (1 + i % 3), 84, 90 + i) for i in range(5)]
    outs, idx = enc.encode_sharded(pages, rank=0, world_size=1, batch_pages=2)
    ordered = gather_pages(outs, idx, len(pages))
    for i, p in enumerate(pages):
        one, _ = enc.encode([p])
        assert torch.equal(ordered[i], one.cpu())


def test_bulk_encode_job_over_reference_jsonl(tmp_path):
    """SURVEY.md section 8 row f4 on the GPU: request JSONL in the reference's schema (RGB and grayscale PNG data-URIs) ->
    result files; each page's stored embedding equals encoding that page on its own."""
    import json

    import numpy as np
    from PIL import Image

    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, bulk
    from tests.test_bulk_formats import make_requests
    cfg = vo.TowerConfig("qwen2_vl", 1, 160, 2, 640, 256)
    tower = KarantaVisionTower(dict(arch="qwen2_vl", depth=1, embed_dim=160, num_heads=2, mlp_hidden=640, out_hidden=256))
    tower.load_state_dict(vo.init_weights(cfg, seed=4))
    enc = PageEncoder(tower)
    pages = [synth_page(140, 112, 1), synth_page(84, 196, 2), synth_page(56, 56, 3), synth_page(280, 280, 4), synth_page(112, 140, 5)]
    path = make_requests(tmp_path, pages, gray={1, 4})
    s = bulk.run_encode_job(path, str(tmp_path / "job"), enc, batch_pages=2)
    assert s["completed"] == 5 and s["failed"] == 0
    for i, p in enumerate(pages):
        tid = f"doc{i}.pdf-{i + 1}"
        img = Image.fromarray(np.transpose(p, (1, 2, 0)))
        if i in (1, 4):
            img = img.convert("L")
        one, grid = enc.encode([img])
        rec = json.load(open(tmp_path / "job" / "results" / f"{tid}.json"))
        assert rec["result"]["image_grid_thw"] == grid.tolist()[0]
        assert torch.equal(bulk.load_embedding(str(tmp_path / "job"), tid), one.cpu())
