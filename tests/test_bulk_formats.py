"""the encode-only bulk job reads the reference's request JSONL
(karanta/data/create_batch_data_prompts.py:84-120, karanta/data/utils.py:269-297) and leaves result files laid out like
bulk_processing/workers/inference_worker.py:205-228. Host-side logic here; the GPU run is in test_gpu_pipeline.py."""
import base64
import io
import json

import numpy as np
import pytest
import torch
from PIL import Image

from karanta_ocr_b200 import bulk
from tests.synth import synth_page


def _png_b64(arr_chw, gray=False):
    img = Image.fromarray(np.transpose(arr_chw, (1, 2, 0)))
    if gray:
        img = img.convert("L")
    buf = io.BytesIO()
    img.save(buf, format="PNG")
    return base64.b64encode(buf.getvalue()).decode()


def make_requests(tmp_path, pages, gray=()):
    """Writes the JSONL exactly as build_page_query_vllm_olmoocr + create_vision_message produce it."""
    path = tmp_path / "requests.jsonl"
    with open(path, "w") as f:
        for i, p in enumerate(pages):
            rec = {"custom_id": f"doc{i}.pdf-{i + 1}", "azure_source_dir": None, "model": "olmocr",
                   "messages": [{"role": "user", "content": [{"type": "text", "text": "prompt"},
                                                             {"type": "image_url", "image_url": {"url": f"data:image/png;base64,{_png_b64(p, i in gray)}"}}]}],
                   "temperature": 0.1, "max_tokens": 6000}
            if i % 2:
                rec = {"custom_id": rec["custom_id"], "method": "POST", "url": "/v1/chat/completions", "body": rec}  # OpenAI batch nesting
            f.write(json.dumps(rec) + "\n")
    return str(path)


class FakeEncoder:
    """Stands in for PageEncoder on a box without a GPU: 'embedding' = per-token mean of the page, to check the plumbing."""
    class P:
        min_pixels, max_pixels = 3136, 12845056
    processor = P()

    def __init__(self):
        self.seen = []

    def encode_to_host(self, pages):
        from karanta_ocr_b200 import smart_resize
        rows, grid = [], []
        for p in pages:
            self.seen.append((p.mode, p.size))
            rh, rw = smart_resize(p.height, p.width, 28, 3136, 12845056)
            n = (rh // 14) * (rw // 14) // 4
            grid.append([1, rh // 14, rw // 14])
            rows.append(torch.full((n, 8), float(np.asarray(p).mean()), dtype=torch.bfloat16))
        return torch.cat(rows), torch.tensor(grid)


def test_read_requests_and_decode(tmp_path):
    pages = [synth_page(140, 112, 1), synth_page(84, 196, 2), synth_page(56, 56, 3)]
    path = make_requests(tmp_path, pages, gray={2})
    reqs = bulk.read_requests(path)
    assert [r[0] for r in reqs] == ["doc0.pdf-1", "doc1.pdf-2", "doc2.pdf-3"]
    imgs = [bulk.decode_data_uri(u) for _, u in reqs]
    assert [im.mode for im in imgs] == ["RGB", "RGB", "L"]
    assert np.array_equal(np.asarray(imgs[0]), np.transpose(pages[0], (1, 2, 0)))  # PNG is lossless
    assert bulk.decode_data_uri(reqs[0][1].split(",", 1)[1]).size == imgs[0].size  # bare base64 is accepted too
    with pytest.raises(ValueError):
        bulk._image_url({"custom_id": "x", "messages": [{"role": "user", "content": "text only"}]})


@pytest.mark.parametrize("world", [1, 2])
def test_job_layout_and_sharding(tmp_path, world):
    pages = [synth_page(140, 112, 1), synth_page(84, 196, 2), synth_page(56, 56, 3), synth_page(280, 280, 4), synth_page(112, 140, 5)]
    path = make_requests(tmp_path, pages)
    out = tmp_path / "job"
    summaries = [bulk.run_encode_job(path, str(out), FakeEncoder(), batch_pages=2, rank=r, world_size=world) for r in range(world)]
    assert sum(s["completed"] for s in summaries) == 5 and all(s["failed"] == 0 for s in summaries)
    for i, p in enumerate(pages):
        tid = f"doc{i}.pdf-{i + 1}"
        rec = json.load(open(out / "results" / f"{tid}.json"))
        assert set(rec) == {"task_id", "result", "timestamp"} and rec["task_id"] == tid
        emb = bulk.load_embedding(str(out), tid)
        assert list(emb.shape) == rec["result"]["shape"] and emb.shape[0] == rec["result"]["num_image_tokens"]
        assert rec["result"]["image_grid_thw"][1] * rec["result"]["image_grid_thw"][2] // 4 == emb.shape[0]
        assert abs(float(emb.float().mean()) - p.mean()) < 1.0  # the right page went to the right file


def test_bad_page_fails_its_task_not_the_job(tmp_path):
    path = make_requests(tmp_path, [synth_page(56, 56, 1), synth_page(56, 56, 2)])

    class Failing(FakeEncoder):
        def encode_to_host(self, pages):
            raise ValueError("absolute aspect ratio must be smaller than 200")
    s = bulk.run_encode_job(path, str(tmp_path / "job"), Failing(), batch_pages=1)
    assert s["failed"] == 2 and s["completed"] == 0
    rec = json.load(open(tmp_path / "job" / "results" / "doc0.pdf-1.json"))
    assert "aspect ratio" in rec["error"]
