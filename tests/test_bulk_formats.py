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


def test_one_bad_page_in_a_batch_fails_alone(tmp_path):
    """a corrupt PNG, an unencodable page and a hostile custom_id each fail their own task; the other
    pages of the same batch are encoded, nothing is double counted and nothing is written outside out_dir."""
    pages = [synth_page(56, 56, 1), synth_page(84, 56, 2), synth_page(56, 84, 3), synth_page(112, 56, 4)]
    path = make_requests(tmp_path, pages)
    lines = open(path).read().splitlines()
    rec = json.loads(lines[1])                                   # page 1: not an image at all
    body = rec.get("body", rec)
    body["messages"][0]["content"][1]["image_url"]["url"] = "data:image/png;base64," + base64.b64encode(b"not a png").decode()
    lines[1] = json.dumps(rec)
    rec = json.loads(lines[3])                                   # page 3: a path as custom_id
    rec["custom_id"] = "../../escape"
    lines[3] = json.dumps(rec)
    open(path, "w").write("\n".join(lines) + "\n")

    class Picky(FakeEncoder):
        def encode_to_host(self, pages):
            if any(p.size == (84, 56) for p in pages):       # page 2 (56 rows x 84 columns) cannot be encoded
                raise ValueError("absolute aspect ratio must be smaller than 200")
            return super().encode_to_host(pages)
    out = tmp_path / "job"
    s = bulk.run_encode_job(path, str(out), Picky(), batch_pages=4)
    assert (s["completed"], s["failed"], s["skipped"]) == (1, 3, 0)
    assert "result" in json.load(open(out / "results" / "doc0.pdf-1.json"))
    assert "error" in json.load(open(out / "results" / "doc1.pdf-2.json"))
    assert "aspect ratio" in json.load(open(out / "results" / "doc2.pdf-3.json"))["error"]
    assert not (tmp_path / "escape.json").exists() and not (tmp_path.parent / "escape.json").exists()
    bad = [f for f in (out / "results").iterdir() if f.name.startswith("invalid_id_")]
    assert len(bad) == 1 and json.load(open(bad[0]))["task_id"] == "../../escape"
    with pytest.raises(ValueError):
        bulk.safe_task_id("a/b")
    with pytest.raises(ValueError):
        bulk.safe_task_id(".hidden")


def test_resume_skips_finished_tasks_and_keeps_state_in_sqlite(tmp_path):
    """bulk_processing/workers/inference_worker.py:315-321 (skip when the result file exists) and
    bulk_processing/utils/database.py:16-49,201-222 (job / task tables, pending = pending or retryable failed)."""
    import sqlite3
    pages = [synth_page(56, 56, 1), synth_page(84, 56, 2), synth_page(56, 84, 3)]
    path = make_requests(tmp_path, pages)
    out, db = tmp_path / "job", str(tmp_path / "state.db")

    class FailsOnce(FakeEncoder):
        armed = True

        def encode_to_host(self, pages):
            if FailsOnce.armed and any(p.size == (84, 56) for p in pages):
                raise RuntimeError("transient")
            return super().encode_to_host(pages)
    s1 = bulk.run_encode_job(path, str(out), FailsOnce(), batch_pages=2, state_db=db, job_id="job1")
    assert (s1["completed"], s1["failed"]) == (2, 1)
    conn = sqlite3.connect(db)
    assert dict(conn.execute("SELECT status, COUNT(*) FROM tasks GROUP BY status").fetchall()) == {"completed": 2, "failed": 1}
    assert conn.execute("SELECT status, total_tasks, completed_tasks, failed_tasks FROM jobs").fetchone() == ("completed_with_errors", 3, 2, 1)
    FailsOnce.armed = False
    enc = FailsOnce()
    s2 = bulk.run_encode_job(path, str(out), enc, batch_pages=2, state_db=db, job_id="job1")
    assert (s2["completed"], s2["failed"], s2["skipped"]) == (1, 0, 2) and len(enc.seen) == 1   # only the failed page ran again
    assert conn.execute("SELECT status, completed_tasks, failed_tasks FROM jobs").fetchone() == ("completed", 3, 0)
    assert conn.execute("SELECT attempts FROM tasks WHERE task_id = 'doc2.pdf-3'").fetchone() == (2,)
    # without a database the result files alone make the job resumable
    s3 = bulk.run_encode_job(path, str(out), FakeEncoder(), batch_pages=2)
    assert (s3["completed"], s3["skipped"]) == (0, 3)
    s4 = bulk.run_encode_job(path, str(out), FakeEncoder(), batch_pages=2, resume=False)
    assert (s4["completed"], s4["skipped"]) == (3, 0)
