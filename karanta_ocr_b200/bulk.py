"""Encode-only bulk job over karanta-ocr's own wire formats.

Input: the JSONL that `karanta/data/create_batch_data_prompts.py:84-120` writes for the vLLM workers - one request per
line, `{"custom_id": "<pdf>-<page>", "messages": [{"role": "user", "content": [{"type": "text", ...}, {"type":
"image_url", "image_url": {"url": "data:image/png;base64,..."}}]}], ...}` (`karanta/data/utils.py:269-297`; the OpenAI
batch variant nests the same thing under "body"). Pages are base64 PNG (grayscale 'L' when `convert_to_grayscale`,
`karanta/data/utils.py:186-251`) or JPEG.

Output: what `bulk_processing/workers/inference_worker.py:205-228` leaves behind - `results/<task_id>.json` with
`{"task_id", "result", "timestamp"}` - where `result` points at the page's embedding file instead of generated text.

PNG pages are decoded on the GPU (inflate + scan-line filters in libkocr.so, row f2): the host only undoes the base64. JPEG and
the PNG flavours the kernels do not take fall to Pillow on a thread pool, the reference's own decode. The pages then go
through PageEncoder (preprocess + tower) in batches, sharded over ranks with shard_pages when world_size > 1.
"""
from __future__ import annotations

import base64
import hashlib
import io
import json
import os
import re
import sqlite3
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .pipeline import page_cost, shard_pages


def _image_url(record: dict) -> str:
    body = record.get("body", record)
    for msg in body.get("messages", []):
        content = msg.get("content")
        if isinstance(content, list):
            for part in content:
                if isinstance(part, dict) and part.get("type") == "image_url":
                    url = part["image_url"]
                    return url["url"] if isinstance(url, dict) else url
    raise ValueError(f"request {record.get('custom_id')!r} carries no image_url part")


def decode_data_uri(url: str):
    """`data:image/<fmt>;base64,<payload>` (or a bare base64 string, as base64_to_grayscale accepts) -> PIL image, 'L' or 'RGB'."""
    from PIL import Image
    payload = url.split(",", 1)[1] if url.startswith("data:") else url
    img = Image.open(io.BytesIO(base64.b64decode(payload)))
    img.load()
    return img if img.mode in ("L", "RGB") else img.convert("RGB")


def read_requests(path: str):
    """[(custom_id, data-URI string)] in file order."""
    out = []
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            line = line.strip()
            if not line:
                continue
            rec = json.loads(line)
            cid = rec.get("custom_id")
            if cid is None:
                raise ValueError(f"{path}:{ln}: request without custom_id")
            out.append((str(cid), _image_url(rec)))
    return out


def _peek_size(url: str):
    from PIL import Image
    payload = url.split(",", 1)[1] if url.startswith("data:") else url
    with Image.open(io.BytesIO(base64.b64decode(payload))) as im:  # header only
        return im.height, im.width


_SAFE_ID = re.compile(r"^[A-Za-z0-9][A-Za-z0-9._@+=,-]*$")


def safe_task_id(custom_id: str) -> str:
    """`custom_id` names files under out_dir: refuse anything that is not a plain file name (no separators, no leading
    dot, no '..'), as request files are untrusted input."""
    if not _SAFE_ID.match(custom_id) or ".." in custom_id or len(custom_id) > 200:
        raise ValueError(f"custom_id {custom_id!r} is not usable as a file name")
    return custom_id


class JobState:
    """Job / task bookkeeping in SQLite with the reference's schema (bulk_processing/utils/database.py:16-49: tables `jobs`
    and `tasks`, statuses pending -> processing -> completed | failed, attempts, error_message, processing_time_ms), so a
    job can be inspected with the reference's own queries and resumed: pending = pending or (failed and attempts < max),
    bulk_processing/utils/database.py:201-222."""

    def __init__(self, db_path: str, job_id: str, config: dict | None = None):
        self.job_id = job_id
        self.conn = sqlite3.connect(db_path)
        self.conn.executescript("""
            CREATE TABLE IF NOT EXISTS jobs (job_id TEXT PRIMARY KEY, status TEXT NOT NULL DEFAULT 'created', config TEXT,
                created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP, updated_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP,
                total_tasks INTEGER DEFAULT 0, completed_tasks INTEGER DEFAULT 0, failed_tasks INTEGER DEFAULT 0,
                processing_tasks INTEGER DEFAULT 0);
            CREATE TABLE IF NOT EXISTS tasks (task_id TEXT PRIMARY KEY, job_id TEXT NOT NULL, status TEXT NOT NULL DEFAULT 'pending',
                request_data TEXT NOT NULL, result_data TEXT, error_message TEXT, attempts INTEGER DEFAULT 0,
                created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP, updated_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP,
                started_at TIMESTAMP, completed_at TIMESTAMP, processing_time_ms INTEGER,
                FOREIGN KEY (job_id) REFERENCES jobs (job_id));
            CREATE INDEX IF NOT EXISTS idx_tasks_job_id ON tasks(job_id);
            CREATE INDEX IF NOT EXISTS idx_tasks_status ON tasks(status);
        """)
        self.conn.execute("INSERT OR IGNORE INTO jobs (job_id, status, config) VALUES (?, 'running', ?)", (job_id, json.dumps(config or {})))
        self.conn.commit()

    def add_tasks(self, task_ids):
        self.conn.executemany("INSERT OR IGNORE INTO tasks (task_id, job_id, request_data) VALUES (?, ?, ?)",
                              [(t, self.job_id, json.dumps({"custom_id": t})) for t in task_ids])
        self.conn.execute("UPDATE jobs SET total_tasks = (SELECT COUNT(*) FROM tasks WHERE job_id = ?) WHERE job_id = ?", (self.job_id, self.job_id))
        self.conn.commit()

    def pending(self, max_attempts: int = 3):
        q = "SELECT task_id FROM tasks WHERE job_id = ? AND (status = 'pending' OR status = 'processing' OR (status = 'failed' AND attempts < ?))"
        return {r[0] for r in self.conn.execute(q, (self.job_id, max_attempts))}

    def update(self, rows):
        """rows: [(task_id, status, error_message or None, processing_time_ms or None)], one transaction (the reference
        batches its DB writes too, inference_worker.py:100-119)."""
        for tid, status, err, ms in rows:
            if status == "processing":
                self.conn.execute("UPDATE tasks SET status = 'processing', attempts = attempts + 1, started_at = CURRENT_TIMESTAMP, "
                                  "updated_at = CURRENT_TIMESTAMP WHERE task_id = ?", (tid,))
            else:
                self.conn.execute("UPDATE tasks SET status = ?, error_message = ?, processing_time_ms = ?, completed_at = CURRENT_TIMESTAMP, "
                                  "updated_at = CURRENT_TIMESTAMP WHERE task_id = ?", (status, err, ms, tid))
        self.conn.execute("""UPDATE jobs SET updated_at = CURRENT_TIMESTAMP,
            completed_tasks = (SELECT COUNT(*) FROM tasks WHERE job_id = :j AND status = 'completed'),
            failed_tasks = (SELECT COUNT(*) FROM tasks WHERE job_id = :j AND status = 'failed'),
            processing_tasks = (SELECT COUNT(*) FROM tasks WHERE job_id = :j AND status = 'processing') WHERE job_id = :j""", {"j": self.job_id})
        self.conn.commit()

    def finish(self):
        self.conn.execute("UPDATE jobs SET status = CASE WHEN failed_tasks > 0 THEN 'completed_with_errors' ELSE 'completed' END, "
                          "updated_at = CURRENT_TIMESTAMP WHERE job_id = ?", (self.job_id,))
        self.conn.commit()

    def counts(self):
        return dict(self.conn.execute("SELECT status, COUNT(*) FROM tasks WHERE job_id = ? GROUP BY status", (self.job_id,)).fetchall())


def _result_done(path: str) -> bool:
    """A task whose result file exists is not processed again (bulk_processing/workers/inference_worker.py:315-321); a file
    that only records an error is."""
    try:
        with open(path) as f:
            return "result" in json.load(f)
    except Exception:
        return False


def run_encode_job(requests_jsonl: str, out_dir: str, encoder, batch_pages: int = 64, rank: int = 0, world_size: int = 1,
                   decode_threads: int | None = None, decode: str = "auto", resume: bool = True, state_db: str | None = None, job_id: str | None = None,
                   max_attempts: int = 3) -> dict:
    """Encode every page of `requests_jsonl` that falls in this rank's shard; write `<out_dir>/results/<task_id>.json` and
    `<out_dir>/embeddings/<task_id>.npy` (bf16 bit patterns as uint16, shape [tokens, out_hidden]). Returns a summary.

    Failure is per page, as in the reference's queue where a task fails alone: a request that cannot be parsed, decoded or
    encoded (corrupt image, aspect ratio > 200, unusable custom_id) gets an `error` result file and the job goes on; when a
    whole batch fails on the GPU its pages are retried one by one to find the bad one. With `resume`, tasks whose result
    file already exists are skipped (inference_worker.py:315-321). With `state_db`, job / task state is kept in SQLite under
    the reference's schema (bulk_processing/utils/database.py:16-49) and only pending / retryable tasks are run.
    `decode`: "gpu" = PNG pages go to the device undecoded, "host" = Pillow on the thread pool as the
    reference does, "auto" = gpu when the encoder takes PNG bytes (PageEncoder does)."""
    import torch
    reqs = read_requests(requests_jsonl)
    res_dir, emb_dir = os.path.join(out_dir, "results"), os.path.join(out_dir, "embeddings")
    os.makedirs(res_dir, exist_ok=True)
    os.makedirs(emb_dir, exist_ok=True)
    minp, maxp = encoder.processor.min_pixels, encoder.processor.max_pixels
    pool = ThreadPoolExecutor(decode_threads or min(32, os.cpu_count() or 4))
    t0 = time.time()
    counts = {"completed": 0, "failed": 0, "skipped": 0, "gpu_decoded": 0}
    state = JobState(state_db, job_id or os.path.basename(requests_jsonl), {"requests": requests_jsonl, "batch_pages": batch_pages}) if state_db else None
    db_rows = []

    def fail_task(i, err, name=None):
        tid = name or reqs[i][0]
        with open(os.path.join(res_dir, f"{tid}.json"), "w") as f:
            json.dump({"task_id": reqs[i][0], "error": err, "timestamp": time.time()}, f, indent=2)
        counts["failed"] += 1
        db_rows.append((tid, "failed", err, None))

    # ---- admission: usable ids, not already done, header readable (one bad request never stops the others)
    def peek(i):
        try:
            return _peek_size(reqs[i][1])
        except Exception as e:
            return e
    names = {}
    for i, (cid, _) in enumerate(reqs):
        try:
            names[i] = safe_task_id(cid)
        except ValueError:
            names[i] = None
    if state:
        state.add_tasks([n for n in names.values() if n])
        retryable = state.pending(max_attempts)
    # the shard of a request depends only on the request file, never on progress, so ranks that start at different times (or a
    # resumed job) agree on who owns what; finished tasks are skipped inside the owner's shard
    cand, sizes = [], {}
    peeked = list(pool.map(peek, range(len(reqs))))
    for i, (cid, _) in enumerate(reqs):
        if names[i] is None:
            if rank == 0:
                fail_task(i, f"ValueError: custom_id {cid!r} is not usable as a file name", name="invalid_id_" + hashlib.sha1(cid.encode()).hexdigest()[:16])
                db_rows.pop()  # not a task the database knows
        elif isinstance(peeked[i], Exception):
            if rank == 0:
                fail_task(i, f"{type(peeked[i]).__name__}: {peeked[i]}", names[i])
        else:
            sizes[i] = peeked[i]
            cand.append(i)
    shard = shard_pages([page_cost(*sizes[i], minp, maxp) for i in cand], world_size)[rank]
    mine = []
    for k in shard:
        i = cand[k]
        if resume and _result_done(os.path.join(res_dir, f"{names[i]}.json")):
            counts["skipped"] += 1
            db_rows.append((names[i], "completed", None, None))
        elif state and names[i] not in retryable:
            counts["skipped"] += 1
        else:
            mine.append(i)
    if state and db_rows:
        state.update(db_rows)
    db_rows.clear()

    gpu_decode = decode == "gpu" or (decode == "auto" and getattr(encoder, "accepts_png_bytes", False))

    def decode_page(i):
        """GPU decode: hand the PNG file bytes through (base64 is undone here, inflate + unfilter run on the device). Anything
        the device kernels do not take - JPEG, palette / 16-bit / interlaced PNG - is decoded by Pillow, as the reference does."""
        try:
            if gpu_decode:
                from .png_decode import is_gpu_decodable, payload_bytes
                data = payload_bytes(reqs[i][1])
                if is_gpu_decodable(data):
                    counts["gpu_decoded"] += 1
                    return data
            return decode_data_uri(reqs[i][1])
        except Exception as e:
            return e

    def encode(pages):
        t1 = time.time()
        emb_host, grid = encoder.encode_to_host(pages)   # raises before anything is written
        return emb_host.view(dtype=torch.uint16).numpy(), grid.tolist(), int((time.time() - t1) * 1e3 / max(len(pages), 1))

    def write(idx, rows, grid, ms):
        off = 0
        for i, g in zip(idx, grid):
            n = g[0] * g[1] * g[2] // 4
            tid = names[i]
            np.save(os.path.join(emb_dir, f"{tid}.npy"), rows[off:off + n])
            off += n
            with open(os.path.join(res_dir, f"{tid}.json"), "w") as f:
                json.dump({"task_id": reqs[i][0],
                           "result": {"embedding_file": f"embeddings/{tid}.npy", "dtype": "bfloat16", "shape": [n, int(rows.shape[1])],
                                      "image_grid_thw": g, "num_image_tokens": n},
                           "timestamp": time.time()}, f, indent=2)
            counts["completed"] += 1
            db_rows.append((tid, "completed", None, ms))

    batches = [mine[b:b + batch_pages] for b in range(0, len(mine), batch_pages)]
    nxt = pool.map(decode_page, batches[0]) if batches else None
    for bi, idx in enumerate(batches):
        decoded = list(nxt)
        if bi + 1 < len(batches):  # decode the next batch while this one is on the GPU
            nxt = pool.map(decode_page, batches[bi + 1])
        if state:
            state.update([(names[i], "processing", None, None) for i in idx])
        good = [(i, p) for i, p in zip(idx, decoded) if not isinstance(p, Exception)]
        for i, p in zip(idx, decoded):
            if isinstance(p, Exception):
                fail_task(i, f"{type(p).__name__}: {p}", names[i])
        if good:
            try:
                out = encode([p for _, p in good])
            except Exception:
                # something in the batch is not encodable (e.g. aspect ratio > 200): one page at a time, only the bad ones fail
                for i, p in good:
                    try:
                        write([i], *encode([p]))
                    except Exception as e:
                        fail_task(i, f"{type(e).__name__}: {e}", names[i])
            else:
                write([i for i, _ in good], *out)
        if state and db_rows:
            state.update(db_rows)
        db_rows.clear()
    pool.shutdown()
    if state:
        state.finish()
    dt = time.time() - t0
    return {"rank": rank, "world_size": world_size, "total_requests": len(reqs), "completed": counts["completed"], "failed": counts["failed"],
            "skipped": counts["skipped"], "gpu_decoded_pages": counts["gpu_decoded"], "seconds": dt, "pages_per_s": counts["completed"] / dt if dt > 0 else None}


def load_embedding(out_dir: str, task_id: str):
    """Read one page's embeddings back as a torch bf16 tensor."""
    import torch
    a = np.load(os.path.join(out_dir, "embeddings", f"{task_id}.npy"))
    return torch.from_numpy(a).view(torch.bfloat16)
