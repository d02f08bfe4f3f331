"""Encode-only bulk job over karanta-ocr's own wire formats.

Input: the JSONL that `karanta/data/create_batch_data_prompts.py:84-120` writes for the vLLM workers - one request per
line, `{"custom_id": "<pdf>-<page>", "messages": [{"role": "user", "content": [{"type": "text", ...}, {"type":
"image_url", "image_url": {"url": "data:image/png;base64,..."}}]}], ...}` (`karanta/data/utils.py:269-297`; the OpenAI
batch variant nests the same thing under "body"). Pages are base64 PNG (grayscale 'L' when `convert_to_grayscale`,
`karanta/data/utils.py:186-251`) or JPEG.

Output: what `bulk_processing/workers/inference_worker.py:205-228` leaves behind - `results/<task_id>.json` with
`{"task_id", "result", "timestamp"}` - where `result` points at the page's embedding file instead of generated text.

Host side only: decoding is PIL on a thread pool (the reference decodes on the CPU too), the pages then go through
PageEncoder (preprocess + tower on the GPU) in batches, sharded over ranks with shard_pages when world_size > 1.
"""
from __future__ import annotations

import base64
import io
import json
import os
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


def run_encode_job(requests_jsonl: str, out_dir: str, encoder, batch_pages: int = 64, rank: int = 0, world_size: int = 1,
                   decode_threads: int | None = None) -> dict:
    """Encode every page of `requests_jsonl` that falls in this rank's shard; write `<out_dir>/results/<task_id>.json` and
    `<out_dir>/embeddings/<task_id>.npy` (bf16 bit patterns as uint16, shape [tokens, out_hidden]). Returns a summary."""
    reqs = read_requests(requests_jsonl)
    res_dir, emb_dir = os.path.join(out_dir, "results"), os.path.join(out_dir, "embeddings")
    os.makedirs(res_dir, exist_ok=True)
    os.makedirs(emb_dir, exist_ok=True)
    minp, maxp = encoder.processor.min_pixels, encoder.processor.max_pixels
    pool = ThreadPoolExecutor(decode_threads or min(32, os.cpu_count() or 4))
    sizes = list(pool.map(lambda r: _peek_size(r[1]), reqs))
    mine = shard_pages([page_cost(h, w, minp, maxp) for h, w in sizes], world_size)[rank]
    t0 = time.time()
    done = failed = 0
    batches = [mine[b:b + batch_pages] for b in range(0, len(mine), batch_pages)]
    nxt = pool.map(lambda i: decode_data_uri(reqs[i][1]), batches[0]) if batches else None
    for bi, idx in enumerate(batches):
        pages = list(nxt)
        if bi + 1 < len(batches):  # decode the next batch while this one is on the GPU
            nxt = pool.map(lambda i: decode_data_uri(reqs[i][1]), batches[bi + 1])
        try:
            emb_host, grid = encoder.encode_to_host(pages)
            rows = emb_host.view(dtype=__import__("torch").uint16).numpy()
            off = 0
            for i, g in zip(idx, grid.tolist()):
                n = g[0] * g[1] * g[2] // 4
                tid = reqs[i][0]
                np.save(os.path.join(emb_dir, f"{tid}.npy"), rows[off:off + n])
                off += n
                with open(os.path.join(res_dir, f"{tid}.json"), "w") as f:
                    json.dump({"task_id": tid,
                               "result": {"embedding_file": f"embeddings/{tid}.npy", "dtype": "bfloat16", "shape": [n, int(rows.shape[1])],
                                          "image_grid_thw": g, "num_image_tokens": n},
                               "timestamp": time.time()}, f, indent=2)
                done += 1
        except Exception as e:  # a bad page fails its batch, like a failed task in the reference's queue; the job goes on
            failed += len(idx)
            for i in idx:
                with open(os.path.join(res_dir, f"{reqs[i][0]}.json"), "w") as f:
                    json.dump({"task_id": reqs[i][0], "error": f"{type(e).__name__}: {e}", "timestamp": time.time()}, f, indent=2)
    pool.shutdown()
    dt = time.time() - t0
    return {"rank": rank, "world_size": world_size, "total_requests": len(reqs), "completed": done, "failed": failed,
            "seconds": dt, "pages_per_s": done / dt if dt > 0 else None}


def load_embedding(out_dir: str, task_id: str):
    """Read one page's embeddings back as a torch bf16 tensor."""
    import torch
    a = np.load(os.path.join(out_dir, "embeddings", f"{task_id}.npy"))
    return torch.from_numpy(a).view(torch.bfloat16)
