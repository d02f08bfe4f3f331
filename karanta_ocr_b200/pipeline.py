"""PageEncoder: host pages -> embeddings in one call (preprocess kernel writing bf16 straight into the tower), and
the page-sharding helpers for one-process-per-GPU bulk encoding.

karanta-ocr's production topology is one engine per GPU with host-side routing of whole pages
(scripts/start_multiple_vllm_servers.sh:272-317, bulk_processing/utils/gpu_router.py:5-21); pages are independent, so
the batch is partitioned by page with no data-path collective and the results are gathered on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from .image_processor import KarantaImageProcessor, smart_resize
from .vision_tower import KarantaVisionTower


def page_cost(height: int, width: int, min_pixels: int, max_pixels: int, alpha: float = 1.0, beta: float = 1.0 / 1200.0):
    """Relative encode cost a*N + b*N^2 of one page (linear layers + full attention), N = patches after smart_resize."""
    rh, rw = smart_resize(height, width, 28, min_pixels, max_pixels)
    n = (rh // 14) * (rw // 14)
    return alpha * n + beta * n * n


def shard_pages(costs, world_size: int):
    """Greedy longest-processing-time partition of page indices over ranks (vLLM balances its DP-ViT the same way,
    vllm model_executor/models/vision.py:314-359). Returns a list of index lists, one per rank, each in input order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def gather_pages(local_items, local_indices, n_total: int, group=None, dst: int = 0):
    """Host-side gather of per-page results in original page order (the reference routes whole pages to independent
    engines and collects on the host: bulk_processing/workers/inference_worker.py:205-228). `local_items[k]` is the
    result for page `local_indices[k]`. Returns the ordered list on rank `dst`, None elsewhere. Works on any
    torch.distributed backend (objects travel through the CPU); without an initialised group it is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * n_total
        for i, it in zip(local_indices, local_items):
            out[i] = it
        return out
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = [(int(i), it.cpu() if hasattr(it, "cpu") else it) for i, it in zip(local_indices, local_items)]
    buf = [None] * world if rank == dst else None
    dist.gather_object(payload, buf, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n_total
    for part in buf:
        for i, it in part:
            out[i] = it
    if any(o is None for o in out):
        raise RuntimeError("gather_pages: some pages were not produced by any rank")
    return out


class PageEncoder:
    """processor + tower on one GPU. `encode(pages)` returns (embeddings bf16 [sum N / 4, out_hidden] on the GPU,
    image_grid_thw int64 [n, 3])."""

    def __init__(self, tower: KarantaVisionTower, min_pixels: int = 3136, max_pixels: int = 12845056,
                 resize_backend: str = "torchvision"):
        self.tower = tower
        self.processor = KarantaImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels, resize_backend=resize_backend,
                                               device=tower.device)
        self.last_launch_count = 0

    @torch.no_grad()
    def encode_sharded(self, pages, rank: int, world_size: int, batch_pages: int = 64):
        """Encode this rank's LPT shard of `pages` (all ranks pass the same list); returns (per-page embeddings on the
        host, page indices). Pair with gather_pages() for the host-side gather."""
        minp, maxp = self.processor.min_pixels, self.processor.max_pixels
        costs = []
        for p in pages:
            h, w = (p.height, p.width) if hasattr(p, "height") else (p.shape[-2:] if p.shape[0] in (1, 3, 4) and p.ndim == 3 else p.shape[:2])
            costs.append(page_cost(int(h), int(w), minp, maxp))
        mine = shard_pages(costs, world_size)[rank]
        outs = []
        for b in range(0, len(mine), batch_pages):
            idx = mine[b:b + batch_pages]
            emb, grid = self.encode([pages[i] for i in idx])
            outs.extend(t.cpu() for t in self.tower.split_per_image(emb, grid))
        return outs, mine

    accepts_png_bytes = True  # pages may be undecoded PNG files (bytes / data URIs): they are decoded on the GPU
    decode_sms = 8            # SMs given to the decode kernels while PNG pages stream through encode_to_host_async

    def _decode_png_entries(self, pages, check: bool):
        """Replace the entries of `pages` that are PNG files (bytes, or base64 / data-URI strings as the reference's requests
        carry them) by device tensors decoded on the GPU. Returns (pages, kernels launched)."""
        idx = [i for i, p in enumerate(pages) if isinstance(p, (bytes, bytearray, memoryview, str))]
        if not idx:
            return pages, 0
        from .png_decode import PngError, decode_png_batch, payload_bytes
        try:
            dec = decode_png_batch([payload_bytes(pages[i]) for i in idx], device=self.tower.device, check=check)
        except PngError as e:
            raise PngError(idx[e.index], str(e).split(": ", 1)[-1]) from e
        self.last_png_status = (idx, decode_png_batch.last_status)
        pages = list(pages)
        for i, t in zip(idx, dec):
            pages[i] = t
        return pages, 2

    @torch.no_grad()
    def encode(self, pages):
        pages, n_dec = self._decode_png_entries(pages, check=True)
        pv, grid = self.processor.preprocess_device(pages, out_dtype=torch.bfloat16)
        emb = self.tower(pv, grid_thw=grid)
        self.last_launch_count = 1 + self.tower.last_launch_count + n_dec
        return emb, grid

    def encode_png(self, files):
        """PNG files -> embeddings with the decode on the GPU: only compressed bytes cross PCIe. Raises PngError (with the
        page's index) for a page that cannot be decoded."""
        return self.encode(list(files))

    @torch.no_grad()
    def encode_to_host_async(self, pages, out_host: torch.Tensor):
        """Pipelined end-to-end call for bulk encoding: the H2D copy + preprocess run on an input stream, the tower on the
        caller's stream, the D2H of the embeddings on an output stream, chained by events, so consecutive calls overlap
        their copies with each other's compute. Returns (event that completes when out_host is filled, grid, n_rows);
        the caller must not reuse `out_host` (or its pages) before that event. Pages given as PNG file bytes are decoded by
        kernels on the input stream as well; after the event `last_png_status` = (their indices, int32 status tensor, 0 = ok)."""
        dev = self.tower.device
        if not hasattr(self, "_s_in"):
            # the input stream outranks the tower's: when an SM frees up, a waiting decode / preprocess CTA goes first
            self._s_in, self._s_out = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev)
        if not getattr(self, "_reserved", False) and any(isinstance(p, (bytes, bytearray, memoryview, str)) for p in pages):
            # undecoded pages: their decode runs beside the previous batch's tower on a few SMs that the tower's persistent
            # GEMM grids leave alone from now on (csrc: kocr_set_reserved_sms)
            from . import _lib
            _lib.check(_lib.load().kocr_set_reserved_sms(_lib.context(dev.index if dev.index is not None else torch.cuda.current_device()), self.decode_sms))
            self._reserved = True
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self._s_in):
            # undecoded PNG files: inflate + unfilter on the input stream too, under the previous batch's tower
            pages, n_dec = self._decode_png_entries(pages, check=False)
            pv, grid = self.processor.preprocess_device(pages, out_dtype=torch.bfloat16)
            ev_in = torch.cuda.Event()
            ev_in.record(self._s_in)
        cur.wait_event(ev_in)
        pv.record_stream(cur)
        emb = self.tower(pv, grid_thw=grid)
        self.last_launch_count = 1 + self.tower.last_launch_count + n_dec
        ev_c = torch.cuda.Event()
        ev_c.record(cur)
        self._s_out.wait_event(ev_c)
        with torch.cuda.stream(self._s_out):
            out_host[: emb.shape[0]].copy_(emb, non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self._s_out)
        emb.record_stream(self._s_out)
        return ev_out, grid, emb.shape[0]

    @torch.no_grad()
    def encode_to_host(self, pages, out_host: torch.Tensor | None = None):
        """End-to-end call: host pages in, embeddings in (pinned) host memory out."""
        emb, grid = self.encode(pages)
        if out_host is None:
            out_host = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
        out_host[: emb.shape[0]].copy_(emb, non_blocking=True)
        torch.cuda.current_stream(self.tower.device).synchronize()
        return out_host[: emb.shape[0]], grid
