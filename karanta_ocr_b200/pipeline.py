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
    def encode(self, pages):
        pv, grid = self.processor.preprocess_device(pages, out_dtype=torch.bfloat16)
        emb = self.tower(pv, grid_thw=grid)
        self.last_launch_count = 1 + self.tower.last_launch_count
        return emb, grid

    @torch.no_grad()
    def encode_to_host(self, pages, out_host: torch.Tensor | None = None):
        """End-to-end call: host pages in, embeddings in (pinned) host memory out."""
        emb, grid = self.encode(pages)
        if out_host is None:
            out_host = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
        out_host[: emb.shape[0]].copy_(emb, non_blocking=True)
        torch.cuda.current_stream(self.tower.device).synchronize()
        return out_host[: emb.shape[0]], grid
