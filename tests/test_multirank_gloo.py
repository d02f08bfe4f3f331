"""N>1 host logic on CPU: world_size-2 gloo processes shard a page list with the LPT partitioner and gather per-page
results on the host in original order (no data-path collective exists on this path: pages are independent)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from karanta_ocr_b200 import gather_pages, page_cost, shard_pages

SHAPES = [(1288, 995), (1288, 420), (640, 880), (256, 256), (1288, 910), (995, 1288), (1288, 995), (300, 200), (2048, 1583)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [page_cost(h, w, 3136, 12845056) for h, w in SHAPES]
        mine = shard_pages(costs, world)[rank]
        # stand-in for the per-page embeddings (the CUDA encoder needs a GPU): a tensor that encodes the page index
        items = [torch.full((3, 4), float(i)) for i in mine]
        out = gather_pages(items, mine, len(SHAPES))
        t = torch.tensor([sum(costs[i] for i in mine)], dtype=torch.float64)
        loads = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(loads, t)
        if rank == 0:
            q.put((mine, [int(o[0, 0].item()) for o in out], [float(x) for x in loads]))
        else:
            assert out is None
            q.put((mine, None, None))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(r[0] for r in res)
    assert sorted(i for s in shards for i in s) == list(range(len(SHAPES)))   # a partition
    ordered = [r[1] for r in res if r[1] is not None][0]
    assert ordered == list(range(len(SHAPES)))                                   # original page order restored
    loads = [r[2] for r in res if r[2] is not None][0]
    assert max(loads) / min(loads) < 1.35                                        # LPT keeps the two ranks balanced


def test_gather_without_process_group_is_identity():
    out = gather_pages(["b", "a"], [1, 0], 2)
    assert out == ["a", "b"]


def test_cost_model_orders_pages_sensibly():
    c = lambda h, w: page_cost(h, w, 3136, 12845056)
    assert c(2048, 1583) > c(1288, 995) > c(640, 880) > c(256, 256)
