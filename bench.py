#!/usr/bin/env python
"""bench.py - pages/sec of the page-image hot path (preprocess + ViT encode) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the box's host cores

Workload (BASELINE.json configs[1], "C2"): olmOCR-7B = Qwen2-VL-7B vision tower (depth 32, D 1280, 16 heads, MLP 5120,
out 3584), 64 synthetic letter pages of uint8 [3,1288,995] per step per GPU -> smart_resize 1288x1008 -> 6624 patches
per page, random-init weights (no checkpoints offline), bf16 compute.  A step = one pass of the hot path over one batch.

Prints ONE JSON line (rank 0).  `value`: pages already resident in HBM when the timed region starts.  `e2e`: the same
through the public API with host buffers, H2D of the pages and D2H of the embeddings inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PAGE_H, PAGE_W = 1288, 995
MIN_PIXELS, MAX_PIXELS = 3136, 12845056
PAGES_PER_STEP = 64
METRIC = "pages/sec (preprocess+ViT encode)"
UNIT = "pages/s"
WORKLOAD = ("C2: Qwen2-VL-7B vision tower (depth 32, D 1280, 16 heads, mlp 5120, out 3584), 64 synthetic letter pages "
            "u8[3,1288,995] per step per GPU -> 1288x1008 -> 6624 patches/page")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


MIXED_SHAPES = [(1288, 420), (640, 880), (256, 256), (1288, 910), (1288, 995), (995, 1288)]
WORKLOADS = {
    "c2": WORKLOAD,
    "c3": ("C3: Qwen2.5-VL-7B vision tower (depth 32, D 1280, 16 heads, gated mlp 3420, out 3584, 112 px windows, full attention "
           "in blocks 7/15/23/31), 64 synthetic letter pages u8[3,1288,995] per step per GPU"),
    "c4": ("C4: Qwen2-VL-7B vision tower, mixed-aspect varlen batch of 64 pages per step per GPU drawn (seeded) from "
           "1288x420, 640x880, 256x256, 1288x910, 1288x995, 995x1288"),
    "c5": ("C5: bulk job of letter pages u8[3,1288,995] (C2's generator) through the Qwen2-VL-7B tower, page-sharded over the GPUs, "
           "64-page batches, host pages in -> host embeddings out"),
}


def make_pages(n, distinct=8, workload="c2"):
    from tests.synth import synth_page
    if workload == "c4":
        rng = np.random.default_rng(4)
        picks = rng.integers(0, len(MIXED_SHAPES), n)
        cache = {}
        return [cache.setdefault(int(k), synth_page(*MIXED_SHAPES[int(k)], 4000 + int(k))) for k in picks]
    base = [synth_page(PAGE_H, PAGE_W, 1234 + i) for i in range(min(n, distinct))]
    return [base[i % len(base)] for i in range(n)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = top half of the clock samples taken while the kernels ran
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_page(dtype=torch.float32, use_hf=True):
    """The reference's CPU path for ONE letter page of the C2 workload, nothing sampled or extrapolated: the transformers
    image processor, then the full 32-block Qwen2-VL-7B vision tower (sdpa attention) on all host threads. Falls back to
    the oracle port (same arithmetic, oracle/) only when transformers cannot be imported.
    Returns (run_once() -> seconds for the page, kind, description)."""
    torch.set_num_threads(os.cpu_count())
    page = make_pages(1)[0]
    kind, proc, model = "port", None, None
    if use_hf:
        try:
            os.environ.setdefault("HF_HUB_OFFLINE", "1")
            from transformers.models.qwen2_vl.configuration_qwen2_vl import Qwen2VLVisionConfig
            from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor
            from transformers.models.qwen2_vl.modeling_qwen2_vl import Qwen2VisionTransformerPretrainedModel
            proc = Qwen2VLImageProcessor(min_pixels=MIN_PIXELS, max_pixels=MAX_PIXELS)
            c = Qwen2VLVisionConfig(depth=32, embed_dim=1280, hidden_size=3584, mlp_ratio=4, num_heads=16)
            c._attn_implementation = "sdpa"
            torch.manual_seed(0)
            model = Qwen2VisionTransformerPretrainedModel(c).eval().to(dtype)
            kind = "reference"
        except Exception:
            proc = model = None
    state = {}

    def run_once():
        t0 = time.perf_counter()
        with torch.no_grad():
            if model is not None:
                f = proc(images=[torch.from_numpy(page)], return_tensors="pt")
                out = model(f["pixel_values"].to(dtype), grid_thw=f["image_grid_thw"])
                out = getattr(out, "pooler_output", out)
            else:
                from oracle import preprocess_oracle as po
                from oracle import vision_oracle as vo
                if "sd" not in state:
                    state["cfg"] = vo.qwen2_vl_7b(32)
                    state["sd"] = vo.init_weights(state["cfg"], seed=0)
                pv_np, grid = po.preprocess([page], MIN_PIXELS, MAX_PIXELS, po.RESIZE_ATEN)
                out = vo.tower_forward(state["cfg"], state["sd"], torch.from_numpy(pv_np), grid, dtype)
        assert tuple(out.shape) == (1656, 3584), out.shape
        return time.perf_counter() - t0
    desc = (f"1 letter page per step (of the workload's 64) on {os.cpu_count()} host threads, {str(dtype).replace('torch.', '')}: image processor "
            f"+ the full 32-block tower, measured (nothing extrapolated), via "
            f"{'transformers ' + __import__('transformers').__version__ if kind == 'reference' else 'the oracle port'}")
    return run_once, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run_once, kind, desc = cpu_reference_page()
    t_first = run_once()
    # every timed step is a real full-depth page; warm-up is cut to the first page only when W + K pages would not fit the budget
    budget = float(os.environ.get("KOCR_REF_BUDGET_S", "1200"))
    warm = max(args.warmup, 1) if t_first * (args.warmup + args.steps) <= budget else 1
    for _ in range(warm - 1):
        run_once()
    t0 = time.perf_counter()
    ts = [run_once() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    spp = wall / args.steps
    value = 1.0 / spp
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "warmup_done": warm, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_step": 1,
                       "note": "a reference step is ONE page of the 64-page batch (bounded sample); pages/s = steps / wall time of the timed steps"},
            "ms_per_page_min_max": [min(ts) * 1e3, max(ts) * 1e3],
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def library_baselines(pages=16, steps=3, timeout_s=300):
    """The two library towers a user of the reference could run on this same GPU today (last row):
    transformers + flash-attn 2 / sdpa, and vLLM's Qwen2VisionTransformer with its FA4 back-end; 16 C2 pages, bf16, tower
    only. Each runs in its own process (vLLM initialises a process group) after this process has finished timing."""
    out = []
    for tool in ("hf_gpu_baseline.py", "vllm_gpu_baseline.py"):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(pages), str(steps)], capture_output=True,
                               text=True, timeout=timeout_s)
            got = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
            out += got if got else [{"impl": tool, "unavailable": (r.stderr or "no output")[-300:]}]
        except Exception as e:
            out.append({"impl": tool, "unavailable": f"{type(e).__name__}: {str(e)[:200]}"})
    return out


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import ctypes as C

    import torch.distributed as dist

    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, _lib, presets, smart_resize

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = SimpleNamespace(**presets.preset("qwen2_5_vl_7b" if args.workload == "c3" else "qwen2_vl_7b"))
    tower = KarantaVisionTower(vars(cfg), device=dev)
    tower.load_state_dict(presets.random_state_dict(vars(cfg), seed=0))
    enc = PageEncoder(tower, MIN_PIXELS, MAX_PIXELS)
    n_pages = args.pages
    pages = make_pages(n_pages, workload=args.workload)
    grid_ref = []
    for p in pages:
        rh, rw = smart_resize(p.shape[1], p.shape[2], 28, MIN_PIXELS, MAX_PIXELS)
        grid_ref.append([1, rh // 14, rw // 14])
    fl = presets.flops_per_batch(vars(cfg), grid_ref)
    flops_step = fl["total"]
    N = int(sum(g[1] * g[2] for g in grid_ref))
    flops_attn_launch = fl["attention"] / cfg.depth

    # device-resident inputs for `value`; pinned host inputs for `e2e`
    d_pages = [torch.from_numpy(p).to(dev) for p in pages]
    h_pages = [torch.from_numpy(p).pin_memory() for p in pages]
    out_hosts = [torch.empty((N // 4, cfg.out_hidden), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    out_host = out_hosts[0]
    lib = _lib.load()
    ctx = _lib.context(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()                      # all D2H copies of the timed steps have landed in host memory
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def step_resident():
        enc.encode(d_pages)

    pending = []

    def step_e2e():
        # public bulk API: H2D of this step's pages and D2H of its embeddings are inside the timed region, pipelined
        # across steps (at most two steps in flight; the timed region ends with every D2H complete)
        if len(pending) >= 2:
            pending.pop(0).synchronize()
        ev, _, _ = enc.encode_to_host_async(h_pages, out_hosts[step_e2e.k % 2])
        step_e2e.k += 1
        pending.append(ev)
    step_e2e.k = 0

    def drain_e2e():
        while pending:
            pending.pop(0).synchronize()

    for _ in range(args.warmup):
        step_resident()
    launches_per_step = enc.last_launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.kocr_profile_begin(ctx)
    ms = timed(step_resident, args.steps)
    ncls = 16  # the library reports as many classes as it has (names past the last one are empty)
    cls_ms = (C.c_double * ncls)()
    cls_n = (C.c_int64 * ncls)()
    lib.kocr_profile_end(ctx, ncls, cls_ms, cls_n)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    drain_e2e()

    def e2e_steps_all():
        step_e2e()
    ms_e2e = timed(e2e_steps_all, args.steps, after=drain_e2e)
    clocks = sampler.stop() if rank == 0 else None

    # ---- extras on every rank (they shard work): the C5 bulk job (strong scaling) and the same pipeline fed
    # with undecoded PNG files (row f2: decode on the GPU inside the timed region)
    extras = {}
    if args.workload == "c2" and not args.no_c5:
        job_pages = args.job_pages if world > 1 else min(args.job_pages, 1024)
        ms_j, _, n_mine, _ = bulk_job(enc, vars(cfg), world, rank, dev, job_pages, 1, 1)
        extras["c5_bulk_job"] = {"value": job_pages / (ms_j / 1e3), "unit": UNIT, "scaling": "strong", "job_pages": job_pages,
                                 "seconds": ms_j / 1e3, "pages_per_rank": n_mine,
                                 "what": "one job of job_pages letter pages sharded over the ranks (LPT, no collective), 64-page batches, "
                                         "pinned host pages in -> host embeddings out; 8192 pages at N > 1, 1024 at N = 1"}
        png_pages = 512 * world
        ms_p, _, _, h2d_p = bulk_job(enc, vars(cfg), world, rank, dev, png_pages, 1, 1, png=True)
        extras["e2e_from_png"] = {"value": png_pages / (ms_p / 1e3), "unit": UNIT, "job_pages": png_pages, "h2d_bytes_per_rank": h2d_p,
                                  "what": "same job with the pages given as RGB PNG files: inflate + scan-line filters on the GPU inside the "
                                          "timed region, only compressed bytes cross PCIe"}

    if rank == 0:
        pk = peaks()
        value = world * n_pages * args.steps / (ms / 1e3)
        e2e = world * n_pages * args.steps / (ms_e2e / 1e3)
        names = [lib.kocr_profile_class_name(i).decode() for i in range(ncls)]
        per_class = {names[i]: {"ms_per_launch": cls_ms[i] / max(cls_n[i], 1), "launches": int(cls_n[i]), "ms_total": cls_ms[i]}
                     for i in range(ncls) if cls_n[i] and names[i]}
        dom = max(per_class, key=lambda k: per_class[k]["ms_total"])
        D, F = cfg.embed_dim, cfg.mlp_hidden
        fc1_mult = 2.0 if cfg.arch == "qwen2_5_vl" else 1.0  # gate and up projections
        flops_by_class = {"attention": flops_attn_launch, "gemm_qkv_rope": 2.0 * N * D * 3 * D, "gemm_proj": 2.0 * N * D * D,
                          "gemm_fc1": 2.0 * N * D * F * fc1_mult, "gemm_fc2": 2.0 * N * D * F, "gemm_patch_embed": 2.0 * N * 1176 * D}
        if cfg.arch == "qwen2_5_vl":  # full-attention layers only: the windowed ones are a different kernel shape with its own class
            n_full = len(cfg.fullatt_block_indexes)
            flops_by_class["attention"] = sum(fl["attention_per_layer"][i] for i in cfg.fullatt_block_indexes) / max(n_full, 1)
        for k, v in per_class.items():
            if k in flops_by_class:
                v["tflops"] = flops_by_class[k] / (v["ms_per_launch"] * 1e-3) / 1e12
        if "attention_windowed" in per_class:  # HBM-bound shape: qkv in (3 S D bf16) + output (S D bf16), ~0.13 TFLOP per layer
            v = per_class["attention_windowed"]
            v["bytes_per_launch"] = 4 * N * D * 2
            v["gbs"] = v["bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9
            v["frac_of_hbm"] = v["gbs"] / pk["hbm_gbs"]
            v["bound"] = "hbm"
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp) and n_pages == PAGES_PER_STEP and args.workload == "c2":  # ncu dram__bytes_read+write per launch at this batch size
            tj = json.load(open(tp))
            traffic, traffic_src = tj["dram_bytes_per_launch"].get(dom), tj["source"]
        if dom in flops_by_class:
            achieved = flops_by_class[dom] / (per_class[dom]["ms_per_launch"] * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                        "flops_per_launch": flops_by_class[dom]}
        else:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None}
        pre = per_class.get("preprocess")
        pre_bytes = int(sum(p.size for p in pages)) + N * 1176 * 2
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "pages_per_step_per_gpu": n_pages, "patches_per_step_per_gpu": N, "parallelism": f"page-sharded replicas x{world}, no collective",
                       "l2": "per-step working set ~11 GB (activations 1.09 GB per tensor) >> 126 MB L2, no flush needed",
                       "weights": "seeded random init (no checkpoints offline)"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(sum(p.numel() for p in h_pages)), "d2h_bytes_per_step": int(out_host.numel() * 2)},
            "gpu_launches": int(launches_per_step * args.steps),
            "tensor_pipe": {"model_tflops": flops_step * value / (world * n_pages) / 1e12 / 1.0,
                            "frac_of_sustained_peak": flops_step * value / (world * n_pages) / 1e12 / pk["tflops_sustained"],
                            "frac_of_burst_peak": flops_step * value / (world * n_pages) / 1e12 / pk["tflops_burst"],
                            "flops_per_page": flops_step / n_pages},
            "roofline": roofline,
            "kernels": per_class,
            "preprocess_hbm": ({"achieved_gbs": pre_bytes / (pre["ms_per_launch"] * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                                "frac": pre_bytes / (pre["ms_per_launch"] * 1e-3) / 1e9 / pk["hbm_gbs"], "bytes_per_launch": pre_bytes}
                               if pre else None),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
            run_once, kind, desc = cpu_reference_page()
            run_once()                       # warm-up page (thread pools, oneDNN primitive caches)
            spp = run_once()                 # one full-depth page: the bounded CPU sample (~10-30 s)
            line["cpu_baseline"] = {"value": 1.0 / spp, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": desc}
        if extras:
            line.setdefault("extra", {}).update(extras)
        if world == 1 and args.library_baselines and args.workload == "c2":
            del d_pages, h_pages
            torch.cuda.empty_cache()
            line.setdefault("extra", {}).update({
                "library_baselines_same_gpu": library_baselines(),
                "library_baselines_note": "tower only, 16 C2 pages per call, bf16, random weights; 'this repo' line in the same list for the like-for-like"})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- C5: bulk job, strong scaling
def bulk_job(enc, cfg, world, rank, dev, total, steps=1, warmup=1, png=False):
    """One job of `total` letter pages (C2's generator), page-sharded over the ranks with shard_pages (no collective), each
    rank streaming its shard through PageEncoder.encode_to_host_async in 64-page batches: pinned host pages in (or, with
    png=True, the pages' PNG files, decoded on the GPU - SURVEY.md section 8 row f2), embeddings in pinned host memory out.
    Returns (ms for `steps` jobs as the max over ranks, launches, pages of this rank, h2d bytes per job of this rank)."""
    import torch.distributed as dist

    from karanta_ocr_b200 import page_cost, shard_pages
    if png:
        import io

        from PIL import Image
        pool = []
        for p in make_pages(8):  # what pdftoppm -png hands over: 8-bit RGB PNG files
            buf = io.BytesIO()
            Image.fromarray(np.ascontiguousarray(p.transpose(1, 2, 0))).save(buf, format="PNG")
            pool.append(buf.getvalue())
        h, w = PAGE_H, PAGE_W
        nbytes = [len(b) for b in pool]
    else:
        pool = [torch.from_numpy(p).pin_memory() for p in make_pages(PAGES_PER_STEP)]  # the job cycles over 64 distinct pinned pages
        h, w = pool[0].shape[1], pool[0].shape[2]
        nbytes = [int(t.numel()) for t in pool]
    cost = page_cost(h, w, MIN_PIXELS, MAX_PIXELS)
    mine = shard_pages([cost] * total, world)[rank]
    rows = (92 * 72) // 4
    outs = [torch.empty((PAGES_PER_STEP * rows, cfg["out_hidden"]), dtype=torch.bfloat16).pin_memory() for _ in range(2)]

    def job(indices):
        pending, k, launches = [], 0, 0
        for b in range(0, len(indices), PAGES_PER_STEP):
            batch = [pool[i % len(pool)] for i in indices[b:b + PAGES_PER_STEP]]
            if len(pending) >= 2:
                pending.pop(0).synchronize()
            ev, _, _ = enc.encode_to_host_async(batch, outs[k % 2])
            launches += enc.last_launch_count
            pending.append(ev)
            k += 1
        for ev in pending:
            ev.synchronize()
        return launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        job(mine[:PAGES_PER_STEP])  # warm-up = one batch per rank, not a whole job
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(steps):
        launches += job(mine)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    barrier()
    return float(ms.item()), launches, len(mine), int(sum(nbytes[i % len(pool)] for i in mine))


def run_bulk_job(args):
    """one job of `--job-pages` letter pages, strong scaling. A step is the whole job."""
    import torch.distributed as dist

    from karanta_ocr_b200 import KarantaVisionTower, PageEncoder, presets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = presets.preset("qwen2_vl_7b")
    tower = KarantaVisionTower(cfg, device=dev)
    tower.load_state_dict(presets.random_state_dict(cfg, seed=0))
    enc = PageEncoder(tower, MIN_PIXELS, MAX_PIXELS)
    total = args.job_pages
    rows = (92 * 72) // 4
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, n_mine, h2d = bulk_job(enc, cfg, world, rank, dev, total, args.steps, args.warmup, png=args.png)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        value = total * args.steps / (ms / 1e3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOADS["c5"], "job_pages": total, "pages_per_rank": n_mine, "batch_pages": PAGES_PER_STEP,
                           "parallelism": f"page-sharded x{world} (LPT), no collective, host-side results",
                           "pages": ("8 distinct RGB PNG files cycled, decoded on the GPU (inflate + scan-line filters)" if args.png
                                     else "cycled from a pool of 64 distinct pinned host pages"), "weights": "seeded random init"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n_mine * rows * cfg["out_hidden"] * 2)},
                "gpu_launches": int(launches), "clocks": clocks,
                "note": "value is measured end to end (host pages in, host embeddings out); there is no device-resident variant of a bulk job"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kocr", choices=["kocr", "reference"])
    ap.add_argument("--pages", type=int, default=PAGES_PER_STEP, help="pages per step per GPU (C2 = 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baselines", dest="library_baselines", action="store_false",
                    help="skip timing the transformers (flash-attn 2 / sdpa) and vLLM (FA4) towers on this GPU after the run (line['extra'])")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = the metric's configuration (default); c3 / c4 = the other BASELINE.json configs, for the record; "
                         "c5 = one bulk job of --job-pages pages sharded over the ranks (strong scaling)")
    ap.add_argument("--job-pages", type=int, default=8192, help="pages in the C5 bulk job (--workload c5, and the c5_bulk_job extra of the default run)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 bulk-job and PNG-input extras of the default run")
    ap.add_argument("--png", action="store_true", help="--workload c5: feed the job with PNG files decoded on the GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_bulk_job(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
