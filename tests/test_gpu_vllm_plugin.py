"""SURVEY.md section 8 row f1, to the row's definition: the plugin inside a LIVE vLLM engine. A tiny Qwen2-VL checkpoint directory
is written locally (tests/tiny_qwen.py; vLLM fills the weights with its seeded dummy loader), `vllm.LLM` is started the way
`vllm serve` builds its engine - once stock, once with the `vllm.general_plugins` entry point active - and one page is sent
through each: the engine process must hold KarantaVllmVisual as `model.visual`, the multimodal processor must run
KarantaImageProcessor, and the image embeddings the language model receives must agree (cosine >= 0.999).
Each engine runs in a child process (one engine per process, like production); skipped where vLLM cannot start an engine."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, {root!r})
os.environ["VLLM_ENABLE_V1_MULTIPROCESSING"] = "0"      # engine core in this process, so the model object can be inspected
os.environ["HF_HUB_OFFLINE"] = "1"
import numpy as np, torch
from PIL import Image
from tests.synth import synth_page
from tests.tiny_qwen import PROMPT
from vllm import LLM, SamplingParams

llm = LLM(model={ckpt!r}, load_format="dummy", enforce_eager=True, max_model_len=4096, gpu_memory_utilization=0.25, dtype="bfloat16",
          limit_mm_per_prompt={{"image": 1, "video": 0}}, seed=0, disable_log_stats=True)
core = llm.llm_engine.engine_core.engine_core
model = core.model_executor.driver_worker.worker.get_model()
page = Image.fromarray(synth_page(420, 322, 7).transpose(1, 2, 0))
captured = {{}}
orig = type(model)._process_image_input
def spy(self, image_input):
    out = orig(self, image_input)
    captured["emb"] = torch.cat([o.float().cpu() for o in out])
    captured["pixel_values_absmax"] = float(image_input["pixel_values"].float().abs().max())
    return out
type(model)._process_image_input = spy
outs = llm.generate([{{"prompt": PROMPT, "multi_modal_data": {{"image": page}}}}], SamplingParams(temperature=0.0, max_tokens=8, logprobs=1))
proc = llm.llm_engine.input_processor if hasattr(llm.llm_engine, "input_processor") else None
from vllm.multimodal import MULTIMODAL_REGISTRY
info = {{
    "model_class": type(model).__name__, "model_module": type(model).__module__,
    "visual_class": type(model.visual).__name__,
    "tokens": list(outs[0].outputs[0].token_ids),
    "emb_shape": list(captured["emb"].shape),
}}
np.save({out!r} + ".emb.npy", captured["emb"].numpy())
# which image processor did the multimodal processor of this engine use?
mm_proc = MULTIMODAL_REGISTRY.create_processor(llm.llm_engine.model_config) if hasattr(MULTIMODAL_REGISTRY, "create_processor") else None
if mm_proc is not None:
    info["image_processor_class"] = type(mm_proc.info.get_image_processor()).__name__
json.dump(info, open({out!r}, "w"))
print("CHILD_OK", info)
'''


def _run(tmp_path, ckpt, tag, plugin: bool):
    out = str(tmp_path / f"{tag}.json")
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")   # the in-tree dist-info makes the entry point discoverable
    env["KOCR_VLLM_PLUGIN"] = "1" if plugin else "0"      # the image's other general plugins stay active in both runs
    env["VLLM_LOGGING_LEVEL"] = "WARNING"
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, ckpt=ckpt, out=out)], env=env, capture_output=True, text=True, timeout=700)
    tail = (r.stdout + r.stderr)[-3000:]
    if r.returncode != 0 or not os.path.exists(out):
        return None, tail
    return json.load(open(out)), np.load(out + ".emb.npy")


@pytest.mark.parametrize("arch", ["qwen2_vl", "qwen2_5_vl"])
def test_plugin_inside_live_engine(tmp_path, arch):
    pytest.importorskip("vllm")
    from tests.tiny_qwen import write_tiny_checkpoint
    ckpt = write_tiny_checkpoint(str(tmp_path / f"ckpt_{arch}"), arch)
    stock, stock_emb = _run(tmp_path, ckpt, "stock", plugin=False)
    if stock is None:
        pytest.skip("vLLM could not start an engine on the tiny checkpoint without the plugin:\n" + stock_emb)
    mine, mine_emb = _run(tmp_path, ckpt, "plugin", plugin=True)
    assert mine is not None, "engine with the plugin failed:\n" + str(mine_emb)
    print("stock:", stock, "\nplugin:", mine)
    assert stock["visual_class"] in ("Qwen2VisionTransformer", "Qwen2_5_VisionTransformer")
    assert mine["model_class"].startswith("Karanta") and mine["model_module"] == "karanta_ocr_b200.vllm_plugin"
    assert mine["visual_class"] == "KarantaVllmVisual"
    assert mine.get("image_processor_class", "KarantaImageProcessor") == "KarantaImageProcessor"
    assert stock.get("image_processor_class", "Qwen2VLImageProcessor") != "KarantaImageProcessor"
    assert mine["emb_shape"] == stock["emb_shape"] == [30 * 24 // 4, 256]      # 420x322 -> 420x336 -> grid 30x24
    a, b = mine_emb.astype(np.float64).ravel(), stock_emb.astype(np.float64).ravel()
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    rel = float(np.abs(a - b).max() / np.abs(b).max())
    print(f"vllm live engine {arch}: image embeddings cosine {cos:.6f} max-rel {rel:.4f}; tokens stock {stock['tokens']} plugin {mine['tokens']}")
    assert cos >= 0.999 and rel <= 0.03
