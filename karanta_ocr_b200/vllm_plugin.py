"""vLLM plugin: the served path of karanta-ocr picks this repo's kernels up without a code change.

Production runs `vllm serve <model> ...` as a child process (karanta/pipeline.py:707-742; one server per GPU in
scripts/start_multiple_vllm_servers.sh:272-317) and posts pages to it (bulk_processing/workers/inference_worker.py:327-339).
vLLM loads every entry point of the group `vllm.general_plugins` in each of its processes (vllm/plugins/__init__.py:
load_general_plugins) before it resolves the model class, so

    [project.entry-points."vllm.general_plugins"]
    karanta_ocr_b200 = "karanta_ocr_b200.vllm_plugin:register"

(pyproject.toml; for a source checkout the same entry point ships as karanta_ocr_b200-*.dist-info next to the package, found
as soon as the checkout is on PYTHONPATH) makes `register()` run there. It re-registers the two architectures karanta-ocr
serves - Qwen2VLForConditionalGeneration (olmOCR-7B-0225) and Qwen2_5_VLForConditionalGeneration (olmOCR-7B-0725, karanta's
own Qwen2.5-VL-3B fine-tunes, configs/training/ocr/*.yaml:2) - with subclasses of vLLM's own model classes that differ in
two places only:

  * the vision tower: after the weights are in place (checkpoint, sharded or dummy loader alike) `self.visual` is swapped for
    KarantaVllmVisual, built from the loaded parameters (vllm_adapter.replace_vllm_visual); `_process_image_input` then calls
    `self.visual(pixel_values, grid_thw=...)` exactly as before (vllm/model_executor/models/qwen2_vl.py:1376);
  * the multimodal processor: the HF processor vLLM builds for the model gets its `image_processor` replaced by
    KarantaImageProcessor with the checkpoint's own min / max pixels, so the resize / normalise / patchify of every request
    runs in the fused kernel instead of on the API server's CPU cores (24-38 ms per page there).

Switches (environment, read in every vLLM process): KOCR_VLLM_TOWER=0 keeps vLLM's tower, KOCR_VLLM_PREPROCESS=0 keeps the
CPU image processor, KOCR_VLLM_PLUGIN=0 makes register() a no-op.
"""
from __future__ import annotations

import os

_ARCHS = {
    "Qwen2VLForConditionalGeneration": "karanta_ocr_b200.vllm_plugin:KarantaQwen2VLForConditionalGeneration",
    "Qwen2_5_VLForConditionalGeneration": "karanta_ocr_b200.vllm_plugin:KarantaQwen2_5_VLForConditionalGeneration",
}


def _on(name: str) -> bool:
    return os.environ.get(name, "1") not in ("0", "false", "False", "off")


def register() -> None:
    """Entry point of the `vllm.general_plugins` group. Idempotent (vLLM may call it once per process, several processes)."""
    if not _on("KOCR_VLLM_PLUGIN"):
        return
    from vllm import ModelRegistry
    for arch, target in _ARCHS.items():
        ModelRegistry.register_model(arch, target)   # lazy "<module>:<class>" form: no CUDA initialisation at import


def swap_image_processor(hf_processor):
    """`processor.image_processor = KarantaImageProcessor(...)` with the limits the checkpoint's
    preprocessor_config.json gave the stock processor. Returns the processor it was given."""
    from .image_processor import KarantaImageProcessor
    ip = hf_processor.image_processor
    if isinstance(ip, KarantaImageProcessor):
        return hf_processor
    size = getattr(ip, "size", None) or {}
    minp = size.get("shortest_edge", getattr(ip, "min_pixels", None))
    maxp = size.get("longest_edge", getattr(ip, "max_pixels", None))
    backend = "pil" if type(ip).__name__.endswith("Pil") else "torchvision"
    hf_processor.image_processor = KarantaImageProcessor(min_pixels=minp, max_pixels=maxp, resize_backend=backend)
    return hf_processor


def _ensure_tower(model) -> None:
    """Swap `model.visual` for the B200 tower once its weights are loaded (first use inside the engine process)."""
    from .vllm_adapter import KarantaVllmVisual, replace_vllm_visual
    v = getattr(model, "visual", None)
    if v is None or isinstance(v, KarantaVllmVisual) or not _on("KOCR_VLLM_TOWER"):
        return
    replace_vllm_visual(model)


def _build_classes():
    """The subclasses are created on first attribute access so that importing this module (which vLLM does in every process,
    also ones that never build a model) stays free of vLLM model imports."""
    from vllm.model_executor.models import qwen2_5_vl as q25
    from vllm.model_executor.models import qwen2_vl as q2
    from vllm.multimodal import MULTIMODAL_REGISTRY

    class KarantaQwen2VLProcessingInfo(q2.Qwen2VLProcessingInfo):
        def get_hf_processor(self, **kwargs):
            proc = super().get_hf_processor(**kwargs)
            return swap_image_processor(proc) if _on("KOCR_VLLM_PREPROCESS") else proc

    class KarantaQwen2_5_VLProcessingInfo(q25.Qwen2_5_VLProcessingInfo):
        def get_hf_processor(self, **kwargs):
            proc = super().get_hf_processor(**kwargs)
            return swap_image_processor(proc) if _on("KOCR_VLLM_PREPROCESS") else proc

    @MULTIMODAL_REGISTRY.register_processor(q2.Qwen2VLMultiModalProcessor, info=KarantaQwen2VLProcessingInfo,
                                            dummy_inputs=q2.Qwen2VLDummyInputsBuilder)
    class KarantaQwen2VLForConditionalGeneration(q2.Qwen2VLForConditionalGeneration):
        def load_weights(self, weights):
            loaded = super().load_weights(weights)
            _ensure_tower(self)
            return loaded

        def _process_image_input(self, image_input):
            _ensure_tower(self)   # loaders that never call load_weights (load_format="dummy") end up here first
            return super()._process_image_input(image_input)

    @MULTIMODAL_REGISTRY.register_processor(q25.Qwen2_5_VLMultiModalProcessor, info=KarantaQwen2_5_VLProcessingInfo,
                                            dummy_inputs=q25.Qwen2_5_VLDummyInputsBuilder)
    class KarantaQwen2_5_VLForConditionalGeneration(q25.Qwen2_5_VLForConditionalGeneration):
        def load_weights(self, weights):
            loaded = super().load_weights(weights)
            _ensure_tower(self)
            return loaded

        def _process_image_input(self, image_input):
            _ensure_tower(self)
            return super()._process_image_input(image_input)

    return {c.__name__: c for c in (KarantaQwen2VLProcessingInfo, KarantaQwen2_5_VLProcessingInfo,
                                    KarantaQwen2VLForConditionalGeneration, KarantaQwen2_5_VLForConditionalGeneration)}


_classes = None


def __getattr__(name):   # PEP 562: vLLM resolves "<module>:<class>" with getattr(module, class)
    global _classes
    if name.startswith("Karanta"):
        if _classes is None:
            _classes = _build_classes()
            for c in _classes.values():
                c.__module__ = __name__
                c.__qualname__ = c.__name__
        if name in _classes:
            return _classes[name]
    raise AttributeError(name)
