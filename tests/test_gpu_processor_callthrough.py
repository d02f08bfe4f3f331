"""the reference's processor call, karanta/training/pipeline_steps.py:289-294

    inputs = self.processor(text=[text], images=[main_image], padding=False, return_tensors="pt")

executed through transformers' Qwen2VLProcessor.__call__ (processing_qwen2_vl.py:61-129) with the image processor swapped
for KarantaImageProcessor the way INTEGRATION.md says (`processor.image_processor = KarantaImageProcessor(...)`), and
compared item by item with the unmodified processor: input_ids (so the <|image_pad|> expansion count), attention_mask,
pixel_values (bit-exact) and image_grid_thw."""
import numpy as np
import pytest
import torch

from tests.synth import synth_page
from tests.tiny_qwen import PROMPT, hf_processor

pytestmark = pytest.mark.gpu
CKPT_MAX = 12845056


def _pil(page_chw, gray=False):
    from PIL import Image
    im = Image.fromarray(np.ascontiguousarray(page_chw.transpose(1, 2, 0)))
    return im.convert("L") if gray else im


@pytest.mark.parametrize("hw,gray", [((1288, 995), False), ((644, 455), True), ((300, 200), False)])
def test_processor_call_site_with_swapped_image_processor(hw, gray):
    pytest.importorskip("transformers")
    from karanta_ocr_b200 import KarantaImageProcessor
    proc = hf_processor(3136, CKPT_MAX)
    img = _pil(synth_page(hw[0], hw[1], 77), gray)
    ref = proc(text=[PROMPT], images=[img], padding=False, return_tensors="pt")
    proc.image_processor = KarantaImageProcessor(min_pixels=3136, max_pixels=CKPT_MAX)   # the INTEGRATION.md one-liner
    got = proc(text=[PROMPT], images=[img], padding=False, return_tensors="pt")
    assert set(got.keys()) == set(ref.keys())
    pad = proc.image_token_id
    n_pad = int((ref["input_ids"] == pad).sum())
    assert n_pad == int(ref["image_grid_thw"][0].prod()) // 4 and n_pad > 1
    for k in ref.keys():
        assert got[k].dtype == ref[k].dtype and got[k].device == ref[k].device, k
        assert torch.equal(got[k], ref[k]), k
    # what Tokenizer.__call__ stores afterwards (pipeline_steps.py:361-371)
    assert got["pixel_values"].shape == (int(got["image_grid_thw"][0].prod()), 1176)


def test_two_images_two_prompts_and_call_time_size():
    pytest.importorskip("transformers")
    from karanta_ocr_b200 import KarantaImageProcessor
    proc = hf_processor(3136, CKPT_MAX)
    imgs = [_pil(synth_page(420, 322, 5)), _pil(synth_page(256, 700, 6))]
    ref = proc(text=[PROMPT, PROMPT], images=imgs, padding=True, return_tensors="pt")
    proc.image_processor = KarantaImageProcessor(min_pixels=3136, max_pixels=CKPT_MAX)
    got = proc(text=[PROMPT, PROMPT], images=imgs, padding=True, return_tensors="pt")
    for k in ref.keys():
        assert torch.equal(got[k], ref[k]), k
    # call-time size= means what it means upstream (it used to be silently ignored)
    kip = proc.image_processor
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor
    small = {"shortest_edge": 3136, "longest_edge": 28 * 28 * 64}
    a = kip(images=imgs, size=small, return_tensors="pt")
    b = Qwen2VLImageProcessor(min_pixels=3136, max_pixels=CKPT_MAX)(images=imgs, size=small, return_tensors="pt")
    assert torch.equal(a["image_grid_thw"], b["image_grid_thw"]) and torch.equal(a["pixel_values"], b["pixel_values"])
    with pytest.raises(ValueError):
        kip(images=imgs, size={"height": 224, "width": 224})
    with pytest.raises(ValueError):
        kip(images=imgs, patch_size=16)
    with pytest.raises(ValueError):
        kip(images=imgs, image_mean=[0.5, 0.5, 0.5])
