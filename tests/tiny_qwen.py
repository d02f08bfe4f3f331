"""A tiny local stand-in for a Qwen2-VL checkpoint's text side: a word-level tokenizer that knows the vision special
tokens, and a transformers Qwen2VLProcessor built around it. No network, no checkpoint."""
SPECIALS = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "<|vision_start|>", "<|vision_end|>", "<|image_pad|>", "<|video_pad|>"]
WORDS = ["system", "user", "assistant", "Read", "this", "page", ".", "Return", "the", "text", "of", "document"]
PROMPT = "<|im_start|> user <|vision_start|> <|image_pad|> <|vision_end|> Return the text of this page . <|im_end|> <|im_start|> assistant"


def tiny_tokenizer():
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {w: i for i, w in enumerate(SPECIALS + WORDS + ["[UNK]"])}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    return PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]", pad_token="<|endoftext|>",
                                   additional_special_tokens=SPECIALS)


def hf_processor(min_pixels=3136, max_pixels=12845056):
    """Qwen2VLProcessor(image_processor=<transformers' own>, tokenizer=<tiny>): what AutoProcessor.from_pretrained
    hands karanta-ocr (karanta/training/data.py:188,208)."""
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor
    from transformers.models.qwen2_vl.processing_qwen2_vl import Qwen2VLProcessor
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
    return Qwen2VLProcessor(image_processor=Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels),
                            tokenizer=tiny_tokenizer(), video_processor=Qwen2VLVideoProcessor())


def write_tiny_checkpoint(path, arch="qwen2_vl", min_pixels=3136, max_pixels=12845056):
    """A checkpoint DIRECTORY (config.json, tokenizer, processor configs, chat template; no weights - vLLM's
    load_format="dummy" or a caller's own state dict supplies them) of a very small Qwen2-VL / Qwen2.5-VL whose vision tower
    has the real head_dim (80): depth 2, embed_dim 160, 2 heads, merged into a 256-wide two-layer language model (head_dim 128)."""
    import json
    import os
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
    os.makedirs(path, exist_ok=True)
    tok = tiny_tokenizer()
    vocab = tok.get_vocab()
    # head_dim 128 like the real models (vLLM's FA4 text attention does not take tiny head sizes)
    text = dict(hidden_size=256, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                vocab_size=64, max_position_embeddings=8192, rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]},
                tie_word_embeddings=False)
    ids = dict(image_token_id=vocab["<|image_pad|>"], video_token_id=vocab["<|video_pad|>"], vision_start_token_id=vocab["<|vision_start|>"],
               vision_end_token_id=vocab["<|vision_end|>"], bos_token_id=vocab["<|endoftext|>"], eos_token_id=vocab["<|im_end|>"])
    if arch == "qwen2_vl":
        from transformers.models.qwen2_vl.configuration_qwen2_vl import Qwen2VLConfig
        from transformers.models.qwen2_vl.processing_qwen2_vl import Qwen2VLProcessor as Proc
        cfg = Qwen2VLConfig(text_config=text, vision_config=dict(depth=2, embed_dim=160, hidden_size=256, num_heads=2, mlp_ratio=4), **ids)
        cfg.architectures = ["Qwen2VLForConditionalGeneration"]
    else:
        from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig
        from transformers.models.qwen2_5_vl.processing_qwen2_5_vl import Qwen2_5_VLProcessor as Proc
        cfg = Qwen2_5_VLConfig(text_config=text, vision_config=dict(depth=2, hidden_size=160, intermediate_size=428, num_heads=2,
                                                                     out_hidden_size=256, fullatt_block_indexes=[1], window_size=112), **ids)
        cfg.architectures = ["Qwen2_5_VLForConditionalGeneration"]
    cfg.torch_dtype = "bfloat16"
    cfg.save_pretrained(path)
    chat = ("{% for m in messages %}<|im_start|> {{ m['role'] }} {% if m['content'] is string %}{{ m['content'] }}{% else %}"
            "{% for c in m['content'] %}{% if c['type'] == 'image' or c['type'] == 'image_url' %}<|vision_start|> <|image_pad|> <|vision_end|> "
            "{% elif c['type'] == 'text' %}{{ c['text'] }} {% endif %}{% endfor %}{% endif %}<|im_end|> {% endfor %}"
            "{% if add_generation_prompt %}<|im_start|> assistant {% endif %}")
    proc = Proc(image_processor=Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels), tokenizer=tok,
                video_processor=Qwen2VLVideoProcessor(), chat_template=chat)
    proc.save_pretrained(path)
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        json.dump({"eos_token_id": ids["eos_token_id"], "do_sample": False}, f)
    return path
