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
