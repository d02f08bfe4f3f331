"""karanta_ocr_b200: B200-native (sm_100a) implementation of karanta-ocr's page-image hot path.

page image -> smart_resize -> resize/normalise/patchify -> pixel_values + image_grid_thw -> vision tower -> embeddings,
behind the transformers call surface the reference uses. See DESIGN.md and INTEGRATION.md.
"""
from . import _lib, presets  # noqa: F401
from .image_processor import KarantaImageProcessor, smart_resize  # noqa: F401
from .llm_handoff import get_rope_index, scatter_image_features  # noqa: F401
from .png_decode import PngError, decode_png_batch, png_info  # noqa: F401
from .pipeline import PageEncoder, gather_pages, page_cost, shard_pages  # noqa: F401
from .vision_tower import KarantaVisionTower, normalize_config  # noqa: F401

__all__ = ["KarantaImageProcessor", "KarantaVisionTower", "PageEncoder", "smart_resize", "shard_pages", "gather_pages", "page_cost",
           "normalize_config", "get_rope_index", "scatter_image_features", "decode_png_batch", "png_info", "PngError"]
