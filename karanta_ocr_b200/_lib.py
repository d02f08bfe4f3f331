"""ctypes binding of libkocr.so (include/kocr.h). The CUDA library is the only implementation: if it is
missing or cannot run on this machine, calls raise -- there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KOCR_LIB") or os.path.join(HERE, "libkocr.so")  # KOCR_LIB: A/B kernel variants in experiments

OK, ERR_INVALID, ERR_ASPECT, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4, -5
RESIZE_PIL, RESIZE_ATEN = 0, 1
LAYOUT_CHW, LAYOUT_HWC, LAYOUT_GRAY = 0, 1, 2
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
ARCH_QWEN2_VL, ARCH_QWEN2_5_VL = 0, 1
EPI_NONE, EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_BIAS_SWIGLU = range(6)


class KocrImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("height", C.c_int32), ("width", C.c_int32), ("layout", C.c_int32),
                ("reserved", C.c_int32)]


class KocrTowerConfig(C.Structure):
    _fields_ = [("arch", C.c_int32), ("depth", C.c_int32), ("embed_dim", C.c_int32), ("num_heads", C.c_int32),
                ("mlp_hidden", C.c_int32), ("out_hidden", C.c_int32), ("patch_size", C.c_int32),
                ("temporal_patch_size", C.c_int32), ("in_channels", C.c_int32), ("spatial_merge_size", C.c_int32),
                ("window_size", C.c_int32), ("n_fullatt", C.c_int32), ("fullatt_block_indexes", C.c_int32 * 8)]


# name -> (restype, argtypes): every symbol include/kocr.h declares
class KocrPngInfo(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32), ("src_channels", C.c_int32),
                ("bit_depth", C.c_int32), ("color_type", C.c_int32), ("interlace", C.c_int32), ("reserved", C.c_int32),
                ("idat_bytes", C.c_int64), ("raw_bytes", C.c_int64)]


SIGNATURES = {
    "kocr_last_error": (C.c_char_p, []),
    "kocr_version": (C.c_char_p, []),
    "kocr_smart_resize": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "kocr_resample_ksize": (C.c_int, [C.c_int, C.c_int]),
    "kocr_resample_coeffs": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "kocr_normalize_lut": (C.c_int, [C.c_int, C.c_void_p]),
    "kocr_num_patches": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]),
    "kocr_pos_ids": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "kocr_cu_seqlens": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "kocr_window_index": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "kocr_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "kocr_destroy": (None, [C.c_void_p]),
    "kocr_set_reserved_sms": (C.c_int, [C.c_void_p, C.c_int]),
    "kocr_preprocess": (C.c_int, [C.c_void_p, C.POINTER(KocrImage), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "kocr_tower_create": (C.c_int, [C.c_void_p, C.POINTER(KocrTowerConfig), C.POINTER(C.c_void_p)]),
    "kocr_tower_destroy": (None, [C.c_void_p]),
    "kocr_tower_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "kocr_tower_finalize": (C.c_int, [C.c_void_p]),
    "kocr_tower_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int]),
    "kocr_tower_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "kocr_tower_plan_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kocr_png_info": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(KocrPngInfo)]),
    "kocr_png_scratch_bytes": (C.c_int64, [C.POINTER(KocrPngInfo), C.c_int]),
    "kocr_png_decode": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_void_p]),
    "kocr_mrope_position_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "kocr_mrope_position_ids_v2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p]),
    "kocr_scatter_image_embeds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int64, C.c_void_p]),
    "kocr_last_launch_count": (C.c_int64, []),
    "kocr_profile_begin": (C.c_int, [C.c_void_p]),
    "kocr_profile_end": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "kocr_profile_class_name": (C.c_char_p, [C.c_int]),
    "kocr_op_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "kocr_op_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                               C.c_int, C.c_void_p]),
    "kocr_op_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """dlopen libkocr.so and type every entry point. Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python karanta_ocr_b200/build.py` (nvcc, sm_100a). "
                "karanta_ocr_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().kocr_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a C status to the exception the transformers call it stands in for would raise."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc in (ERR_INVALID, ERR_ASPECT):
        raise ValueError(msg)
    raise RuntimeError(msg)


_ctx = {}


def context(device_index: int):
    """One KocrCtx per device per process."""
    if device_index not in _ctx:
        h = C.c_void_p()
        check(load().kocr_create(device_index, C.byref(h)))
        _ctx[device_index] = h
    return _ctx[device_index]
