"""KarantaVisionTower: the Qwen2-VL / Qwen2.5-VL vision tower call surface on sm_100a kernels.

Mirrors `Qwen2VisionTransformerPretrainedModel` (transformers models/qwen2_vl/modeling_qwen2_vl.py:687-795) and
`Qwen2_5_VisionTransformerPretrainedModel` (models/qwen2_5_vl/modeling_qwen2_5_vl.py:345-518): the `visual` module
karanta-ocr's model calls at karanta/training/ocr_training.py:86,670 (through get_image_features :1118-1136) and that
vLLM calls as `self.visual(pixel_values, grid_thw=grid_thw)` (vllm model_executor/models/qwen2_vl.py:1376).

    tower = KarantaVisionTower(hf_model.visual.config)      # or a dict with the same field names
    tower.load_state_dict(hf_model.visual.state_dict())     # HF key names
    emb = tower(pixel_values, grid_thw=image_grid_thw)      # bf16 [sum N / 4, out_hidden]

Inference only (no autograd). The forward is a fixed schedule of hand-written kernels in libkocr.so; there is no
PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _cfg_get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def normalize_config(config) -> dict:
    """Accepts a Qwen2VLVisionConfig / Qwen2_5_VLVisionConfig (or a dict with their field names)."""
    model_type = _cfg_get(config, "model_type", "") or ""
    is25 = "qwen2_5" in model_type or _cfg_get(config, "intermediate_size") is not None or _cfg_get(config, "arch") == "qwen2_5_vl"
    if _cfg_get(config, "arch") == "qwen2_vl":
        is25 = False
    if not is25:
        embed = _cfg_get(config, "embed_dim", 1280)
        mlp_hidden = _cfg_get(config, "mlp_hidden")
        if mlp_hidden is None:
            mlp_hidden = int(embed * _cfg_get(config, "mlp_ratio", 4))
        out_hidden = _cfg_get(config, "out_hidden", None) or _cfg_get(config, "hidden_size", 3584)
        full = []
    else:
        embed = _cfg_get(config, "embed_dim", None) or _cfg_get(config, "hidden_size", 1280)
        mlp_hidden = _cfg_get(config, "mlp_hidden", None) or _cfg_get(config, "intermediate_size", 3420)
        out_hidden = _cfg_get(config, "out_hidden", None) or _cfg_get(config, "out_hidden_size", 3584)
        full = list(_cfg_get(config, "fullatt_block_indexes", [7, 15, 23, 31]))
    return dict(arch="qwen2_5_vl" if is25 else "qwen2_vl", depth=_cfg_get(config, "depth", 32), embed_dim=embed,
                num_heads=_cfg_get(config, "num_heads", 16), mlp_hidden=mlp_hidden, out_hidden=out_hidden,
                patch_size=_cfg_get(config, "patch_size", 14), temporal_patch_size=_cfg_get(config, "temporal_patch_size", 2),
                in_channels=_cfg_get(config, "in_channels", 3) or _cfg_get(config, "in_chans", 3),
                spatial_merge_size=_cfg_get(config, "spatial_merge_size", 2), window_size=_cfg_get(config, "window_size", 112),
                fullatt_block_indexes=full)


_DT = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


class GraphedForward:
    """A captured tower forward (KarantaVisionTower.capture): copy the page's pixel_values into `.pixel_values`, call replay()."""

    def __init__(self, graph, pixel_values, output, grid_thw):
        self.graph, self.pixel_values, self.output, self.grid_thw = graph, pixel_values, output, grid_thw

    def replay(self, pixel_values=None):
        if pixel_values is not None:
            self.pixel_values.copy_(pixel_values.reshape(self.pixel_values.shape), non_blocking=True)
        self.graph.replay()
        return self.output


class KarantaVisionTower(torch.nn.Module):
    def __init__(self, config, device=None, hf_output: bool = False):
        super().__init__()
        # transformers 5.x towers return BaseModelOutputWithPooling (get_image_features reads .pooler_output);
        # 4.53.3 (the reference's pin) and vLLM return the merged tensor. hf_output selects the former.
        self.hf_output = bool(hf_output)
        if not torch.cuda.is_available():
            raise RuntimeError("KarantaVisionTower needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.cfg = normalize_config(config)
        self.config = config
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("KarantaVisionTower runs on CUDA devices only")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        self.spatial_merge_size = self.cfg["spatial_merge_size"]
        self.patch_size = self.cfg["patch_size"]
        self.out_hidden_size = self.cfg["out_hidden"]
        c = _lib.KocrTowerConfig()
        c.arch = _lib.ARCH_QWEN2_5_VL if self.cfg["arch"] == "qwen2_5_vl" else _lib.ARCH_QWEN2_VL
        for k in ("depth", "embed_dim", "num_heads", "mlp_hidden", "out_hidden", "patch_size", "temporal_patch_size",
                  "in_channels", "spatial_merge_size", "window_size"):
            setattr(c, k, int(self.cfg[k]))
        full = self.cfg["fullatt_block_indexes"]
        if len(full) > 8:
            raise ValueError("at most 8 full-attention blocks are supported")
        c.n_fullatt = len(full)
        for i, v in enumerate(full):
            c.fullatt_block_indexes[i] = int(v)
        self._h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kocr_tower_create(_lib.context(dev.index), C.byref(c), C.byref(self._h)))
        self._ws = None
        self.last_launch_count = 0

    def __del__(self):
        try:
            h = self.__dict__.get("_h")
            if h is not None and h.value:
                _lib.load().kocr_tower_destroy(h)
                h.value = None
        except Exception:  # interpreter teardown
            pass

    # ---- nn.Module-compatible surface the call sites touch
    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        return self._device

    def get_dtype(self):
        return torch.bfloat16

    def get_device(self):
        return self._device

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """HF key names (patch_embed.proj.weight, blocks.{i}.attn.qkv.weight, ..., merger.mlp.2.bias).
        rotary_pos_emb.inv_freq is a non-persistent buffer upstream and is ignored."""
        lib = _lib.load()
        unexpected = []
        with torch.cuda.device(self._device):
            for name, t in state_dict.items():
                if name.endswith("inv_freq"):
                    continue
                if t.dtype not in _DT:
                    raise ValueError(f"{name}: unsupported dtype {t.dtype}")
                d = t.detach().to(self._device).contiguous()
                shape = (C.c_int64 * d.ndim)(*d.shape)
                torch.cuda.current_stream(self._device).synchronize()
                rc = lib.kocr_tower_set_weight(self._h, name.encode(), d.data_ptr(), _DT[d.dtype], shape, d.ndim)
                if rc == _lib.ERR_INVALID and "unknown weight" in _lib.last_error():
                    if strict:
                        unexpected.append(name)
                    continue
                _lib.check(rc)
        if strict:
            if unexpected:
                raise RuntimeError(f"Unexpected key(s) in state_dict: {unexpected[:8]}")
            rc = lib.kocr_tower_finalize(self._h)
            if rc:
                raise RuntimeError("Missing key(s) in state_dict: " + _lib.last_error())
        return torch.nn.modules.module._IncompatibleKeys([], unexpected)

    def workspace_bytes(self, grid_thw) -> int:
        g = np.ascontiguousarray(np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3))
        return _lib.check(_lib.load().kocr_tower_workspace_bytes(self._h, g.ctypes.data, g.shape[0]))

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, grid_thw, return_hidden: bool = False, **kwargs):
        """hidden_states: pixel_values [..., 1176] (f32 or bf16; leading dims are flattened like PatchEmbed.forward's
        view(-1, ...)); grid_thw: LongTensor [n, 3] or list. Returns the merged embeddings [sum N / 4, out_hidden] bf16
        (transformers 4.53.3 / vLLM convention; `.pooler_output` of the 5.x return type)."""
        g = grid_thw.detach().cpu().numpy() if isinstance(grid_thw, torch.Tensor) else np.asarray(grid_thw)
        g = np.ascontiguousarray(g.astype(np.int64).reshape(-1, 3))
        dev = self._device
        x = hidden_states
        if x.device != dev:
            x = x.to(dev, non_blocking=True)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.bfloat16)
        patch_dim = self.cfg["in_channels"] * self.cfg["temporal_patch_size"] * self.cfg["patch_size"] ** 2
        x = x.reshape(-1, patch_dim).contiguous()
        S = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
        if x.shape[0] != S:
            raise ValueError(f"pixel_values has {x.shape[0]} patches but grid_thw describes {S}")
        m2 = self.spatial_merge_size ** 2
        with torch.cuda.device(dev):
            need = self.workspace_bytes(g)
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty((S // m2, self.cfg["out_hidden"]), dtype=torch.bfloat16, device=dev)
            # 5.x callers get last_hidden_state filled like upstream's BaseModelOutputWithPooling (one extra D2D copy)
            want_hidden = return_hidden or self.hf_output
            hid = torch.empty((S, self.cfg["embed_dim"]), dtype=torch.bfloat16, device=dev) if want_hidden else None
            rc = _lib.load().kocr_tower_forward(self._h, x.data_ptr(), _DT[x.dtype], g.ctypes.data, g.shape[0], out.data_ptr(),
                                                hid.data_ptr() if hid is not None else None, self._ws.data_ptr(),
                                                self._ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc)
            self.last_launch_count = int(_lib.load().kocr_last_launch_count())
            if not torch.cuda.is_current_stream_capturing():
                x.record_stream(torch.cuda.current_stream(dev))
        if self.hf_output:
            from transformers.modeling_outputs import BaseModelOutputWithPooling
            return BaseModelOutputWithPooling(last_hidden_state=hid, pooler_output=out)
        if return_hidden:
            return out, hid
        return out

    @torch.no_grad()
    def capture(self, grid_thw, dtype=torch.bfloat16):
        """CUDA-graph form of forward() for one grid_thw (a serving loop sees the same page shape again and again): returns
        a GraphedForward whose `.pixel_values` [sum N, 1176] is the static input and whose `replay()` launches the whole
        tower as one graph and returns the static output tensor. One eager forward is run first (plan, rotary table and
        shared-memory opt-ins must exist before the capture); the tables of that plan stay pinned for the tower's lifetime."""
        g = np.ascontiguousarray(np.asarray(grid_thw.cpu() if isinstance(grid_thw, torch.Tensor) else grid_thw, dtype=np.int64).reshape(-1, 3))
        S = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
        patch_dim = self.cfg["in_channels"] * self.cfg["temporal_patch_size"] * self.cfg["patch_size"] ** 2
        dev = self._device
        with torch.cuda.device(dev):
            pv = torch.zeros((S, patch_dim), dtype=dtype, device=dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self.forward(pv, g)  # warm-up outside the capture
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(pv, g)
        return GraphedForward(graph, pv, out, g)

    @classmethod
    def replace_visual(cls, model, device=None):
        """Swap a loaded Qwen2-VL / Qwen2.5-VL model's vision tower for this one (inference): finds `model.visual`
        (transformers 4.53.3, vLLM) or `model.model.visual` (transformers 5.x), builds the tower from its config and
        state dict, and assigns it in place. Returns the new tower."""
        owner, hf5 = None, False
        if hasattr(model, "visual"):
            owner = model
        elif hasattr(model, "model") and hasattr(model.model, "visual"):
            owner, hf5 = model.model, True
        if owner is None:
            raise ValueError("model has no `visual` / `model.visual` module")
        old = owner.visual
        if not hf5:
            try:  # a 5.x Qwen2VLModel used directly also returns the structured output
                import transformers
                hf5 = int(transformers.__version__.split(".")[0]) >= 5 and old.__class__.__module__.startswith("transformers.")
            except Exception:
                hf5 = False
        dev = device if device is not None else next(old.parameters()).device
        tower = cls(old.config, device=dev, hf_output=hf5)
        tower.load_state_dict(old.state_dict())
        owner.visual = tower
        return tower

    def split_per_image(self, embeddings: torch.Tensor, grid_thw):
        """get_image_features' per-image split (modeling_qwen2_vl.py:1132-1134)."""
        g = grid_thw.detach().cpu().numpy() if isinstance(grid_thw, torch.Tensor) else np.asarray(grid_thw)
        sizes = (g.reshape(-1, 3).prod(-1) // self.spatial_merge_size ** 2).tolist()
        return torch.split(embeddings, sizes)
