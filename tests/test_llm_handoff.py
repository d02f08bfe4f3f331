"""LLM hand-off: M-RoPE position ids are host planning (CPU tests, bit-exact against the
transformers golden and the oracle); the masked scatter is a CUDA kernel (gpu test, bit-exact: it only moves bf16 rows)."""
import os

import numpy as np
import pytest
import torch

from karanta_ocr_b200 import get_rope_index, scatter_image_features
from oracle import vision_oracle as vo

G = os.path.join(os.path.dirname(__file__), "golden")


def test_position_ids_equal_transformers_golden():
    z = np.load(os.path.join(G, "g6_llm_handoff.npz"))
    img = int(z["image_token_id"])
    names = sorted({k.split(".")[0] for k in z.files if "." in k})
    assert len(names) == 4
    for n in names:
        pos, delta = get_rope_index(z[f"{n}.input_ids"], z[f"{n}.grid"], z[f"{n}.attention_mask"], image_token_id=img)
        assert pos.dtype == torch.int64 and tuple(pos.shape) == z[f"{n}.position_ids"].shape
        assert np.array_equal(pos.numpy(), z[f"{n}.position_ids"]), n
        assert np.array_equal(delta.numpy(), z[f"{n}.deltas"]), n
        opos, odelta = vo.mrope_position_ids(z[f"{n}.input_ids"], z[f"{n}.grid"], z[f"{n}.attention_mask"], img)
        assert np.array_equal(opos, z[f"{n}.position_ids"]) and np.array_equal(odelta, z[f"{n}.deltas"]), n


def test_position_ids_random_prompts_vs_oracle():
    rng = np.random.default_rng(9)
    IMG = 151655
    for _ in range(25):
        B = int(rng.integers(1, 4))
        rows, grids = [], []
        for _b in range(B):
            toks = list(rng.integers(1, 1000, int(rng.integers(0, 6))))
            for _i in range(int(rng.integers(0, 3))):
                h, w = int(rng.integers(1, 12)) * 2, int(rng.integers(1, 12)) * 2
                grids.append([1, h, w])
                toks += [IMG] * (h * w // 4) + list(rng.integers(1, 1000, int(rng.integers(1, 5))))
            rows.append(toks or [7])
        L = max(len(r) for r in rows)
        ids = np.zeros((B, L), dtype=np.int64)
        mask = np.zeros((B, L), dtype=np.int64)
        for i, r in enumerate(rows):
            ids[i, L - len(r):] = r
            mask[i, L - len(r):] = 1
        grid = np.asarray(grids, dtype=np.int64).reshape(-1, 3)
        pos, delta = get_rope_index(ids, grid if len(grid) else None, mask, image_token_id=IMG)
        opos, odelta = vo.mrope_position_ids(ids, grid, mask, IMG)
        assert np.array_equal(pos.numpy(), opos) and np.array_equal(delta.numpy(), odelta)
    # no padding mask
    ids = np.asarray([[5, 6, IMG, IMG, IMG, IMG, 9]], dtype=np.int64)
    pos, delta = get_rope_index(ids, [[1, 4, 4]], None, image_token_id=IMG)
    assert pos[:, 0].tolist() == [[0, 1, 2, 2, 2, 2, 4], [0, 1, 2, 2, 3, 3, 4], [0, 1, 2, 3, 2, 3, 4]] and delta.tolist() == [[-2]]


def test_position_ids_errors():
    IMG = 151655
    with pytest.raises(ValueError):
        get_rope_index(np.asarray([[1, IMG, IMG, 2]]), [[1, 4, 4]], None, image_token_id=IMG)      # 2 placeholders, grid needs 4
    with pytest.raises(ValueError):
        get_rope_index(np.asarray([[1, IMG, 2, IMG]]), [[1, 2, 2]], None, image_token_id=IMG)      # two runs, one grid row


@pytest.mark.gpu
def test_scatter_equals_masked_scatter():
    IMG = 151655
    g = torch.Generator().manual_seed(0)
    B, L, H = 3, 700, 3584
    ids = torch.randint(1, 1000, (B, L), generator=g)
    ids[0, 10:10 + 414] = IMG
    ids[1, 0:64] = IMG
    ids[2, 300:699] = IMG
    n = int((ids == IMG).sum())
    emb = torch.randn(B, L, H, generator=g).to(torch.bfloat16)
    img = torch.randn(n, H, generator=g).to(torch.bfloat16)
    ref = vo.scatter_image_features(emb.clone(), ids, img, IMG)
    out = scatter_image_features(emb.cuda(), ids, img.cuda(), IMG)
    assert torch.equal(out.cpu(), ref)
    with pytest.raises(ValueError, match="Image features and image tokens do not match"):
        scatter_image_features(emb.cuda(), ids, img[:-1].cuda(), IMG)


@pytest.mark.gpu
def test_scatter_full_page_batch():
    """C2-sized hand-off: 8 prompts x 1656 image tokens of width 3584; every placeholder row replaced, text rows untouched."""
    IMG = 151655
    B, L, H, T = 8, 1700, 3584, 1656
    ids = torch.full((B, L), 11, dtype=torch.int64)
    ids[:, 20:20 + T] = IMG
    emb = torch.zeros(B, L, H, dtype=torch.bfloat16, device="cuda")
    img = torch.arange(B * T, dtype=torch.float32, device="cuda").remainder(251).to(torch.bfloat16).unsqueeze(1).expand(-1, H).contiguous()
    out = scatter_image_features(emb, ids, img, IMG)
    assert torch.equal(out[:, 20:20 + T].reshape(B * T, H), img)
    assert (out[:, :20] == 0).all() and (out[:, 20 + T:] == 0).all()
