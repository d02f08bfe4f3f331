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


def test_g6_golden_is_reproducible_from_transformers():
    """tests/golden/make_golden.py g6 (the prompts are written out in the script) run live against the installed
    transformers: the committed position ids / deltas are what Qwen2VLModel.get_rope_index returns for these prompts."""
    pytest.importorskip("transformers")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(G, "g6_llm_handoff.npz"))
    assert sorted(mg.G6_PROMPTS) == sorted({k.split(".")[0] for k in z.files if "." in k})
    for name in mg.G6_PROMPTS:
        pos, delta = mg._hf_rope_index(z[f"{name}.input_ids"], z[f"{name}.grid"], z[f"{name}.attention_mask"], mg.IMG_TOKEN)
        assert np.array_equal(pos, z[f"{name}.position_ids"]) and np.array_equal(delta, z[f"{name}.deltas"]), name


def test_position_ids_transformers_4_5x_semantics():
    """The line the reference pins (transformers 4.53.3) differs from 5.x only at padding and at image location
    (ADVICE round 1): padded positions hold 1, delta is taken against the padded length, adjacent images are split by
    grid size. Unpadded single-image prompts must agree between the two modes."""
    IMG, VS = 151655, 151652
    z = np.load(os.path.join(G, "g6_llm_handoff.npz"))
    # unpadded, one image, no <|vision_start|> in the golden prompts -> locate by runs: identical to 5.x
    for n in ("one_page", "image_first"):
        p5, d5 = get_rope_index(z[f"{n}.input_ids"], z[f"{n}.grid"], z[f"{n}.attention_mask"], image_token_id=IMG)
        p4, d4 = get_rope_index(z[f"{n}.input_ids"], z[f"{n}.grid"], z[f"{n}.attention_mask"], image_token_id=IMG, semantics="4.5x",
                                vision_start_token_id=None)
        assert torch.equal(p4, p5) and torch.equal(d4, d5), n
    # left-padded batch: pads are 1 and delta shrinks by the pad count
    ids, mask, grid = z["batch_padded.input_ids"], z["batch_padded.attention_mask"], z["batch_padded.grid"]
    p5, d5 = get_rope_index(ids, grid, mask, image_token_id=IMG)
    p4, d4 = get_rope_index(ids, grid, mask, image_token_id=IMG, semantics="4.5x", vision_start_token_id=None)
    m = torch.from_numpy(mask).bool()
    assert torch.equal(p4[:, m], p5[:, m]) and (p4[:, ~m] == 1).all() and (p5[:, ~m] == 0).all()
    pads = torch.from_numpy((mask == 0).sum(-1, keepdims=True))
    assert torch.equal(d4, d5 - pads)
    # <|vision_start|> locates the images; two images back to back are split by their grids (5.x refuses the merged run)
    row = [7, VS] + [IMG] * 4 + [IMG] * 6 + [9]
    with pytest.raises(ValueError):
        get_rope_index(np.asarray([row]), [[1, 4, 4], [1, 4, 6]], None, image_token_id=IMG)
    row = [7, VS] + [IMG] * 4 + [VS] + [IMG] * 6 + [9]
    p4, d4 = get_rope_index(np.asarray([row]), [[1, 4, 4], [1, 4, 6]], None, image_token_id=IMG, semantics="4.5x")
    p5, d5 = get_rope_index(np.asarray([row]), [[1, 4, 4], [1, 4, 6]], None, image_token_id=IMG)
    assert torch.equal(p4, p5) and torch.equal(d4, d5)
    # an image token that no <|vision_start|> announces is text in 4.5x
    p4, _ = get_rope_index(np.asarray([[5, IMG, 6]]), None, None, image_token_id=IMG, semantics="4.5x")
    assert p4[:, 0].tolist() == [[0, 1, 2]] * 3
