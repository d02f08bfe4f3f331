"""Seeded synthetic pages: white background 250+-5 with dark text-like
rectangles. Same generator as tests/golden/make_golden.py so golden files and live tests agree."""
import numpy as np
import torch


def synth_page(h, w, seed, gray=False):
    g = torch.Generator().manual_seed(seed)
    img = (250 + torch.randint(-5, 6, (1, h, w), generator=g)).clamp(0, 255).repeat(3, 1, 1)
    n = int(h * w / 2500)
    ys = torch.randint(0, max(h - 12, 1), (n,), generator=g)
    xs = torch.randint(0, max(w - 60, 1), (n,), generator=g)
    ws = torch.randint(4, 60, (n,), generator=g)
    hs = torch.randint(2, 12, (n,), generator=g)
    cs = torch.randint(0, 90, (n, 3), generator=g)
    for i in range(n):
        img[:, ys[i]:ys[i] + hs[i], xs[i]:xs[i] + ws[i]] = cs[i][:, None, None]
    out = img.to(torch.uint8).numpy()
    if gray:
        out = np.repeat(out[:1], 3, axis=0)
    return out
