"""Synthetic JPEG generation shared by tests and bench.py (seeded smooth random images)."""
from __future__ import annotations

import io

import numpy as np
from PIL import Image

SUBSAMPLING = {"444": 0, "422": 1, "420": 2}


def smooth_image(seed: int, width: int, height: int, gray: bool = False, noise: int = 4) -> Image.Image:
    """low-resolution uint8 noise, bicubically up-sampled, plus +-noise fine noise"""
    rng = np.random.default_rng(20260000 + seed)
    ch = 1 if gray else 3
    lo = rng.integers(0, 256, size=(height // 32 + 2, width // 32 + 2, ch)).astype(np.uint8)
    if gray:
        im = Image.fromarray(lo[:, :, 0], "L").resize((width, height), Image.BICUBIC)
        a = np.asarray(im).astype(np.int16) + rng.integers(-noise, noise + 1, size=(height, width))
        return Image.fromarray(np.clip(a, 0, 255).astype(np.uint8), "L")
    im = Image.fromarray(lo, "RGB").resize((width, height), Image.BICUBIC)
    a = np.asarray(im).astype(np.int16) + rng.integers(-noise, noise + 1, size=(height, width, 3))
    return Image.fromarray(np.clip(a, 0, 255).astype(np.uint8), "RGB")


def encode(im: Image.Image, quality: int = 90, subsampling: str = "420", progressive: bool = False,
           restart_rows: int = 0) -> bytes:
    bio = io.BytesIO()
    kw = dict(quality=quality, progressive=progressive)
    if im.mode != "L":
        kw["subsampling"] = SUBSAMPLING[subsampling]
    if restart_rows:
        kw["restart_marker_rows"] = restart_rows
    im.save(bio, "JPEG", **kw)
    return bio.getvalue()


def synth_jpeg(seed, width, height, subsampling="420", quality=90, progressive=False, gray=False, restart_rows=0) -> bytes:
    return encode(smooth_image(seed, width, height, gray), quality, subsampling, progressive, restart_rows)
