"""A spec-correct (ITU T.81 / libjpeg-style) reconstruction of coefficient planes in numpy: float IDCT, "fancy" triangle
up-sampling with replicated edges, JFIF YCbCr -> RGB.  TEST INFRASTRUCTURE: it has none of the reference's pixel-path quirks
(Q1 dropped rows, Q4 flat-strip up-samplers, Q5 row tails), so comparing ITS output with libjpeg's
checks the HOST STAGE's coefficient planes alone (tests/test_quirks.py)."""
from __future__ import annotations

import numpy as np
from scipy.fft import idctn


def _tri(a: np.ndarray, axis: int) -> np.ndarray:
    a = np.moveaxis(a, axis, -1)
    p = np.concatenate([a[..., :1], a[..., :-1]], -1)
    n = np.concatenate([a[..., 1:], a[..., -1:]], -1)
    o = np.empty(a.shape[:-1] + (2 * a.shape[-1],), np.float32)
    o[..., 0::2] = (3 * a + p) / 4
    o[..., 1::2] = (3 * a + n) / 4
    return np.moveaxis(o, -1, axis)


def sane_pixels(img, planes, width: int, height: int):
    """(ZjImage, planes) of Decoder.decode_coefficients -> (rows x width x 3 int16 RGB, rows); rows < height when the
    baseline driver never decoded the last MCU row (Q1: the planes end there)."""
    comps = []
    for z in range(img.n_comp):
        c = img.comp[z]
        bpr = c.width_stride // 8
        p = planes[z].astype(np.float32).reshape(-1, bpr, 8, 8) * np.array(c.qt, np.float32).reshape(8, 8)
        s = idctn(p, axes=(2, 3), norm="ortho") + 128.0
        comps.append(np.clip(s.transpose(0, 2, 1, 3).reshape(-1, bpr * 8), 0, 255))
    y = comps[0]
    if img.n_comp == 1:
        rows = min(y.shape[0], height)
        g = np.clip(np.rint(y[:rows, :width]), 0, 255).astype(np.int16)
        return np.stack([g, g, g], -1), rows
    hs, vs = img.comp[0].h_samp, img.comp[0].v_samp
    up = []
    for c in comps[1:]:
        if hs == 2:
            c = _tri(c, 1)
        if vs == 2:
            c = _tri(c, 0)
        up.append(c)
    rows = min(y.shape[0], up[0].shape[0], height)
    y, cb, cr = y[:rows, :width], up[0][:rows, :width] - 128, up[1][:rows, :width] - 128
    rgb = np.stack([y + 1.402 * cr, y - 0.344136 * cb - 0.714136 * cr, y + 1.772 * cb], -1)
    return np.clip(np.rint(rgb), 0, 255).astype(np.int16), rows
