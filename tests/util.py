"""Shared helpers for the tests: random coefficient planes and zj_image descriptors."""
from __future__ import annotations

import ctypes as C

import numpy as np

from zune_jpeg_b200._ffi import ZjImage

# Annex-K luminance / chrominance tables (natural order), scaled like libjpeg quality 90
_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], np.int32)
_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, np.int32)


def std_qt(chroma: bool, quality: int = 90) -> np.ndarray:
    base = _CHROMA if chroma else _LUMA
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((base * scale + 50) // 100, 1, 255).astype(np.int32)


def geometry(width, height, h, v):
    mcu_x = -(-width // (8 * h))
    mcu_y = -(-height // (8 * v))
    return mcu_x, mcu_y


def random_planes(rng, width, height, n_comp, h, v, density=0.15, dc_only_frac=0.3, amp=40, extreme=False):
    """Whole-image coefficient planes as the progressive driver allocates them
    (reference src/mcu_prog.rs:73-79): comp z has mcu_x*h_z x mcu_y*v_z blocks."""
    mcu_x, mcu_y = geometry(width, height, h, v)
    planes = []
    for z in range(n_comp):
        hz, vz = (h, v) if z == 0 else (1, 1)
        nblk = mcu_x * hz * mcu_y * vz
        if extreme:
            c = rng.integers(-32768, 32768, size=(nblk, 64)).astype(np.int16)
        else:
            c = np.zeros((nblk, 64), np.int16)
            mask = rng.random((nblk, 64)) < density
            # favour low frequencies like real images
            fall = 1.0 / (1 + 0.35 * (np.arange(64) // 8 + np.arange(64) % 8))
            vals = (rng.normal(0, amp, size=(nblk, 64)) * fall).round()
            c[mask] = vals[mask].astype(np.int16)
            c[:, 0] = rng.integers(-1024 // 3, 1024 // 3, size=nblk).astype(np.int16)
        dc_only = rng.random(nblk) < dc_only_frac
        c[dc_only, 1:] = 0
        planes.append(np.ascontiguousarray(c.reshape(-1)))
    return planes


def make_image(width, height, planes, qts, h, v, out_cs, variant, progressive=False, ptrs=None) -> ZjImage:
    """Build a zj_image.  planes are numpy int16 arrays (kept alive by the caller) unless `ptrs` gives raw
    device/host addresses."""
    img = ZjImage()
    img.width, img.height = width, height
    img.n_comp = len(planes)
    img.out_cs = out_cs
    img.variant = variant
    img.flags = 1 if progressive else 0
    mcu_x, _ = geometry(width, height, h, v)
    for z in range(len(planes)):
        c = img.comp[z]
        c.coeff = ptrs[z] if ptrs is not None else planes[z].ctypes.data
        c.n_i16 = planes[z].size
        for i in range(64):
            c.qt[i] = int(qts[z][i])
        c.h_samp, c.v_samp = (h, v) if z == 0 else (1, 1)
        c.width_stride = c.h_samp * mcu_x * 8
    return img
