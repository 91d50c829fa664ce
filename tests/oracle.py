"""ctypes wrapper around oracle/libzj_oracle*.so (TEST INFRASTRUCTURE -- the checker, never the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from zune_jpeg_b200._ffi import ZjImage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ERR_PANIC = -5


def _has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return " avx2" in txt and " sse4_1" in txt
    except OSError:
        return False


def build() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


_libs = {}


def lib(emulated: bool | None = None) -> C.CDLL:
    """emulated=None: the real-SIMD build when the host has AVX2, else the emulation."""
    if emulated is None:
        emulated = not _has_avx2()
    name = "libzj_oracle_emul.so" if emulated else "libzj_oracle.so"
    if name not in _libs:
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        sz, vp = C.c_size_t, C.c_void_p
        for fn in ("zjo_idct_scalar", "zjo_idct_avx2"):
            getattr(L, fn).argtypes = [vp, sz, vp, sz, sz, sz, vp]
        for fn in ("zjo_upsample_horizontal_scalar", "zjo_upsample_horizontal_sse", "zjo_upsample_vertical",
                   "zjo_upsample_hv_scalar", "zjo_upsample_hv_simd"):
            getattr(L, fn).argtypes = [vp, sz, vp, sz]
        L.zjo_reconstruct_image.argtypes = [C.POINTER(ZjImage), vp, sz, C.c_int]
        L.zjo_output_size.argtypes = [C.POINTER(ZjImage)]
        L.zjo_output_size.restype = sz
        _libs[name] = L
    return _libs[name]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def idct(coef: np.ndarray, qt: np.ndarray, stride: int, samp: int, v_samp: int, variant: int, emulated=None) -> np.ndarray:
    coef = np.ascontiguousarray(coef, np.int16).reshape(-1)
    qt = np.ascontiguousarray(qt, np.int32)
    out = np.zeros(coef.size, np.int16)
    fn = lib(emulated).zjo_idct_avx2 if variant == 0 else lib(emulated).zjo_idct_scalar
    rc = fn(_ptr(coef), coef.size, _ptr(qt), stride, samp, v_samp, _ptr(out))
    if rc:
        raise RuntimeError(f"oracle idct rc={rc}")
    return out


def upsample(name: str, x: np.ndarray, out_len: int, emulated=None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.int16)
    out = np.zeros(out_len, np.int16)
    rc = getattr(lib(emulated), "zjo_upsample_" + name)(_ptr(x), x.size, _ptr(out), out_len)
    if rc:
        raise RuntimeError(f"oracle upsample_{name} rc={rc}")
    return out


def reconstruct(img: ZjImage, threads: int = 1, emulated=None) -> np.ndarray:
    """zjo_reconstruct_image -> uint8 array; raises RuntimeError(rc) on error."""
    L = lib(emulated)
    n = L.zjo_output_size(C.byref(img))
    if n == 0:
        # let the library say why
        out = np.zeros(1, np.uint8)
        rc = L.zjo_reconstruct_image(C.byref(img), _ptr(out), 0, threads)
        raise RuntimeError(f"oracle rc={rc}")
    out = np.empty(n, np.uint8)
    rc = L.zjo_reconstruct_image(C.byref(img), _ptr(out), n, threads)
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return out
