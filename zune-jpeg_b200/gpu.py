"""Thin Python view of the GPU half of the C ABI (include/zune_jpeg_b200.h).  Every call goes through
libzune_jpeg_b200.so; nothing here computes pixels and there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import ZjImage


class ZjError(RuntimeError):
    def __init__(self, status: int, where: str = ""):
        lib = _ffi.load()
        msg = lib.zj_gpu_strerror(status).decode()
        if status == _ffi.ERR_CUDA or status == _ffi.ERR_OOM:
            msg += " -- " + lib.zj_gpu_last_cuda_error().decode()
        super().__init__(f"{where}: [{status}] {msg}" if where else f"[{status}] {msg}")
        self.status = status


def _check(status: int, where: str = "") -> None:
    if status != 0:
        raise ZjError(status, where)


def device_count() -> int:
    return _ffi.load().zj_gpu_device_count()


def launch_count() -> int:
    return int(_ffi.load().zj_gpu_launch_count())


def output_size(img: ZjImage) -> int:
    return int(_ffi.load().zj_output_size(C.byref(img)))


def validate(img: ZjImage) -> int:
    return int(_ffi.load().zj_validate_image(C.byref(img)))


def _img_array(images):
    arr = (ZjImage * len(images))()
    for i, im in enumerate(images):
        C.memmove(C.byref(arr[i]), C.byref(im), C.sizeof(ZjImage))
    return arr


def reconstruct(images, device: int = 0, stream=None):
    """zj_gpu_reconstruct: host coefficient planes in, host pixels out (numpy uint8 arrays)."""
    lib = _ffi.load()
    n = len(images)
    arr = _img_array(images)
    outs = []
    ptrs = (C.c_void_p * n)()
    lens = (C.c_size_t * n)()
    for i, im in enumerate(images):
        sz = output_size(im)
        if sz == 0:
            _check(validate(im) or _ffi.ERR_INVALID_ARG, "zj_validate_image")
        o = np.empty(sz, np.uint8)
        outs.append(o)
        ptrs[i] = o.ctypes.data
        lens[i] = sz
    _check(lib.zj_gpu_reconstruct(device, stream, arr, n, ptrs, lens), "zj_gpu_reconstruct")
    return outs


class DeviceBuffer:
    """cudaMalloc'd bytes owned through the C ABI."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        _check(_ffi.load().zj_gpu_device_alloc(device, self.nbytes, C.byref(p)), "zj_gpu_device_alloc")
        self.ptr = p.value

    def upload(self, host: np.ndarray, stream=None, offset: int = 0):
        host = np.ascontiguousarray(host)
        _check(_ffi.load().zj_gpu_memcpy_h2d(self.device, stream, self.ptr + offset, host.ctypes.data, host.nbytes), "h2d")

    def download(self, nbytes: int | None = None, stream=None, offset: int = 0) -> np.ndarray:
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes, np.uint8)
        lib = _ffi.load()
        _check(lib.zj_gpu_memcpy_d2h(self.device, stream, out.ctypes.data, self.ptr + offset, nbytes), "d2h")
        _check(lib.zj_gpu_stream_synchronize(self.device, stream), "sync")
        return out

    def memset(self, value: int, stream=None):
        _check(_ffi.load().zj_gpu_memset(self.device, stream, self.ptr, value, self.nbytes), "memset")

    def free(self):
        if self.ptr:
            _ffi.load().zj_gpu_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """cudaHostAlloc'd bytes exposed as a numpy array."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _check(_ffi.load().zj_gpu_pinned_alloc(self.nbytes, C.byref(p)), "zj_gpu_pinned_alloc")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            _ffi.load().zj_gpu_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Batch:
    """zj_batch_*: a reusable launch plan over device-resident planes and outputs."""

    def __init__(self, images, out_ptrs, out_lens, device: int = 0):
        lib = _ffi.load()
        n = len(images)
        self.device = device
        arr = _img_array(images)
        ptrs = (C.c_void_p * n)(*out_ptrs)
        lens = (C.c_size_t * n)(*out_lens)
        h = C.c_void_p()
        _check(lib.zj_batch_create(device, arr, n, ptrs, lens, C.byref(h)), "zj_batch_create")
        self.handle = h

    def run(self, stream=None):
        _check(_ffi.load().zj_batch_run(self.handle, stream), "zj_batch_run")

    @property
    def launches(self) -> int:
        return _ffi.load().zj_batch_launches(self.handle)

    @property
    def algorithmic_bytes(self) -> int:
        return int(_ffi.load().zj_batch_algorithmic_bytes(self.handle))

    def destroy(self):
        if self.handle:
            _ffi.load().zj_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def synchronize(device: int = 0, stream=None):
    _check(_ffi.load().zj_gpu_stream_synchronize(device, stream), "sync")


class Stream:
    def __init__(self, device: int = 0):
        self.device = device
        p = C.c_void_p()
        _check(_ffi.load().zj_gpu_stream_create(device, C.byref(p)), "stream_create")
        self.ptr = p.value

    def synchronize(self):
        synchronize(self.device, self.ptr)

    def destroy(self):
        if self.ptr:
            _ffi.load().zj_gpu_stream_destroy(self.device, self.ptr)
            self.ptr = None


class Event:
    def __init__(self, device: int = 0):
        self.device = device
        p = C.c_void_p()
        _check(_ffi.load().zj_gpu_event_create(device, C.byref(p)), "event_create")
        self.ptr = p.value

    def record(self, stream=None):
        _check(_ffi.load().zj_gpu_event_record(self.device, self.ptr, stream), "event_record")

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        _check(_ffi.load().zj_gpu_event_elapsed_ms(self.device, self.ptr, stop.ptr, C.byref(ms)), "event_elapsed")
        return float(ms.value)
