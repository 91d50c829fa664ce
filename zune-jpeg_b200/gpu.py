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


def reconstruct_multi(images, devices):
    """zj_gpu_reconstruct_multi: one batch over several devices of one box (image ranges, or strip ranges when there are fewer
    images than devices); host planes in, host pixels out."""
    lib = _ffi.load()
    n = len(images)
    arr = _img_array(images)
    outs, ptrs, lens = [], (C.c_void_p * n)(), (C.c_size_t * n)()
    for i, im in enumerate(images):
        sz = output_size(im)
        if sz == 0:
            _check(validate(im) or _ffi.ERR_INVALID_ARG, "zj_validate_image")
        o = np.empty(sz, np.uint8)
        outs.append(o)
        ptrs[i], lens[i] = o.ctypes.data, sz
    devs = (C.c_int * len(devices))(*devices)
    _check(lib.zj_gpu_reconstruct_multi(devs, len(devices), arr, n, ptrs, lens), "zj_gpu_reconstruct_multi")
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


# ------------------------------------------------------------------ device-side consumers
_NP_DTYPES = {_ffi.DTYPE_U8: np.uint8, _ffi.DTYPE_F16: np.float16, _ffi.DTYPE_F32: np.float32}
_TYPESTR = {_ffi.DTYPE_U8: "|u1", _ffi.DTYPE_F16: "<f2", _ffi.DTYPE_F32: "<f4"}


class OutputDesc:
    """zj_output_desc: layout 'HWC' | 'CHW', dtype 'u8' | 'f16' | 'f32', half = 2x2 box down-scale, channels 0 (all) | 3,
    per-channel mean / inv_std in u8 units.  Float values are (float(u8) - mean) * inv_std of the reference's exact bytes."""

    def __init__(self, layout: str = "HWC", dtype: str = "u8", half: bool = False, channels: int = 0,
                 mean=(0.0, 0.0, 0.0, 0.0), inv_std=(1.0, 1.0, 1.0, 1.0)):
        d = _ffi.ZjOutputDesc()
        d.layout = {"HWC": _ffi.LAYOUT_HWC, "CHW": _ffi.LAYOUT_CHW}[layout]
        d.dtype = {"u8": _ffi.DTYPE_U8, "f16": _ffi.DTYPE_F16, "f32": _ffi.DTYPE_F32}[dtype]
        d.scale_log2 = 1 if half else 0
        d.channels = channels
        mean, inv_std = list(mean) + [0.0] * 4, list(inv_std) + [1.0] * 4
        for c in range(4):
            d.mean[c], d.inv_std[c] = float(mean[c]), float(inv_std[c])
        self.c = d

    @property
    def chw(self) -> bool:
        return self.c.layout == _ffi.LAYOUT_CHW

    @property
    def np_dtype(self):
        return _NP_DTYPES[self.c.dtype]

    def shape_of(self, img: ZjImage):
        w, h, ch = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(_ffi.load().zj_consumer_output_shape(C.byref(img), C.byref(self.c), C.byref(w), C.byref(h), C.byref(ch)), "zj_consumer_output_shape")
        return (ch.value, h.value, w.value) if self.chw else (h.value, w.value, ch.value)

    def expected(self, u8: np.ndarray, width: int, height: int, nc: int) -> np.ndarray:
        """The specification, applied to the reference's bytes with numpy (tests: `u8` comes from the oracle)."""
        a = np.asarray(u8, np.uint8).reshape(height, width, nc)
        oc = 3 if (self.c.channels == 3 and nc == 4) else nc
        a = a[:, :, :oc]
        mean = np.array(list(self.c.mean)[:oc], np.float32)
        inv = np.array(list(self.c.inv_std)[:oc], np.float32)
        if self.c.scale_log2:
            h2, w2 = height // 2, width // 2
            s = a[: 2 * h2, : 2 * w2].astype(np.uint32).reshape(h2, 2, w2, 2, oc).sum(axis=(1, 3))
            r = ((s + 2) >> 2).astype(np.uint8) if self.c.dtype == _ffi.DTYPE_U8 else ((s.astype(np.float32) * np.float32(0.25) - mean) * inv)
        else:
            r = a if self.c.dtype == _ffi.DTYPE_U8 else ((a.astype(np.float32) - mean) * inv)
        r = r.astype(self.np_dtype)
        return np.ascontiguousarray(r.transpose(2, 0, 1)) if self.chw else np.ascontiguousarray(r)


class DeviceArray:
    """A typed view of device memory produced by the consumer kernels.  Exposes ``__cuda_array_interface__`` (version 3), so
    ``torch.as_tensor(a, device='cuda')``, ``cupy.asarray(a)`` and numba read it in place; ``__dlpack__`` goes through torch."""

    def __init__(self, buf: DeviceBuffer, shape, dtype_code: int):
        self.buf, self.shape, self.dtype_code = buf, tuple(int(x) for x in shape), dtype_code

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": _TYPESTR[self.dtype_code], "data": (int(self.buf.ptr or 0), False), "version": 3, "strides": None}

    def to_torch(self):
        import torch
        return torch.as_tensor(self, device=f"cuda:{self.buf.device}")

    def __dlpack__(self, stream=None):
        return self.to_torch().__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return (2, self.buf.device)   # kDLCUDA

    def download(self) -> np.ndarray:
        n = int(np.prod(self.shape)) * np.dtype(_NP_DTYPES[self.dtype_code]).itemsize
        return self.buf.download(n).view(_NP_DTYPES[self.dtype_code]).reshape(self.shape)


def reconstruct_device_ex(images, desc: OutputDesc, device: int = 0, stream=None):
    """zj_gpu_reconstruct_device_ex: `images` carry DEVICE coefficient planes; returns one DeviceArray per image."""
    lib = _ffi.load()
    n = len(images)
    arr = _img_array(images)
    outs, ptrs, lens = [], (C.c_void_p * n)(), (C.c_size_t * n)()
    for i, im in enumerate(images):
        sz = int(lib.zj_consumer_output_size(C.byref(im), C.byref(desc.c)))
        if sz == 0 and output_size(im) == 0:
            _check(validate(im) or _ffi.ERR_INVALID_ARG, "zj_validate_image")
        b = DeviceBuffer(max(sz, 1), device)
        outs.append(DeviceArray(b, desc.shape_of(im), desc.c.dtype))
        ptrs[i], lens[i] = b.ptr, sz
    _check(lib.zj_gpu_reconstruct_device_ex(device, stream, arr, n, C.byref(desc.c), ptrs, lens), "zj_gpu_reconstruct_device_ex")
    return outs


def convert_device(src: DeviceBuffer, width: int, height: int, nc: int, desc: OutputDesc, device: int = 0, stream=None) -> DeviceArray:
    """zj_gpu_convert_device: the consumer alone over interleaved u8 pixels already in device memory."""
    lib = _ffi.load()
    ow, oh = width >> desc.c.scale_log2, height >> desc.c.scale_log2
    oc = 3 if (desc.c.channels == 3 and nc == 4) else nc
    sz = ow * oh * oc * np.dtype(desc.np_dtype).itemsize
    b = DeviceBuffer(max(sz, 1), device)
    _check(lib.zj_gpu_convert_device(device, stream, src.ptr, width, height, nc, C.byref(desc.c), b.ptr, sz), "zj_gpu_convert_device")
    synchronize(device, stream)
    return DeviceArray(b, (oc, oh, ow) if desc.chw else (oh, ow, oc), desc.c.dtype)
