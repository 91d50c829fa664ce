#!/usr/bin/env python
"""tools/zero_copy_probe.py [batch] -- can the fused kernel read its coefficient planes from, and write its pixels to, PINNED HOST
memory directly (zero-copy over PCIe, no staging copies)?  Times zj_gpu_reconstruct_device with host / device pointers on either
side against the staged zj_gpu_reconstruct on the c2 workload (3840x2160 4:2:0 -> RGB)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import bench
from zune_jpeg_b200 import _ffi, gpu
from zune_jpeg_b200._ffi import ZjImage

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
w, h = bench.CONFIGS[cfg][0], bench.CONFIGS[cfg][1]
pool = bench.make_pool(cfg, 8, 0)
lib = _ffi.load()
out_bytes = gpu.output_size(pool[0][0])
stream = gpu.Stream(0)
# pinned host planes (one block per image) and outputs; device copies of both
host_planes, dev_planes = [], []
for (img, planes, _) in pool:
    hp, dp = [], []
    for p in planes:
        pb = gpu.PinnedBuffer(p.nbytes); pb.array[:] = p.view(np.uint8); hp.append(pb)
        db = gpu.DeviceBuffer(p.nbytes); db.upload(p); dp.append(db)
    host_planes.append(hp); dev_planes.append(dp)
host_out = gpu.PinnedBuffer(out_bytes * batch)
dev_out = gpu.DeviceBuffer(out_bytes * batch)
plane_bytes = sum(p.nbytes for p in pool[0][1])


def images(on_host):
    arr = (ZjImage * batch)()
    for b in range(batch):
        img = pool[b % 8][0]
        C.memmove(C.byref(arr[b]), C.byref(img), C.sizeof(ZjImage))
        for z in range(img.n_comp):
            arr[b].comp[z].coeff = (host_planes if on_host else dev_planes)[b % 8][z].ptr
    return arr


def outs(on_host):
    base = host_out.ptr if on_host else dev_out.ptr
    return (C.c_void_p * batch)(*[base + b * out_bytes for b in range(batch)]), (C.c_size_t * batch)(*[out_bytes] * batch)


res = {}
for name, (ph, oh) in {"dev_in_dev_out": (False, False), "host_in_dev_out": (True, False), "dev_in_host_out": (False, True), "host_in_host_out": (True, True)}.items():
    arr = images(ph)
    optr, olen = outs(oh)
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        rc = lib.zj_gpu_reconstruct_device(0, stream.ptr, arr, batch, optr, olen)
        dt = time.perf_counter() - t0
        if rc:
            print(name, "rc", rc, lib.zj_gpu_last_cuda_error().decode()); break
        best = min(best, dt)
    res[name] = {"ms": round(best * 1e3, 2), "in_GBps": round(plane_bytes * batch / best / 1e9, 1), "out_GBps": round(out_bytes * batch / best / 1e9, 1), "MP/s": round(batch * w * h / 1e6 / best)}
    print(name, res[name], flush=True)
arr = images(True)
optr, olen = outs(True)
best = 1e9
for it in range(4):
    t0 = time.perf_counter()
    rc = lib.zj_gpu_reconstruct(0, stream.ptr, arr, batch, optr, olen)
    best = min(best, time.perf_counter() - t0)
res["staged"] = {"ms": round(best * 1e3, 2), "in_GBps": round(plane_bytes * batch / best / 1e9, 1), "out_GBps": round(out_bytes * batch / best / 1e9, 1), "MP/s": round(batch * w * h / 1e6 / best)}
print("staged", res["staged"])
# the zero-copy result must be the staged one
a = host_out.array[:out_bytes].copy()
lib.zj_gpu_reconstruct_device(0, stream.ptr, images(True), batch, *outs(True))
print("equal:", bool(np.array_equal(a, host_out.array[:out_bytes])))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"zero_copy_{cfg}_{batch}.json"), "w"), indent=1)
