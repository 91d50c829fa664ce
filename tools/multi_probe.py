#!/usr/bin/env python
"""tools/multi_probe.py [n_dev] -- the in-process multi-device entry point on a multi-GPU box: zj_gpu_reconstruct_multi over
1 .. n_dev devices for (a) a batch of 4K 4:2:0 images (image ranges per device) and (b) ONE 8192x8192 image (strip ranges per
device), pinned host planes in, pinned host pixels out; checks (b) against the single-device result."""
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

lib = _ffi.load()
n_max = int(sys.argv[1]) if len(sys.argv) > 1 else gpu.device_count()
res = {}


def pinned_images(cfg, n_distinct, batch):
    pool = bench.make_pool(cfg, n_distinct, 0)
    out_bytes = gpu.output_size(pool[0][0])
    blocks, arr = [], (ZjImage * batch)()
    for (img, planes, _) in pool:
        offs, tot = [], 0
        for p in planes:
            offs.append(tot); tot += (p.nbytes + 255) & ~255
        blk = gpu.PinnedBuffer(tot)
        for p, o in zip(planes, offs):
            blk.array[o:o + p.nbytes] = p.view(np.uint8)
        blocks.append((blk, offs))
    for b in range(batch):
        img = pool[b % n_distinct][0]
        C.memmove(C.byref(arr[b]), C.byref(img), C.sizeof(ZjImage))
        blk, offs = blocks[b % n_distinct]
        for z in range(img.n_comp):
            arr[b].comp[z].coeff = blk.ptr + offs[z]
    out = gpu.PinnedBuffer(out_bytes * batch)
    optr = (C.c_void_p * batch)(*[out.ptr + b * out_bytes for b in range(batch)])
    olen = (C.c_size_t * batch)(*[out_bytes] * batch)
    return pool, arr, out, optr, olen, out_bytes, blocks


for name, cfg, batch in (("batch_4k", "c2", 64), ("one_8192", "c5", 1)):
    pool, arr, out, optr, olen, out_bytes, keep = pinned_images(cfg, min(batch, 8), batch)
    w, h = bench.CONFIGS[cfg][0], bench.CONFIGS[cfg][1]
    ref = None
    for nd in [n for n in (1, 2, 4, 8) if n <= n_max]:
        devs = (C.c_int * nd)(*range(nd))
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            rc = lib.zj_gpu_reconstruct_multi(devs, nd, arr, batch, optr, olen)
            dt = time.perf_counter() - t0
            if rc:
                raise SystemExit(f"{name} n_dev={nd}: rc {rc} {lib.zj_gpu_last_cuda_error().decode()}")
            best = min(best, dt)
        px = out.array[:out_bytes].copy()
        if ref is None:
            ref = px
        same = bool(np.array_equal(px, ref))
        res[f"{name}_n{nd}"] = {"ms": round(best * 1e3, 2), "MP/s": round(batch * w * h / 1e6 / best), "equal_to_one_device": same}
        print(name, "devices", nd, res[f"{name}_n{nd}"], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"multi_probe_n{n_max}.json"), "w"), indent=1)
