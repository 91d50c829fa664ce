"""Phase times of zj_decode_batch_gpu (run on the GPU box): python tools/gpu_entropy_probe.py [c5|c2r] [images]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import jpeg_util  # noqa: E402
from zune_jpeg_b200 import gpu  # noqa: E402
from zune_jpeg_b200.decoder import ZuneJpegOptions, decode_batch  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
w, h = {"c5": (8192, 8192), "c2r": (3840, 2160)}[cfg]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
jpegs = [jpeg_util.synth_jpeg(i, w, h, "420", 90, False, False, 1) for i in range(4)]
jpegs = [jpegs[i % 4] for i in range(n)]
out_bytes = w * h * 3
pinned = gpu.PinnedBuffer(out_bytes * n)
outs = [pinned.array[b * out_bytes:(b + 1) * out_bytes] for b in range(n)]
o = ZuneJpegOptions()
for rep in range(3):
    stats = {}
    if rep == 2:
        os.environ["ZJ_GPU_ENTROPY_TRACE"] = "1"
    t0 = time.perf_counter()
    decode_batch(jpegs, o, threads=0, out=outs, gpu_entropy=True, stats=stats)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {n} x {w}x{h}: {1e3 * dt:.1f} ms, {n * w * h / 1e6 / dt:.0f} MP/s, on GPU: {stats}", flush=True)
os.environ.pop("ZJ_GPU_ENTROPY_TRACE")
t0 = time.perf_counter()
decode_batch(jpegs, o, threads=0, out=outs)
dt = time.perf_counter() - t0
print(f"host stage route: {1e3 * dt:.1f} ms, {n * w * h / 1e6 / dt:.0f} MP/s")
# pixels left on the device, JPEG bytes from pinned memory
import numpy as np  # noqa: E402
pin = gpu.PinnedBuffer(sum(len(j) for j in jpegs))
ins, off = [], 0
for j in jpegs:
    pin.array[off:off + len(j)] = np.frombuffer(j, np.uint8)
    ins.append(pin.array[off:off + len(j)])
    off += len(j)
dev = [gpu.DeviceBuffer(out_bytes) for _ in range(min(n, 64))]
targets = [(dev[i % len(dev)].ptr, out_bytes) for i in range(n)]
for rep in range(3):
    stats = {}
    t0 = time.perf_counter()
    decode_batch(ins, o, threads=0, device_out=targets, stats=stats)
    dt = time.perf_counter() - t0
    print(f"device out rep {rep}: {1e3 * dt:.1f} ms, {n * w * h / 1e6 / dt:.0f} MP/s, on GPU: {stats}", flush=True)
