"""Where the whole-decode throughput goes (run on the GPU box): the host stage alone (headers + Huffman -> planes) on T
threads, against zj_decode_batch (the same plus upload, kernels, download) on T threads.
    python tools/decode_scaling.py [config] [images]"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import jpeg_util  # noqa: E402
from zune_jpeg_b200 import _ffi, gpu  # noqa: E402
from zune_jpeg_b200._ffi import ZjImage  # noqa: E402
from zune_jpeg_b200.decoder import Decoder, ZuneJpegOptions, decode_batch  # noqa: E402

w, h, sub, rst = {"c2": (3840, 2160, "420", 0), "c5": (8192, 8192, "420", 1)}[sys.argv[1] if len(sys.argv) > 1 else "c2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
jpegs = [jpeg_util.synth_jpeg(i, w, h, sub, 90, False, False, rst) for i in range(8)]
jpegs = [jpegs[i % 8] for i in range(n)]
lib = _ffi.load()
opts = ZuneJpegOptions().set_num_threads(1)
mp = w * h / 1e6
for T in (1, 2, 4, 8, 16, os.cpu_count() or 16):
    tls = threading.local()

    def host_only(j):
        if not hasattr(tls, "d"):
            tls.d = Decoder.new_with_options(opts)
        img = ZjImage()
        rc = lib.zj_decoder_decode_coefficients(tls.d._h, j, len(j), C.byref(img))
        assert rc == 0
    m = min(n, 8 * T)
    with ThreadPoolExecutor(max_workers=T) as ex:
        list(ex.map(host_only, jpegs[:T]))
        t0 = time.perf_counter()
        list(ex.map(host_only, jpegs[:m]))
        dt_h = time.perf_counter() - t0
    out_bytes = w * h * 3
    pinned = gpu.PinnedBuffer(out_bytes * m)
    outs = [pinned.array[b * out_bytes:(b + 1) * out_bytes] for b in range(m)]
    o = ZuneJpegOptions()
    decode_batch(jpegs[:min(m, 3 * T)], o, threads=T, out=outs[:min(m, 3 * T)])   # warm-up: every worker allocates its two sets of pinned planes
    t0 = time.perf_counter()
    decode_batch(jpegs[:m], o, threads=T, out=outs)
    dt_b = time.perf_counter() - t0
    pinned.free()
    print(f"T={T:3d} images={m:4d} host stage only {m * mp / dt_h:8.1f} MP/s ({1e3 * dt_h * T / m:6.2f} ms per image and thread) | "
          f"zj_decode_batch {m * mp / dt_b:8.1f} MP/s ({1e3 * dt_b * T / m:6.2f} ms)", flush=True)
