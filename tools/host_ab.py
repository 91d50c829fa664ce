"""A/B of host-stage builds (run on the GPU box, quiet cores): python tools/host_ab.py name=path[:ENV=1] ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, ctypes as C
sys.path[:0] = [%r, %r]
import jpeg_util
from zune_jpeg_b200.decoder import Decoder, ZuneJpegOptions
from zune_jpeg_b200._ffi import ZjImage
data = jpeg_util.synth_jpeg(5, 3840, 2160, "420", quality=90)
d = Decoder.new_with_options(ZuneJpegOptions().set_num_threads(1))
img = ZjImage(); ts = []
for _ in range(25):
    t0 = time.perf_counter(); rc = d._lib.zj_decoder_decode_coefficients(d._h, data, len(data), C.byref(img)); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort(); print("%%.2f %%.2f" %% (ts[0], ts[len(ts) // 2]))
''' % (ROOT, os.path.join(ROOT, "tests"))
variants = [a.split("=", 1) for a in sys.argv[1:]]
res = {n: [] for n, _ in variants}
for rep in range(3):
    for name, spec in variants:
        path, *envs = spec.split(":")
        env = dict(os.environ, ZJ_LIB_PATH=os.path.join(ROOT, path))
        for e in envs:
            k, v = e.split("=")
            env[k] = v
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        res[name].append(out.stdout.strip() or out.stderr[-300:])
for n, r in res.items():
    print(n, r)
