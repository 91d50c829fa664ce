"""One-off differential fuzz (run on the GPU box): random sizes / sub-samplings / qualities / restart spacings, a share of
them damaged, through zj_decode_batch_gpu and through the host stage; pixels and per-image errors must agree.
    python tools/fuzz_gpu_entropy.py [cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import jpeg_util  # noqa: E402
from zune_jpeg_b200.decoder import ColorSpace, DecodeErrors, ZuneJpegOptions, decode_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4321)
jpegs = []
for it in range(n):
    w, h = int(rng.integers(64, 1500)), int(rng.integers(64, 1200))
    sub, gray = ["420", "422", "444"][it % 3], it % 7 == 0
    q = int(rng.choice([40, 75, 85, 90, 95, 98, 100]))
    d = bytearray(jpeg_util.synth_jpeg(it, w, h, sub, q, False, gray, int(rng.integers(1, 4))))
    sos = bytes(d).index(b"\xff\xda")
    kind = it % 5
    if kind == 1:
        for p in rng.integers(sos + 14, len(d) - 2, size=3):
            d[p] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2:
        i = bytes(d).index(b"\xff\xdd")
        v = int(rng.integers(1, 300))
        d[i + 4:i + 6] = bytes([v >> 8, v & 255])
    elif kind == 3:
        del d[int(rng.integers(sos + 14, len(d) - 2)):]
    jpegs.append(bytes(d))
bad = 0
for cs in (ColorSpace.RGB, ColorSpace.GRAYSCALE, ColorSpace.RGBA, ColorSpace.YCbCr):
    opts = ZuneJpegOptions().set_out_colorspace(cs)
    want = decode_batch(jpegs, opts, threads=0)
    stats = {}
    got = decode_batch(jpegs, opts, threads=0, gpu_entropy=True, stats=stats)
    for k, (g, w_) in enumerate(zip(got, want)):
        same = (isinstance(g, DecodeErrors) and isinstance(w_, DecodeErrors) and g.status == w_.status) or (not isinstance(w_, DecodeErrors) and g == w_)
        if not same:
            bad += 1
            print("MISMATCH", cs, k)
    print(cs.name, "entropy-decoded on the GPU:", stats["gpu_entropy"], "of", n, "errors:", sum(isinstance(w_, DecodeErrors) for w_ in want), flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
