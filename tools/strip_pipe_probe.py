#!/usr/bin/env python
"""tools/strip_pipe_probe.py [w h] -- one large baseline JPEG with restart markers through Decoder.decode_into: one shot
(ZJ_STRIP_RANGES=1) against the strip pipeline, sequential and interval-parallel host stage.  ZJ_PIPE_TRACE=1 prints when each
strip range was queued."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import jpeg_util
from zune_jpeg_b200 import gpu
from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 8192)
data = jpeg_util.synth_jpeg(1, w, h, "420", 90, restart_rows=1)
pin = gpu.PinnedBuffer(w * h * 3)
for nt in (1, os.cpu_count() or 4):
    d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB).set_num_threads(nt))
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        d.decode_into(data, pin.array)
        best = min(best, time.perf_counter() - t0)
    print(f"ranges={os.environ.get('ZJ_STRIP_RANGES', '8')} threads={nt}: {best * 1e3:.2f} ms, intervals side by side {d.entropy_segments()}", flush=True)
