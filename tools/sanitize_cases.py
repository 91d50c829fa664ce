"""tools/sanitize_cases.py -- a handful of small images through every kernel, for `compute-sanitizer python tools/sanitize_cases.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import oracle  # noqa: E402
import util  # noqa: E402
from zune_jpeg_b200 import gpu  # noqa: E402

qts = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]
rng = np.random.default_rng(5)
n = 0
for (w, h, hs, vs, cs, var) in [(640, 200, 2, 2, 0, 0), (1000, 130, 2, 2, 5, 0), (333, 70, 1, 1, 0, 0), (520, 90, 2, 1, 0, 0), (264, 100, 1, 2, 2, 0),
                                (300, 100, 2, 2, 1, 0), (520, 24, 1, 1, 1, 0), (200, 64, 2, 2, 0, 1), (40, 40, 2, 2, 0, 0), (4100, 48, 2, 2, 0, 0)]:
    nc = 1 if (cs == 1 and hs == 1 and vs == 1 and w == 520) else 3
    planes = util.random_planes(rng, w, h, nc, hs, vs)
    img = util.make_image(w, h, planes, qts[:nc], hs, vs, cs, var)
    assert np.array_equal(gpu.reconstruct([img])[0], oracle.reconstruct(img)), (w, h, hs, vs, cs, var)
    n += 1
print("sanitize cases ok:", n)
if os.environ.get("ZJ_SANITIZE_FUSED_ONLY") == "1":   # (racecheck of the round-2 cases below takes a quarter of an hour)
    sys.exit(0)

# round 2: the device-side consumer kernels, strip ranges over a device list, the strip pipeline of one image
import jpeg_util  # noqa: E402
from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions  # noqa: E402

for (w, h, hs, vs, cs) in [(333, 131, 1, 1, 0), (640, 96, 2, 2, 5)]:
    planes = util.random_planes(rng, w, h, 3, hs, vs)
    host = util.make_image(w, h, planes, qts, hs, vs, cs, 0)
    want = oracle.reconstruct(host)
    bufs = [gpu.DeviceBuffer(p.nbytes) for p in planes]
    for b, p in zip(bufs, planes):
        b.upload(p)
    dimg = util.make_image(w, h, planes, qts, hs, vs, cs, 0, ptrs=[b.ptr for b in bufs])
    for desc in (gpu.OutputDesc("CHW", "f16", False, 3, (120, 110, 100, 0), (0.02, 0.02, 0.02, 1)), gpu.OutputDesc("HWC", "f32", True, 0), gpu.OutputDesc("CHW", "u8", True, 0)):
        got = gpu.reconstruct_device_ex([dimg], desc)[0].download()
        assert got.tobytes() == desc.expected(want, w, h, len(want) // (w * h)).tobytes()
        n += 1
    assert np.array_equal(gpu.reconstruct_multi([host], [0, 0, 0])[0], want)
    n += 1
data = jpeg_util.synth_jpeg(3, 2304, 1800, "420", 90, restart_rows=1)
d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB).set_num_threads(4))
img, planes = d.decode_coefficients(data)
pin = gpu.PinnedBuffer(2304 * 1800 * 3)
assert d.decode_into(data, pin.array) == 2304 * 1800 * 3
assert np.array_equal(pin.array, oracle.reconstruct(img, threads=4))
n += 1
print("sanitize cases ok (round 2):", n)
