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
