"""oracle built with real SSE4.1/AVX2 intrinsics == oracle built with the portable emulation (simd_compat.h)."""
import numpy as np
import pytest

import oracle
import util

pytestmark = pytest.mark.skipif(not oracle._has_avx2(), reason="host has no AVX2: only the emulated build runs")


def test_idct_avx2_real_vs_emulated():
    rng = np.random.default_rng(0)
    qt = util.std_qt(False, 50)
    for extreme in (False, True):
        planes = util.random_planes(rng, 256, 64, 1, 1, 1, extreme=extreme, dc_only_frac=0.2)
        a = oracle.idct(planes[0], qt, 256, 1, 1, 0, emulated=False)
        b = oracle.idct(planes[0], qt, 256, 1, 1, 0, emulated=True)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name,mult", [("horizontal_sse", 2), ("hv_simd", 4)])
def test_upsamplers_real_vs_emulated(name, mult):
    rng = np.random.default_rng(1)
    for w in (32, 40, 64, 120, 1920):
        x = rng.integers(0, 256, size=16 * w).astype(np.int16)
        a = oracle.upsample(name, x, mult * x.size, emulated=False)
        b = oracle.upsample(name, x, mult * x.size, emulated=True)
        assert np.array_equal(a, b)


def test_whole_image_real_vs_emulated():
    rng = np.random.default_rng(2)
    qts = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]
    for (h, v) in [(1, 1), (2, 1), (1, 2), (2, 2)]:
        planes = util.random_planes(rng, 520, 100, 3, h, v)
        img = util.make_image(520, 100, planes, qts, h, v, 0, 0)
        assert np.array_equal(oracle.reconstruct(img, emulated=False), oracle.reconstruct(img, emulated=True))
