"""The literal oracle (oracle/zj_oracle.c) against the independently derived closed-form model
(tests/ref_model.py): two readings of the reference that must agree bit-for-bit, panics included."""
import numpy as np
import pytest

import oracle
import ref_model as M
import util

QTS = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]


def _both(w, h, hs, vs, out_cs, variant, planes, n_comp=3, progressive=False):
    img = util.make_image(w, h, planes, QTS[:n_comp], hs, vs, out_cs, variant, progressive)
    try:
        a = oracle.reconstruct(img)
    except RuntimeError as e:
        a = "panic" if "rc=-5" in str(e) else str(e)
    try:
        b = M.reconstruct(w, h, planes, QTS[:n_comp], hs, vs, out_cs, variant, progressive)
    except M.RefPanic:
        b = "panic"
    if isinstance(a, str) or isinstance(b, str):
        assert isinstance(a, str) and isinstance(b, str) and a == b, (w, h, hs, vs, out_cs, variant, a if isinstance(a, str) else "ok", b if isinstance(b, str) else "ok")
        return
    assert np.array_equal(a, b), (w, h, hs, vs, out_cs, variant, np.nonzero(a != b)[0][:8])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("hv", [(1, 1), (2, 1), (1, 2), (2, 2)])
def test_random_images(hv, variant):
    rng = np.random.default_rng(100 + 10 * hv[0] + hv[1] + variant)
    for (w, h) in [(64, 64), (100, 70), (1000, 96), (33, 17), (520, 40), (16, 16), (7, 9), (264, 130)]:
        for out_cs in (0, 5, 6, 1, 2, 3):
            planes = util.random_planes(rng, w, h, 3, hv[0], hv[1])
            _both(w, h, hv[0], hv[1], out_cs, variant, planes)


@pytest.mark.parametrize("variant", [0, 1])
def test_extreme_coefficients(variant):
    rng = np.random.default_rng(7)
    for hv in [(1, 1), (2, 1), (1, 2), (2, 2)]:
        planes = util.random_planes(rng, 520, 64, 3, hv[0], hv[1], extreme=True)
        _both(520, 64, hv[0], hv[1], 0, variant, planes)


@pytest.mark.parametrize("variant", [0, 1])
def test_grayscale(variant):
    rng = np.random.default_rng(8)
    for (w, h) in [(64, 64), (1001, 40), (57, 9), (200, 8), (9, 64)]:
        planes = util.random_planes(rng, w, h, 1, 1, 1)
        _both(w, h, 1, 1, 1, variant, planes, n_comp=1)


def test_dropped_rows_q1():
    """H and HV with an odd MCU-row count leave the last MCU row zero."""
    rng = np.random.default_rng(9)
    for (hs, vs, h) in [(2, 1, 24), (2, 2, 48)]:
        planes = util.random_planes(rng, 64, h, 3, hs, vs)
        img = util.make_image(64, h, planes, QTS, hs, vs, 0, 0)
        out = oracle.reconstruct(img).reshape(h, 64 * 3)
        rows = 8 * vs
        assert (out[h - rows:] == 0).all() and out[: h - rows].any()


def test_threads_match_serial():
    rng = np.random.default_rng(10)
    planes = util.random_planes(rng, 640, 480, 3, 2, 2)
    img = util.make_image(640, 480, planes, QTS, 2, 2, 0, 0)
    assert np.array_equal(oracle.reconstruct(img, threads=1), oracle.reconstruct(img, threads=5))
