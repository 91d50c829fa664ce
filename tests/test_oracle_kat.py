"""The oracle against every vector the reference's own tests hold for this path
(reference src/idct.rs:66-127 and src/upsampler.rs:118-151), in both oracle builds."""
import numpy as np
import pytest

import oracle

MAX_OUT = [0, 255, 0, 255, 0, 0, 255, 255, 255, 0, 0, 255, 0, 255, 0, 0, 0, 0, 255, 0, 255, 0, 255,
           255, 255, 255, 0, 255, 0, 255, 0, 0, 0, 0, 255, 0, 255, 0, 255, 255, 0, 255, 0, 255, 0,
           158, 0, 49, 255, 0, 255, 0, 255, 0, 255, 255, 255, 0, 255, 0, 255, 49, 255, 255]  # idct.rs:92-96
MIN_OUT = [255, 0, 255, 0, 255, 255, 0, 0, 0, 255, 255, 0, 255, 0, 255, 255, 255, 255, 0, 255, 0, 255,
           0, 0, 0, 0, 255, 0, 255, 0, 255, 255, 255, 255, 0, 255, 0, 255, 0, 0, 255, 0, 255, 0, 255,
           98, 255, 207, 0, 255, 0, 255, 0, 255, 0, 0, 0, 255, 0, 255, 0, 207, 0, 0]  # idct.rs:117-121
QT1 = np.ones(64, np.int32)


@pytest.mark.parametrize("emulated", [False, True])
@pytest.mark.parametrize("variant", [0, 1])
def test_idct_kats(variant, emulated):
    if not emulated and not oracle._has_avx2():
        pytest.skip("host has no AVX2")
    f = lambda c: oracle.idct(np.full(64, c, np.int16), QT1, 8, 1, 1, variant, emulated)
    assert (f(0) == 128).all()                                   # test_zeroes, idct.rs:69-84
    assert np.array_equal(f(32767), np.array(MAX_OUT, np.int16))  # test_max, idct.rs:88-106
    assert np.array_equal(f(-32768), np.array(MIN_OUT, np.int16))  # test_min, idct.rs:111-126


@pytest.mark.parametrize("emulated", [False, True])
def test_upsample_ramps(emulated):
    if not emulated and not oracle._has_avx2():
        pytest.skip("host has no AVX2")
    for v in (np.arange(128, dtype=np.int16), np.arange(1280, dtype=np.int16)[::-1].copy()):  # upsampler.rs:126-150
        a = oracle.upsample("horizontal_sse", v, 2 * v.size, emulated)
        b = oracle.upsample("horizontal_scalar", v, 2 * v.size, emulated)
        assert np.array_equal(a, b)


def test_kat_model_agrees():
    """the closed-form model reproduces the same KATs"""
    import ref_model as M
    for variant in (0, 1):
        for c, want in ((0, [128] * 64), (32767, MAX_OUT), (-32768, MIN_OUT)):
            got = M.idct_blocks(np.full((1, 64), c, np.int16), QT1, variant).reshape(-1)
            assert np.array_equal(got, np.array(want, np.int16))
