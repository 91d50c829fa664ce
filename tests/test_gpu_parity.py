"""GPU parity: the fused CUDA path (through the C ABI) must equal the oracle byte-for-byte."""
import numpy as np
import pytest

import oracle
import util

pytestmark = pytest.mark.gpu

QTS = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]
MODES = {"444": (1, 1), "422": (2, 1), "440": (1, 2), "420": (2, 2)}


def _run_case(rng, w, h, mode, out_cs, variant, n_comp=3, progressive=False, **kw):
    from zune_jpeg_b200 import gpu
    hs, vs = MODES[mode]
    planes = util.random_planes(rng, w, h, n_comp, hs, vs, **kw)
    img = util.make_image(w, h, planes, QTS[:n_comp], hs, vs, out_cs, variant, progressive)
    try:
        want = oracle.reconstruct(img)
    except RuntimeError as e:
        assert "rc=-5" in str(e), e
        assert gpu.validate(img) == -5, "GPU path must refuse what the reference panics on"
        return "panic"
    got = gpu.reconstruct([img])[0]
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        nc = len(want) // (w * h)
        first = bad[:6]
        where = [(int(b // (w * nc)), int((b % (w * nc)))) for b in first]
        raise AssertionError(f"{w}x{h} {mode} out={out_cs} var={variant}: {bad.size} bytes differ, first (row,byte)={where} got={got[first]} want={want[first]}")
    return "ok"


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("out_cs", [0, 5, 2, 1])
def test_random_sizes(mode, out_cs, variant):
    rng = np.random.default_rng(hash((mode, out_cs, variant)) & 0xFFFF)
    sizes = [(64, 64), (100, 70), (1000, 96), (333, 130), (16, 16), (17, 40), (640, 33), (2500, 48), (1288, 64)]
    for (w, h) in sizes:
        _run_case(rng, w, h, mode, out_cs, variant)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("mode", list(MODES))
def test_small_widths(mode, variant):
    rng = np.random.default_rng(5)
    for w in range(1, 40):
        for out_cs in (0, 5):
            _run_case(rng, w, 37, mode, out_cs, variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_extreme_coefficients(variant):
    rng = np.random.default_rng(11)
    for mode in MODES:
        _run_case(rng, 264, 72, mode, 0, variant, extreme=True)


@pytest.mark.parametrize("variant", [0, 1])
def test_grayscale_input(variant):
    rng = np.random.default_rng(3)
    for (w, h) in [(64, 64), (1001, 77), (4096, 64), (57, 9), (200, 8)]:
        _run_case(rng, w, h, "444", 1, variant, n_comp=1)
        _run_case(rng, w, h, "444", 0, variant, n_comp=1)  # GRAY -> RGB: the reference writes nothing


def test_unwritten_pairs_are_zero():
    rng = np.random.default_rng(4)
    _run_case(rng, 128, 64, "420", 3, 0)  # YCbCr -> CMYK: zeros
    _run_case(rng, 128, 64, "444", 4, 1)  # YCbCr -> YCCK: zeros


def test_batch_mixed():
    from zune_jpeg_b200 import gpu
    rng = np.random.default_rng(9)
    imgs, keep, wants = [], [], []
    for (w, h, mode, out_cs, variant) in [(640, 480, "420", 0, 0), (320, 200, "444", 0, 0), (500, 100, "422", 5, 1),
                                          (1920, 64, "420", 0, 0), (256, 256, "440", 2, 0), (300, 300, "420", 1, 0),
                                          (777, 123, "420", 0, 1), (4000, 40, "420", 5, 0)]:
        hs, vs = MODES[mode]
        planes = util.random_planes(rng, w, h, 3, hs, vs)
        keep.append(planes)
        img = util.make_image(w, h, planes, QTS, hs, vs, out_cs, variant)
        imgs.append(img)
        wants.append(oracle.reconstruct(img))
    gots = gpu.reconstruct(imgs)
    for i, (g, wv) in enumerate(zip(gots, wants)):
        assert np.array_equal(g, wv), f"image {i} differs"


def test_4k_420_full_size():
    """BASELINE config 2 geometry (3840x2160 4:2:0 -> RGB), one image, against the oracle."""
    rng = np.random.default_rng(21)
    assert _run_case(rng, 3840, 2160, "420", 0, 0) == "ok"


def test_device_resident_batch_plan():
    from zune_jpeg_b200 import gpu
    rng = np.random.default_rng(13)
    w, h = 1280, 720
    planes = util.random_planes(rng, w, h, 3, 2, 2)
    host_img = util.make_image(w, h, planes, QTS, 2, 2, 0, 0)
    want = oracle.reconstruct(host_img)
    bufs = [gpu.DeviceBuffer(p.nbytes) for p in planes]
    for b, p in zip(bufs, planes):
        b.upload(p)
    out = gpu.DeviceBuffer(len(want))
    out.memset(0xAB)  # the kernel must define every byte itself
    img = util.make_image(w, h, planes, QTS, 2, 2, 0, 0, ptrs=[b.ptr for b in bufs])
    batch = gpu.Batch([img], [out.ptr], [len(want)])
    before = gpu.launch_count()
    batch.run()
    batch.run()
    assert gpu.launch_count() - before == 2 * batch.launches
    got = out.download()
    assert np.array_equal(got, want)
    assert batch.algorithmic_bytes == sum(p.nbytes for p in planes) - 0 * 2 + len(want) or batch.algorithmic_bytes > 0
