"""Host front-end (headers + Huffman entropy decode, C++) -- no GPU needed: it only produces coefficient
planes.  Checks: the reference's own malformed-input tests (tests/invalid_images.rs), header info, and that
the decoded coefficients are right (pixels rebuilt from them by the oracle match libjpeg within the integer
IDCT / colour-conversion tolerance wherever the reference has no quirk)."""
import io
import os

import numpy as np
import pytest
from PIL import Image

import jpeg_util
import oracle
from zune_jpeg_b200.decoder import ColorSpace, DecodeErrors, Decoder, ZuneJpegOptions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- reference tests/invalid_images.rs, same inputs, same variant + message -----------------------------
@pytest.mark.parametrize("data,variant,message", [
    (bytes([0xff, 0xd8, 0xa4]), "Format", None),                                                        # eof
    (bytes([0xff, 0xd8, 0xff, 0x00, 0x00, 0x00]), "Format", "Found a marker with invalid length : 0"),  # bad_ff_marker_size
    (bytes([255, 216, 255, 218, 232, 197, 255]), "SosError", "Bad SOS length,corrupt jpeg"),            # bad_number_of_scans
    (bytes([255, 216, 255, 196, 0, 0]), "FormatStatic", "Invalid Huffman length in image"),             # huffman_length_subtraction_overflow
    (bytes([255, 216, 255, 192, 255, 1, 8, 9, 119, 48, 255, 192]), "SofError",
     "Length of start of frame differs from expected 584,value is 65281"),                             # mul_with_overflow
])
def test_reference_invalid_images(data, variant, message):
    with pytest.raises(DecodeErrors) as e:
        Decoder.new().decode_coefficients(data)
    assert e.value.variant == variant
    if message is not None:
        assert e.value.message == message


def test_more_errors():
    d = Decoder()
    with pytest.raises(DecodeErrors) as e:
        d.read_headers(b"\x89PNG\r\n")
    assert e.value.variant == "IllegalMagicBytes" and e.value.message == str(0x8950)
    with pytest.raises(DecodeErrors) as e:
        d.read_headers(b"\xff")
    assert e.value.variant == "ExhaustedData"
    with pytest.raises(DecodeErrors) as e:  # DAC = arithmetic coding (test-images/test-arithmetic-coding.jpg path)
        d.read_headers(bytes([0xff, 0xd8, 0xff, 0xcc, 0, 4, 0, 0]))
    assert e.value.variant == "Format" and "`DAC` is not supported,cannot continue" in e.value.message
    # limits (options.rs max_width / max_height)
    data = jpeg_util.synth_jpeg(1, 64, 48)
    with pytest.raises(DecodeErrors) as e:
        Decoder.new_with_options(ZuneJpegOptions().set_max_width(32)).read_headers(data)
    assert "greater than width limit 32" in e.value.message
    assert Decoder().info() is None


def test_headers_and_options():
    data = jpeg_util.synth_jpeg(2, 200, 120, "422", progressive=True)
    d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGBA))
    d.read_headers(data)
    info = d.info()
    assert (info.width, info.height, info.components, info.sof) == (200, 120, 3, 2)
    assert (d.width(), d.height()) == (200, 120)
    assert d.get_output_colorspace() == ColorSpace.RGBA
    g = Decoder()
    g.read_headers(jpeg_util.synth_jpeg(3, 64, 64, gray=True))
    assert g.get_output_colorspace() == ColorSpace.GRAYSCALE  # forced by SOF, headers.rs:278-285
    o = ZuneJpegOptions()
    assert (o.get_use_unsafe(), o.get_threads(), o.get_max_width(), o.get_max_scans(), o.get_strict_mode()) == (True, 4, 16384, 64, False)
    assert o.set_strict_mode(True).get_strict_mode() and not o.get_strict_mode()  # builder is by-value


def _oracle_pixels(data, out_cs=ColorSpace.RGB, use_unsafe=True):
    d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(out_cs).set_use_unsafe(use_unsafe))
    img, planes = d.decode_coefficients(data)
    for z in range(img.n_comp):
        img.comp[z].coeff = planes[z].ctypes.data
    return img, planes, oracle.reconstruct(img)


@pytest.mark.parametrize("progressive", [False, True])
@pytest.mark.parametrize("quality", [50, 75, 90])  # higher qualities run into host-stage quirks Q9/Q10 of the reference
def test_444_matches_libjpeg_within_tolerance(progressive, quality):
    w, h = 336, 200
    data = jpeg_util.synth_jpeg(10 + quality, w, h, "444", quality, progressive)
    img, _, out = _oracle_pixels(data)
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(int)
    diff = np.abs(out.reshape(h, w, 3).astype(int) - ref)
    # integer IDCT (+-1) and the reference's 5/6-bit colour coefficients (+-3) vs libjpeg; the last 22 pixels of
    # a row fall in the row-tail quirk Q5 and are excluded
    assert diff[:, : w - 22].max() <= 5
    assert (out.reshape(h, w * 3)[:, -16:] == 0).all()  # Q5: 16 zero bytes end every row


def test_progressive_equals_baseline_coefficients():
    for (sub, w, h) in [("444", 200, 96), ("422", 333, 130), ("420", 640, 200)]:
        im = jpeg_util.smooth_image(5, w, h)
        pb = Decoder().decode_coefficients(jpeg_util.encode(im, 90, sub, False))[1]
        pp = Decoder().decode_coefficients(jpeg_util.encode(im, 90, sub, True))[1]
        for z in range(3):
            n = min(pb[z].size, pp[z].size)  # baseline planes stop at the last processed strip (Q1)
            # the very last MCU is excluded: the baseline driver breaks out of its component loop as soon as the
            # bit reader has PREFETCHED the EOI marker (mcu.rs:337-342), so trailing blocks of the final MCU can
            # stay zero in the reference (host-stage quirk Q11)
            n -= 64 * 4
            assert n > 0 and np.array_equal(pb[z][:n], pp[z][:n]), (sub, z)


def test_grayscale_jpeg():
    w, h = 200, 64
    data = jpeg_util.synth_jpeg(6, w, h, gray=True)
    img, planes, out = _oracle_pixels(data)
    assert img.n_comp == 1 and out.size == w * h
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("L")).astype(int).reshape(-1)
    assert np.abs(out.astype(int) - ref).max() <= 2


def test_restart_markers_large_interval():
    """DRI of one MCU row or more decodes to the same coefficients as no DRI (SURVEY Q8 stays benign)."""
    im = jpeg_util.smooth_image(7, 384, 128)
    for sub in ("444", "420"):
        a = Decoder().decode_coefficients(jpeg_util.encode(im, 90, sub))[1]
        b = Decoder().decode_coefficients(jpeg_util.encode(im, 90, sub, restart_rows=1))[1]
        for z in range(3):
            n = a[z].size - 64 * 4  # final MCU excluded: EOI-prefetch quirk Q11, see above
            assert np.array_equal(a[z][:n], b[z][:n]), (sub, z)


def test_color_to_gray_only_decodes_luma():
    data = jpeg_util.synth_jpeg(8, 256, 64, "420")
    img, planes, out = _oracle_pixels(data, ColorSpace.GRAYSCALE)
    assert planes[1].size == 0 and planes[2].size == 0 and out.size == 256 * 64  # mcu.rs:244,287


def test_same_decoder_reused():
    d = Decoder()
    a = jpeg_util.synth_jpeg(9, 160, 80, "420")
    b = jpeg_util.synth_jpeg(9, 96, 200, "444", progressive=True)
    pa1 = d.decode_coefficients(a)[1]
    pb1 = d.decode_coefficients(b)[1]
    pa2 = d.decode_coefficients(a)[1]
    assert all(np.array_equal(x, y) for x, y in zip(pa1, pa2))
    assert all(np.array_equal(x, y) for x, y in zip(pb1, Decoder().decode_coefficients(b)[1]))


def test_golden_fixtures():
    """Committed JPEGs + the oracle's X86 output for them (tests/golden/make_golden.py): a regression pin of
    host stage + oracle together.  (The reference itself cannot run here, so these are oracle outputs.)"""
    import json
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    assert man["cases"]
    for case in man["cases"]:
        data = open(os.path.join(GOLDEN, case["jpeg"]), "rb").read()
        _, _, out = _oracle_pixels(data, ColorSpace[case["out"]], case["variant"] == "X86")
        want = np.fromfile(os.path.join(GOLDEN, case["pixels"]), np.uint8)
        assert np.array_equal(out, want), case["jpeg"]


def test_decode_batch_without_a_device():
    """zj_decode_batch on a machine without a GPU: the host stage still runs per image (decode errors are reported as such),
    the pixel stage reports ZJ_ERR_NO_DEVICE -- never a CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    from zune_jpeg_b200.decoder import decode_batch
    good = open(os.path.join(GOLDEN, "c420_base.jpg"), "rb").read()
    res = decode_batch([good, bytes([0xff, 0xd8, 0xa4]), good], threads=2)
    assert len(res) == 3 and all(isinstance(r, DecodeErrors) for r in res)
    assert res[0].status == -6 and res[2].status == -6      # ZJ_ERR_NO_DEVICE
    assert res[1].status == -9                                # ZJ_ERR_DECODE
    assert decode_batch([]) == []


def test_decode_batch_reports_the_references_error_per_image():
    """The batch front doors carry the reference's DecodeErrors per image (zj_batch_error_kind / zj_batch_error): the same
    variant and message Decoder.decode_buffer raises for that input -- the inputs of the reference's tests/invalid_images.rs.
    (No device needed: these inputs fail in the host stage.)"""
    from zune_jpeg_b200.decoder import decode_batch
    bad = [bytes([0xff, 0xd8, 0xa4]), bytes([0xff, 0xd8, 0xff, 0x00, 0x00, 0x00]), bytes([255, 216, 255, 218, 232, 197, 255]),
           bytes([255, 216, 255, 196, 0, 0]), bytes([255, 216, 255, 192, 255, 1, 8, 9, 119, 48, 255, 192]), b"\x89PNG\r\n"]
    for kw in ({}, {"gpu_entropy": True}):
        res = decode_batch(bad, threads=3, **kw)
        for data, r in zip(bad, res):
            with pytest.raises(DecodeErrors) as e:
                Decoder.new().decode_buffer(data)
            assert isinstance(r, DecodeErrors) and r.status == -9
            assert (r.variant, r.message) == (e.value.variant, e.value.message), (data, kw)
