"""Restart-interval-parallel entropy decode -- no GPU needed.

The parallel form must be indistinguishable from the reference's sequential MCU loop (mcu.rs:253-351, 386-418):
same coefficient planes, same errors, also on streams where the reference's restart bookkeeping misbehaves (Q8:
the countdown ticks once per component) or the data is damaged.  `set_num_threads(1)` runs the sequential loop."""
import numpy as np
import pytest

import jpeg_util
from zune_jpeg_b200.decoder import ColorSpace, DecodeErrors, Decoder, ZuneJpegOptions


def _decode(data, threads, out_cs=ColorSpace.RGB):
    d = Decoder.new_with_options(ZuneJpegOptions().set_num_threads(threads).set_out_colorspace(out_cs))
    try:
        img, planes = d.decode_coefficients(data)
    except DecodeErrors as e:
        return ("error", e.variant, e.message), d.entropy_segments()
    return ("ok", [p.tobytes() for p in planes]), d.entropy_segments()


def _same(data, threads=4, out_cs=ColorSpace.RGB):
    seq, n0 = _decode(data, 1, out_cs)
    par, n1 = _decode(data, threads, out_cs)
    assert n0 == 0
    assert seq == par
    return n1, seq


@pytest.mark.parametrize("sub,gray,w,h,rows", [
    ("420", False, 1024, 768, 1),     # DRI = one MCU row (BASELINE configs[4] style)
    ("420", False, 1000, 1016, 2),    # odd number of MCU rows (Q1: the last one is never decoded), ragged width
    ("422", False, 1280, 720, 1),
    ("444", False, 800, 600, 3),
    ("420", True, 2048, 1024, 1),     # single component
])
def test_parallel_equals_sequential(sub, gray, w, h, rows):
    data = jpeg_util.synth_jpeg(11, w, h, sub, quality=85, gray=gray, restart_rows=rows)
    assert b"\xff\xdd" in data
    n, res = _same(data)
    assert res[0] == "ok"
    assert n >= 2, "the parallel path did not run"


def test_gray_output_of_colour_image():
    # out_colorspace GRAYSCALE: chroma blocks are decoded into a scratch block (mcu.rs:244, 316-320)
    data = jpeg_util.synth_jpeg(12, 1024, 512, "420", restart_rows=1)
    n, res = _same(data, out_cs=ColorSpace.GRAYSCALE)
    assert res[0] == "ok" and n >= 2


def _set_dri(data: bytes, dri: int) -> bytes:
    i = data.index(b"\xff\xdd")
    return data[:i + 4] + bytes([dri >> 8, dri & 255]) + data[i + 6:]


@pytest.mark.parametrize("dri", [1, 2, 5, 7, 63, 64, 65, 200])
def test_mismatched_dri_falls_back(dri):
    """A DRI that does not match where the encoder put its markers: the reference's loop resets at the wrong places
    or not at all; whatever it makes of the stream, both forms must make the same."""
    data = _set_dri(jpeg_util.synth_jpeg(13, 1024, 512, "420", quality=80, restart_rows=1), dri)
    _same(data)


@pytest.mark.parametrize("seed", range(8))
def test_damaged_streams(seed):
    rng = np.random.default_rng(seed)
    data = bytearray(jpeg_util.synth_jpeg(14, 1024, 640, "420", quality=80, restart_rows=1))
    sos = bytes(data).index(b"\xff\xda")
    kind = seed % 4
    if kind == 0:      # flipped bytes inside the entropy-coded data
        for p in rng.integers(sos + 20, len(data) - 2, size=6):
            data[p] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:    # truncated
        del data[int(rng.integers(sos + 20, len(data) - 2)):]
    elif kind == 2:    # a restart marker removed
        marks = [i for i in range(sos, len(data) - 1) if data[i] == 0xFF and 0xD0 <= data[i + 1] <= 0xD7]
        p = marks[int(rng.integers(0, len(marks)))]
        del data[p:p + 2]
    else:              # a marker inside the scan that is not RSTn
        p = int(rng.integers(sos + 20, len(data) - 2))
        data[p:p] = b"\xff\xd9"
    _same(bytes(data))


def test_small_intervals_many_segments():
    # DRI of a few MCUs as the encoder wrote it (Pillow has no knob: patch the rows-based stream's geometry instead):
    # a 16-px-high... image: one MCU row per interval, 64 MCUs wide -> 3 * 64 ticks per interval
    data = jpeg_util.synth_jpeg(15, 1024, 4096, "420", quality=75, restart_rows=1)
    n, res = _same(data, threads=8)
    assert res[0] == "ok" and n == 4096 // 16


def test_batch_uses_spare_threads():
    from zune_jpeg_b200 import _ffi
    assert hasattr(_ffi.load(), "zj_decoder_entropy_segments")


def test_fill_bytes_before_markers():
    """0xFF fill bytes in front of RSTn (legal, B.1.1.2): the marker scan and the bit reader (bitstream.rs:200-215) agree on where
    the next interval starts, so the intervals still run side by side."""
    data = bytearray(jpeg_util.synth_jpeg(16, 1024, 512, "420", quality=85, restart_rows=1))
    sos = bytes(data).index(b"\xff\xda")
    marks = [i for i in range(sos, len(data) - 1) if data[i] == 0xFF and 0xD0 <= data[i + 1] <= 0xD7]
    for k, p in enumerate(reversed(marks)):
        data[p:p] = b"\xff" * (1 + k % 3)
    n, res = _same(bytes(data))
    assert res[0] == "ok" and n == 512 // 16
    clean, _ = _decode(jpeg_util.synth_jpeg(16, 1024, 512, "420", quality=85, restart_rows=1), 1)
    assert res == clean
