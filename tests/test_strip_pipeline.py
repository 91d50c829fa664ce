"""SURVEY 8(f).2 for one image: zj_decoder_decode_into streams finished strip ranges through the GPU while the host is still
entropy-decoding (reference src/mcu.rs:230-369).  The pixels must be those of the one-shot path / the oracle, for the sequential
loop, for restart intervals side by side, and when the interval-parallel form is turned down half-way (planes redone)."""
import numpy as np
import pytest

import jpeg_util
import oracle

pytestmark = pytest.mark.gpu


def _expect(data, out_cs=0):
    from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions
    d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace(out_cs)).set_num_threads(1))
    img, planes = d.decode_coefficients(data)
    return oracle.reconstruct(img, threads=4)


@pytest.mark.parametrize("sub,rst,threads", [("420", 0, 1), ("420", 1, 8), ("422", 2, 8), ("444", 1, 4), ("420", 1, 1)])
def test_decode_into_pipeline_matches_oracle(sub, rst, threads):
    from zune_jpeg_b200 import gpu
    from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions
    w, h = 2560, 1723                    # 4.4 MP: above the pipeline's threshold; odd height (partial last strip)
    data = jpeg_util.synth_jpeg(77, w, h, sub, 90, restart_rows=rst)
    want = _expect(data)
    pin = gpu.PinnedBuffer(len(want) + 64)
    pin.array[:] = 0xCD
    d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB).set_num_threads(threads))
    before = gpu.launch_count()
    n = d.decode_into(data, pin.array)
    assert n == len(want)
    assert gpu.launch_count() - before >= 8          # one launch per strip range, not one for the image
    assert np.array_equal(pin.array[:n], want)
    assert np.all(pin.array[n:] == 0xCD)
    if rst and threads > 1:
        assert d.entropy_segments() > 1
    # pageable destination and the allocating front door take the same route
    assert d.decode_buffer(data) == want.tobytes()
    small = jpeg_util.synth_jpeg(5, 320, 200, sub, 90, restart_rows=rst)
    before = gpu.launch_count()
    assert d.decode_buffer(small) == _expect(small).tobytes()
    assert gpu.launch_count() - before == 1          # below the threshold: one shot


def test_pipeline_restart_when_intervals_are_turned_down():
    """A stream whose DRI does not match its markers: the interval-parallel form starts, reports progress, is turned down, and
    the sequential loop redoes the planes -- ranges already queued must not survive in the output."""
    from zune_jpeg_b200 import gpu
    from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions
    w, h = 2560, 1723
    data = bytearray(jpeg_util.synth_jpeg(78, w, h, "420", 90, restart_rows=1))
    i = data.find(b"\xff\xdd")
    assert i > 0
    dri = int.from_bytes(data[i + 4:i + 6], "big")
    # corrupt one RSTn marker two thirds into the stream into stuffed data: the interval before it no longer ends with a reset
    marks = [k for k in range(i + 6, len(data) - 1) if data[k] == 0xFF and 0xD0 <= data[k + 1] <= 0xD7]
    k = marks[(2 * len(marks)) // 3]
    data[k + 1] = 0x00
    data = bytes(data)
    d1 = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB).set_num_threads(1))
    try:
        want = d1.decode_buffer(data)
    except Exception as e:               # the sequential loop may reject the stream: then the pipelined call must too
        want = e
    d8 = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB).set_num_threads(8))
    pin = gpu.PinnedBuffer(w * h * 3)
    if isinstance(want, Exception):
        with pytest.raises(type(want)):
            d8.decode_into(data, pin.array)
    else:
        n = d8.decode_into(data, pin.array)
        assert pin.array[:n].tobytes() == want
        assert d8.entropy_segments() == 0        # the sequential loop produced the planes
