"""The C-ABI library loads, exports every symbol include/zune_jpeg_b200.h declares, validates descriptors the
way the oracle does, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from zune_jpeg_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QTS = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _ffi.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "zune_jpeg_b200.h")).read()
    declared = set(re.findall(r"ZJ_API[^;(]*?\b(zj_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes():
    assert C.sizeof(_ffi.ZjComponent) == 8 + 8 + 256 + 16
    assert C.sizeof(_ffi.ZjImage) == 24 + 3 * C.sizeof(_ffi.ZjComponent)
    assert C.sizeof(_ffi.ZjOptions) == 32


def test_validate_matches_oracle_statuses(lib):
    import oracle
    rng = np.random.default_rng(0)
    seen = set()
    for (w, h, hs, vs, out_cs, n_comp) in [(64, 64, 2, 2, 0, 3), (33, 17, 1, 1, 1, 3), (100, 70, 2, 1, 1, 3), (9, 64, 1, 1, 1, 1),
                                           (64, 64, 1, 1, 1, 1), (640, 33, 2, 2, 5, 3), (12, 40, 2, 2, 0, 3), (5, 5, 1, 1, 0, 3)]:
        planes = util.random_planes(rng, w, h, n_comp, hs, vs)
        img = util.make_image(w, h, planes, QTS[:n_comp], hs, vs, out_cs, 0)
        rc = lib.zj_validate_image(C.byref(img))
        try:
            oracle.reconstruct(img)
            want = 0
        except RuntimeError as e:
            want = int(str(e).split("rc=")[1])
        assert rc == want, (w, h, hs, vs, out_cs, rc, want)
        seen.add(rc)
    assert 0 in seen and -5 in seen


def test_validate_rejects_bad_descriptors(lib):
    rng = np.random.default_rng(1)
    planes = util.random_planes(rng, 64, 64, 3, 2, 2)
    ok = lambda: util.make_image(64, 64, planes, QTS, 2, 2, 0, 0)
    img = ok(); img.n_comp = 2
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_INVALID_ARG
    img = ok(); img.comp[0].h_samp = 4
    assert lib.zj_validate_image(C.byref(img)) in (_ffi.ERR_UNSUPPORTED, _ffi.ERR_INVALID_ARG)
    img = ok(); img.comp[1].h_samp = 2
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_UNSUPPORTED  # chroma must be 1x1 (decoder.rs:634)
    img = ok(); img.comp[0].width_stride = 60
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_INVALID_ARG
    img = ok(); img.comp[2].n_i16 = 10
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_SHORT_PLANE
    img = ok(); img.out_cs = 9
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_INVALID_ARG
    img = ok(); img.width = 0
    assert lib.zj_validate_image(C.byref(img)) == _ffi.ERR_INVALID_ARG
    assert lib.zj_output_size(C.byref(ok())) == 64 * 64 * 3


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute entry points must fail loudly, not compute on the CPU."""
    if lib.zj_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    rng = np.random.default_rng(2)
    planes = util.random_planes(rng, 64, 64, 3, 2, 2)
    img = util.make_image(64, 64, planes, QTS, 2, 2, 0, 0)
    out = np.zeros(64 * 64 * 3, np.uint8)
    ptrs = (C.c_void_p * 1)(out.ctypes.data)
    lens = (C.c_size_t * 1)(out.size)
    arr = (_ffi.ZjImage * 1)(img)
    assert lib.zj_gpu_reconstruct(0, None, arr, 1, ptrs, lens) == _ffi.ERR_NO_DEVICE
    assert not out.any()
    plan = C.c_void_p()
    assert lib.zj_batch_create(0, arr, 1, ptrs, lens, C.byref(plan)) == _ffi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.zj_gpu_strerror(_ffi.ERR_NO_DEVICE)


def test_product_does_not_touch_oracle():
    """Nothing under the package may import, link or load oracle/ (it is the checker)."""
    pkg = os.path.join(ROOT, "zune-jpeg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "zj_oracle" not in txt and "libzj_oracle" not in txt and "import oracle" not in txt, f


@pytest.mark.gpu
def test_release_caches_while_calls_are_in_flight():
    """zj_release_device_caches / zj_release_host_caches may be called at any time: here from a thread of its own while four
    threads run the batch front doors (host route, GPU-entropy route, device outputs).  Results stay correct, nothing crashes,
    and after a final release the next call simply allocates again."""
    import threading

    import numpy as np

    import jpeg_util
    from zune_jpeg_b200 import _ffi, gpu
    from zune_jpeg_b200.decoder import ColorSpace, ZuneJpegOptions, decode_batch
    lib = _ffi.load()
    opts = ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB)
    datas = [jpeg_util.synth_jpeg(20 + i, 640 + 32 * i, 360 + 16 * i, ("420", "444", "422")[i % 3], 90, restart_rows=i % 2) for i in range(6)]
    want = decode_batch(datas, opts, threads=2)
    assert all(isinstance(w, bytes) for w in want)
    stop = threading.Event()
    errors = []

    def releaser():
        while not stop.is_set():
            lib.zj_release_device_caches()
            lib.zj_release_host_caches()

    def worker(k):
        try:
            for it in range(6):
                got = decode_batch(datas, opts, threads=2, gpu_entropy=bool((k + it) & 1))
                if got != want:
                    errors.append((k, it, "pixels differ"))
        except Exception as e:      # noqa: BLE001
            errors.append((k, repr(e)))

    r = threading.Thread(target=releaser)
    ws = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    r.start()
    for w in ws:
        w.start()
    for w in ws:
        w.join()
    stop.set()
    r.join()
    assert not errors, errors[:3]
    lib.zj_release_device_caches()
    assert decode_batch(datas, opts, threads=2, gpu_entropy=True) == want
