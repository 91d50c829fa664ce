"""N>1 path on CPU: world_size-2 gloo processes shard a batch with no data-path collective and agree on the
max-over-ranks timing and the summed unit count (the only things that cross ranks)."""
import os
import socket
import sys

import pytest

from zune_jpeg_b200.sharding import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 256, 1024, 1031):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                p = partition(n, world, r)
                got.extend(p)
            assert got == list(range(n))
            sizes = [len(partition(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from zune_jpeg_b200.sharding import Plumbing, partition
    pl = Plumbing(backend="gloo")
    mine = partition(37, pl.world, pl.rank)
    pl.barrier()
    t = pl.max(10.0 + 5.0 * rank)      # per-rank device time -> the job takes the slowest
    total = pl.sum(len(mine))          # units all ranks processed
    pl.close()
    q.put((rank, list(mine), t, total))


def test_two_rank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == list(range(37))
    assert res[0][2] == res[1][2] == 15.0
    assert res[0][3] == res[1][3] == 37.0


# ------------------------------------------------------------------ strips of one image over several devices
import numpy as np

import oracle
import util

QTS = [util.std_qt(False), util.std_qt(True), util.std_qt(True)]
MODES = {"444": (1, 1), "422": (2, 1), "440": (1, 2), "420": (2, 2)}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("out_cs", [0, 1, 2])
def test_strip_ranges_concatenate_to_the_whole_image(mode, out_cs):
    """The claim the strip split rests on, checked against the ORACLE (CPU): reconstructing contiguous strip ranges as images
    of their own gives exactly the bytes of the whole image -- odd heights (dropped MCU row, partial last strip) included."""
    from zune_jpeg_b200.sharding import strip_ranges
    rng = np.random.default_rng(hash((mode, out_cs)) & 0xFFFF)
    hs, vs = MODES[mode]
    for (w, h) in [(256, 256), (200, 203), (520, 77), (64, 41)]:
        planes = util.random_planes(rng, w, h, 3, hs, vs)
        for variant in (0, 1):
            img = util.make_image(w, h, planes, QTS, hs, vs, out_cs, variant)
            try:
                whole = oracle.reconstruct(img)
            except RuntimeError:
                continue        # a geometry the reference panics on
            for world in (2, 3, 8):
                got = np.full(len(whole), 0xEE, np.uint8)
                covered = 0
                for sub, off, nb in strip_ranges(img, world):
                    piece = oracle.reconstruct(sub)
                    assert len(piece) == nb
                    got[off:off + nb] = piece
                    covered += nb
                assert covered == len(whole) and np.array_equal(got, whole), f"{w}x{h} {mode} out={out_cs} var={variant} world={world}"


def _devices(gpu, want):
    n = gpu.device_count()
    return list(range(min(n, want))) if n >= 2 else [0] * want     # one device: the shares run side by side on it


@pytest.mark.gpu
@pytest.mark.parametrize("n_dev", [2, 4])
def test_reconstruct_multi_images_and_strips(n_dev):
    """zj_gpu_reconstruct_multi: a batch larger than the device list goes by image ranges, a single image by strip ranges;
    both must equal the oracle.  (On a one-GPU box the device list repeats device 0: same code path, shares run concurrently.)"""
    from zune_jpeg_b200 import gpu
    rng = np.random.default_rng(40 + n_dev)
    devs = _devices(gpu, n_dev)
    cases = [(640, 352, "420", 0), (333, 131, "444", 0), (1000, 64, "422", 5), (72, 40, "440", 1), (520, 203, "420", 0), (64, 64, "420", 2), (1288, 96, "420", 0)]
    imgs, wants, keep = [], [], []
    for (w, h, mode, out_cs) in cases:
        hs, vs = MODES[mode]
        planes = util.random_planes(rng, w, h, 3, hs, vs)
        keep.append(planes)
        imgs.append(util.make_image(w, h, planes, QTS, hs, vs, out_cs, 0))
        wants.append(oracle.reconstruct(imgs[-1]))
    for got, want in zip(gpu.reconstruct_multi(imgs, devs), wants):
        assert np.array_equal(got, want)
    # one large image: strips over the devices
    w, h = 2048, 1531
    planes = util.random_planes(rng, w, h, 3, 2, 2)
    img = util.make_image(w, h, planes, QTS, 2, 2, 0, 0)
    want = oracle.reconstruct(img, threads=4)
    before = gpu.launch_count()
    got = gpu.reconstruct_multi([img], devs)[0]
    assert np.array_equal(got, want)
    assert gpu.launch_count() - before == len(devs)      # one launch per strip range


@pytest.mark.gpu
def test_decode_batch_multi():
    """JPEG bytes over a device list: image ranges (n >= devices) and one image with restart markers cut into strip ranges."""
    import ctypes as C

    import jpeg_util
    from zune_jpeg_b200 import _ffi, gpu
    from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions
    lib = _ffi.load()
    devs = _devices(gpu, 2)

    def run(datas):
        n = len(datas)
        opt = _ffi.ZjOptions()
        lib.zj_options_default(C.byref(opt))
        opt.num_threads = 4
        bufs = (C.c_void_p * n)(*[C.cast(C.c_char_p(d), C.c_void_p).value for d in datas])
        lens = (C.c_size_t * n)(*[len(d) for d in datas])
        outs = (C.c_void_p * n)()
        olens = (C.c_size_t * n)()
        status = (C.c_int * n)()
        rc = lib.zj_decode_batch_multi(C.byref(opt), (C.c_int * len(devs))(*devs), len(devs), bufs, lens, n, outs, olens, status)
        assert rc == 0 and list(status) == [0] * n
        res = []
        for i in range(n):
            res.append(np.ctypeslib.as_array((C.c_uint8 * olens[i]).from_address(outs[i])).copy())
            lib.zj_buffer_free(outs[i])
        return res

    def expect(data):
        d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace.RGB))
        img, planes = d.decode_coefficients(data)
        return oracle.reconstruct(img)

    datas = [jpeg_util.synth_jpeg(10 + i, 320 + 16 * i, 200 + 8 * i, ("420", "444", "422")[i % 3], 90) for i in range(5)]
    for got, data in zip(run(datas), datas):
        assert np.array_equal(got, expect(data))
    big = jpeg_util.synth_jpeg(3, 1024, 777, "420", 90, restart_rows=1)
    assert np.array_equal(run([big])[0], expect(big))
