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
