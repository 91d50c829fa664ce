"""Multi-GPU partitioning of the path.  Images (and, inside one image, strips) are independent -- every
up-sampler / IDCT / colour rule of the reference is strip-local -- so a batch shards across
ranks with NO data-path collective: each rank reconstructs its contiguous slice on its own GPU and stream.
torch.distributed is used only for the plumbing (barrier + max-over-ranks of the device time)."""
from __future__ import annotations

import os


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous, balanced slice of [0, n_items) owned by `rank` (sizes differ by at most one): zj_partition, the rule the
    library's own multi-device entry points (zj_gpu_reconstruct_multi, zj_decode_batch_multi) cut batches with."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    import ctypes as C
    from . import _ffi
    lo, hi = C.c_size_t(), C.c_size_t()
    _ffi.load().zj_partition(n_items, world, rank, C.byref(lo), C.byref(hi))
    return range(lo.value, hi.value)


def strip_ranges(img, world: int):
    """A single huge image: [(sub-image descriptor, byte offset, byte count)] of the contiguous strip ranges the devices of
    one box take (zj_image_strip_range); ranges without rows are left out.  Every rule of the path is strip-local, so the
    ranges' pixels concatenate to the whole image's."""
    import ctypes as C
    from . import _ffi
    lib = _ffi.load()
    ns = C.c_uint32()
    rc = lib.zj_image_strip_range(C.byref(img), 0, 0, None, None, None, C.byref(ns))
    if rc:
        raise ValueError(f"zj_image_strip_range: {rc}")
    out = []
    for r in range(world):
        p = partition(ns.value, world, r)
        if len(p) == 0 and not (ns.value == 0 and r == 0):
            continue
        sub, off, nb = _ffi.ZjImage(), C.c_size_t(), C.c_size_t()
        rc = lib.zj_image_strip_range(C.byref(img), p.start, p.stop, C.byref(sub), C.byref(off), C.byref(nb), None)
        if rc:
            raise ValueError(f"zj_image_strip_range: {rc}")
        if nb.value:
            out.append((sub, off.value, nb.value))
    return out


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


class Plumbing:
    """Barrier and max-reduction over ranks via torch.distributed (nccl on GPUs, gloo on CPU)."""

    def __init__(self, backend: str | None = None):
        self.rank, self.local_rank, self.world = env_world()
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device(f"cuda:{self.local_rank}")
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
            self.dist, self.torch, self.backend = dist, torch, backend

    def _tensor(self, v):
        dev = f"cuda:{self.local_rank}" if self.backend == "nccl" else "cpu"
        return self.torch.tensor([float(v)], dtype=self.torch.float64, device=dev)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, v: float) -> float:
        if self.dist is None:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if self.dist is None:
            return float(v)
        t = self._tensor(v)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
