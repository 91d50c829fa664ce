"""Multi-GPU partitioning of the path.  Images (and, inside one image, strips) are independent -- every
up-sampler / IDCT / colour rule of the reference is strip-local -- so a batch shards across
ranks with NO data-path collective: each rank reconstructs its contiguous slice on its own GPU and stream.
torch.distributed is used only for the plumbing (barrier + max-over-ranks of the device time)."""
from __future__ import annotations

import os


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous, balanced slice of [0, n_items) owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def strip_partition(n_strips: int, world: int, rank: int) -> range:
    """A single huge image: contiguous strip ranges per GPU (restart intervals make the host entropy stage
    splittable on the same boundaries)."""
    return partition(n_strips, world, rank)


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
