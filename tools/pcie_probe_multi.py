#!/usr/bin/env python
"""tools/pcie_probe_multi.py N [GB] -- what the host links give when N GPUs of one box copy at the same time: one process per
GPU (as under torchrun), pinned 1 GB (or GB) buffers, H2D and D2H at once on two streams, all processes released together by a
barrier.  Prints GB/s per direction per GPU and the aggregate; run for N = 1, 2, 4, 8 to see where the e2e leg of bench.py
stops scaling (host DRAM / PCIe root / IOMMU side of the box, not the kernels)."""
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def worker(rank, n_proc, gb, barrier, q):
    torch.cuda.set_device(rank)
    n = gb << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):          # first pass = warm-up
            torch.cuda.synchronize()
            barrier.wait()
            t = time.perf_counter()
            for _ in range(6):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            res[mode] = 6 * n / (time.perf_counter() - t) / 1e9
            barrier.wait()
    q.put((rank, res))


def main():
    n_proc = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    gb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n_proc), ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, n_proc, gb, barrier, q)) for r in range(n_proc)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join()
    for mode in ("h2d", "d2h", "both"):
        per = [r[1][mode] for r in res]
        print(f"N={n_proc} {mode:5s}: per GPU {' '.join('%.1f' % v for v in per)} GB/s each way -> aggregate {sum(per):.1f} GB/s per direction", flush=True)


if __name__ == "__main__":
    main()
