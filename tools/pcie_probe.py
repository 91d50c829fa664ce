"""tools/pcie_probe.py -- what the host link gives (run on the GPU box): pinned H2D, D2H, and both at once, as 1 GB copies and
cut into chunks of the size zj_gpu_reconstruct moves (one plane / one image's pixels per copy)."""
import sys
import time
import torch

import os
n = int(os.environ.get("PROBE_GB", "1")) << 30   # PROBE_GB=6: buffers of the size the bench's e2e leg moves per step
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunk, reps=5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        for o in range(0, n, chunk):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return reps * n / dt / 1e9


run(True, True, n, 1)
for chunk in ([n] + [int(a) << 20 for a in sys.argv[1:]]) if len(sys.argv) > 1 else [n, 64 << 20, 16 << 20, 4 << 20]:
    print("chunk %5d MB: H2D alone %.1f  D2H alone %.1f  both at once %.1f GB/s each" % (chunk >> 20, run(True, False, chunk), run(False, True, chunk), run(True, True, chunk)), flush=True)
