#!/usr/bin/env python
"""tools/update_traffic.py TAG -- fill profiles/traffic.json from gpurun_out/TAG_ncu_full_<config>_summary.txt (written by
tools/profile_round.sh): dram__bytes_read.sum + dram__bytes_write.sum of one launch per config, stamped with the hash of the
kernel sources in the tree (bench.py reports the figure only while that hash matches)."""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
path = os.path.join(ROOT, "profiles", "traffic.json")
tr = json.load(open(path))
sha = bench.kernel_sha16()
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
for cfg in ("c1", "c2", "c3_444", "c3_gray", "c4", "c5"):
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_full_{cfg}_summary.txt")
    if not os.path.exists(src):
        continue
    txt = open(src).read()
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(key) + r"\s+([0-9.]+)\s+(\w+)", txt)
        tot += float(m.group(1)) * UNIT[m.group(2)]
    for suffix in ("summary", "hotspots", "mix"):
        f = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_full_{cfg}_{suffix}.txt")
        if os.path.exists(f):
            shutil.copy(f, os.path.join(ROOT, "profiles", os.path.basename(f)))
    tr[cfg] = {"dram_bytes_per_launch": int(round(tot)), "kernel_sha16": sha, "source": f"profiles/{tag}_ncu_full_{cfg}_summary.txt"}
    print(cfg, int(round(tot)), sha)
json.dump(tr, open(path, "w"), indent=1)
