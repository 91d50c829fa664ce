#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX -- text summaries of one `ncu --set full --import-source on` capture:
OUT_PREFIX_summary.txt (headline metrics + stall reasons) and OUT_PREFIX_hotspots.txt (instruction share / stall
samples per source line).  Needs the `ncu` CLI (it only reads the report; no GPU)."""
import collections, csv, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
m = dict(zip(rows[0], rows[2]))
units = dict(zip(rows[0], rows[1]))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__icc_request_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
with open(out + "_summary.txt", "w") as f:
    f.write(f"# {rep}: kernel {m.get('Kernel Name', '?')}\n")
    for k in keys:
        if k in m:
            f.write(f"{k:75s} {m[k]:>18s} {units.get(k, '')}\n")
    f.write("--- stalls (warps per issue active)\n")
    for k in sorted(m):
        if "warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try:
                v = float(m[k])
            except ValueError:
                continue
            if v >= 0.05:
                f.write(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}\n")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
cols = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
lines = []
for r in rows:
    if len(r) <= iI or r[0] in ("", "Line No"):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = {k: int(r[c]) if r[c].isdigit() else 0 for k, c in cols.items()}
    lines.append((ln, r[1].strip(), int(r[iS]) if r[iS].isdigit() else 0, int(r[iI]) if r[iI].isdigit() else 0, d))
S = sum(l[2] for l in lines) or 1
I = sum(l[3] for l in lines) or 1
with open(out + "_hotspots.txt", "w") as f:
    f.write(f"# {rep}: {I} warp-instructions, {S} stall samples; per source line of zj_kernels.cu (inlined callees count at their own lines)\n")
    tot = collections.Counter()
    for l in lines:
        for k, v in l[4].items():
            tot[k] += v
    f.write("stall samples by reason: " + ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in tot.most_common(10)) + "\n")
    f.write("--- by instruction share\n")
    for l in sorted(lines, key=lambda x: -x[3])[:40]:
        f.write(f"{l[0]:5d} inst {100 * l[3] / I:5.2f}% samples {100 * l[2] / S:5.2f}% | {l[1][:120]}\n")
    f.write("--- by stall samples\n")
    for l in sorted(lines, key=lambda x: -x[2])[:25]:
        top = sorted(l[4].items(), key=lambda kv: -kv[1])[:2]
        f.write(f"{l[0]:5d} samples {100 * l[2] / S:5.2f}% inst {100 * l[3] / I:5.2f}% {[(k[6:], v) for k, v in top]} | {l[1][:100]}\n")
print("wrote", out + "_summary.txt", out + "_hotspots.txt")
