#!/usr/bin/env python
"""tools/sass_mix.py REPORT.ncu-rep [OUT.txt] -- executed-instruction mix of one `ncu --set full --import-source on` capture:
warp-instructions per SASS opcode and per issue pipe (fma: IMAD/IDP/VIADD..., alu: IADD3/LOP3/SHF/PRMT/VIMNMX/I2IP...,
lsu: LDS/STS/LDG/STG/LDGSTS..., other), plus the same split per code region when region boundaries (SASS line numbers of the
source page) are given with --regions name:first-last,...   Needs the `ncu` CLI (reads the report; no GPU)."""
import collections, csv, subprocess, sys

FMA = ("IMAD", "IDP", "VIADD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL")
ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "VIMNMX", "I2IP", "LEA", "ISETP", "SEL", "MOV", "SHL", "SHR", "VABSDIFF", "IABS", "FMNMX", "PLOP3", "FSETP", "ICMP", "SGXT", "BMSK", "FLO", "POPC", "LOP", "I2I", "VIMNMX3")
LSU = ("LDS", "STS", "LDG", "STG", "LDGSTS", "LD", "ST", "LDC", "ATOMS", "ATOMG", "RED", "LDSM", "LDGDEPBAR", "DEPBAR", "SHFL", "MATCH", "VOTE", "CCTL", "ULDC")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA: return "fma"
    if base in ALU: return "alu"
    if base in LSU: return "lsu/mio"
    return "other"


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else sys.stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    iS, iI, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ops, pipes = collections.Counter(), collections.Counter()
    total = 0
    per_line = []
    for r in rows[hi + 1:]:
        if len(r) <= iT: continue
        txt = r[iS].strip()
        toks = txt.split()
        if not toks: continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        n = int(r[iI] or 0)
        ops[op.split(".")[0]] += n
        pipes[pipe(op)] += n
        total += n
        per_line.append((n, txt))
    out.write(f"# {rep}: {rows[0][1] if rows and len(rows[0]) > 1 else ''}\n# {total} warp-instructions executed, {len(per_line)} SASS lines\n")
    out.write("--- per pipe\n")
    for k, v in pipes.most_common():
        out.write(f"  {k:10s} {v:14d} {100.0 * v / total:6.2f}%\n")
    out.write("--- per opcode\n")
    for k, v in ops.most_common(40):
        out.write(f"  {k:10s} {pipe(k):8s} {v:14d} {100.0 * v / total:6.2f}%\n")


if __name__ == "__main__":
    main()
