#!/usr/bin/env python
"""Coarse profile of a kernel along its SASS: per bin of N instructions, executed warp-instructions, stall samples, FP64 share.
usage: ncu_bins.py report.ncu-rep [bin=250]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 250
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0].startswith("0x")]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
cols = ["stall_no_inst", "stall_wait", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_branch_resolving"]
ci = [h.index(c) for c in cols]
print("bin   first  exec(M)  fp64%  samples  " + "  ".join(c[6:13] for c in cols))
for b in range(0, len(body), N):
    chunk = body[b:b + N]
    ex = sum(float(r[ie] or 0) for r in chunk)
    fp = sum(float(r[ie] or 0) for r in chunk if re.match(r"\s*(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)", r[1]))
    sm = sum(float(r[isamp] or 0) for r in chunk)
    st = [sum(float(r[k] or 0) for r in chunk) for k in ci]
    print(f"{b // N:3d} {b:6d} {ex / 1e6:8.1f} {100 * fp / max(ex, 1):6.1f} {sm:8.0f}  " + "  ".join(f"{x:7.0f}" for x in st))
