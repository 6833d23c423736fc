#!/usr/bin/env python
"""Top instructions of an ncu report by stall reason.  usage: ncu_hot.py report.ncu-rep [reason ...]
reasons: stall_long_sb stall_no_inst stall_barrier stall_wait stall_short_sb ... (columns of the source page)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
reasons = sys.argv[2:] or ["stall_long_sb", "stall_no_inst", "stall_barrier", "stall_wait", "stall_short_sb"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0].startswith("0x")]
tot = sum(float(r[h.index("# Samples")] or 0) for r in body)
print(f"{len(body)} instructions, {tot:.0f} samples")
for reason in reasons:
    k = h.index(reason)
    s = sum(float(r[k] or 0) for r in body)
    print(f"== {reason}: {s:.0f} samples ({100 * s / tot:.1f}% of all)")
    top = sorted(enumerate(body), key=lambda ir: -float(ir[1][k] or 0))[:12]
    for i, r in top:
        print(f"   #{i:5d} {float(r[k] or 0):7.0f}  exec {float(r[h.index('Instructions Executed')] or 0):10.0f}  {r[1].strip()[:90]}")
