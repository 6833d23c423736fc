#!/usr/bin/env python
"""Executed warp-instructions and stall samples of an ncu report by SOURCE LINE (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [units]   (units: divide the executed counts by this number, e.g. warp-steps of the launch)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur_file = "?"
agg = collections.OrderedDict()
hdr = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":          # a source line row (aggregated)
        key = (cur_file, int(r[0]))
        agg[key] = [r[1].strip(), float(r[7] or 0), float(r[6] or 0), 0.0]
        last = key
    else:                   # a SASS row under the last source line
        op = r[3].strip().split()
        name = next((t for t in op if not t.startswith("@")), "")
        if name.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H")):
            agg[last][3] += float(r[7] or 0)
tot = sum(v[1] for v in agg.values())
tots = sum(v[2] for v in agg.values())
print(f"total executed {tot:.0f} ({tot / units:.1f} per unit), samples {tots:.0f}")
print("   exec/unit  fp64/unit  share  samp%   file:line  source")
for (f, ln), (src, ex, sm, fp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"   {ex / units:8.1f}  {fp / units:8.1f}  {100 * ex / tot:5.1f}  {100 * sm / max(tots, 1):5.1f}   {f}:{ln}  {src[:110]}")
