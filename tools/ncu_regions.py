#!/usr/bin/env python
"""Region view of an ncu source page: consecutive SASS instructions with (nearly) equal executed counts are merged into
regions; prints executed warp-instructions, stall samples and the dominant stall reasons per region.
usage: ncu_regions.py src.csv (made by: ncu -i rep --page source --csv --print-source sass) [units] [min_share]
units = number of 32-cell group-steps in the launch (to print instructions per group-step)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0].startswith("0x")]
ix = {n: i for i, n in enumerate(h)}
E = [float(r[ix["Instructions Executed"]] or 0) for r in body]
S = [float(r[ix["# Samples"]] or 0) for r in body]
reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
totE, totS = sum(E), sum(S)
regions = []
start = 0
for i in range(1, len(body) + 1):
    txt = body[i - 1][1]
    brk = i == len(body) or abs(E[i] - E[start]) > 0.02 * max(E[start], 1.0) or "BAR" in txt or "CALL" in txt or "RET" in txt
    if brk:
        regions.append((start, i))
        start = i
print(f"{len(body)} instr, {totE:.0f} executed ({totE / units:.0f} per unit), {totS:.0f} samples")
print(f"{'idx':>6} {'len':>5} {'exec/instr/unit':>10} {'instr/unit':>10} {'share':>6} {'samp%':>6} {'fp64':>5}  stalls  | first instruction")
for a, b in regions:
    e = sum(E[a:b]); s = sum(S[a:b])
    if e < min_share * totE and s < min_share * totS:
        continue
    fp64 = sum(E[i] for i in range(a, b) if any(op in body[i][1] for op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")))
    st = {n: sum(float(body[i][ix[n]] or 0) for i in range(a, b)) for n in reasons}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    tops = " ".join(f"{n[6:]}:{100 * v / max(s, 1):.0f}" for n, v in top)
    print(f"{a:6d} {b - a:5d} {E[a] / units:10.2f} {e / units:10.1f} {100 * e / totE:6.2f} {100 * s / totS:6.2f} {100 * fp64 / max(e, 1):5.0f}  {tops:38s} | {body[a][1].strip()[:60]}")
