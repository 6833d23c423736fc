#!/usr/bin/env python
"""Condenses an ncu report (.ncu-rep, read here without a GPU) into the text summary kept under profiles/:
key raw metrics of the profiled launch, warp-stall breakdown, executed-instruction mix by opcode and the
distribution of stall samples over the kernel's code.  usage: ncu_summary.py report.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
for v in rows[2:]:
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print("== launch:", name)
    for i, n in enumerate(h):
        if n in KEEP:
            print(f"{n:72s} {u[i]:14s} {v[i]}")
    print("-- warp stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active)")
    st = [(float(v[i] or 0), n) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio")]
    for val, n in sorted(st, reverse=True)[:10]:
        print(f"   {n.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {val:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    h = rows[hi[0]]
    isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    byop = collections.defaultdict(lambda: [0.0, 0.0])
    n_inst = 0
    for r in rows[hi[0] + 1:]:
        if len(r) <= iex or not r[0].startswith("0x"):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = m.group(2).split(".")[0] if m else "?"
        byop[op][0] += float(r[ismp] or 0)
        byop[op][1] += float(r[iex] or 0)
        n_inst += 1
    tot_s = sum(v[0] for v in byop.values())
    tot_e = sum(v[1] for v in byop.values())
    print(f"-- SASS: {n_inst} instructions in the kernel, {tot_e:.0f} warp-instructions executed, {tot_s:.0f} stall samples")
    print("   opcode      executed    share   stall-sample share")
    for op, (s, e) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"   {op:10s} {e:12.0f}  {100 * e / tot_e:5.1f}%   {100 * s / max(tot_s, 1):5.1f}%")
