#!/usr/bin/env python
"""Per-address-bucket stall profile of one captured kernel: where in the SASS do the no-instruction / wait / long-scoreboard
samples sit.   python tools/ncu_stalls.py report.ncu-rep [--index 0] [--bucket 64]"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("--index", type=int, default=0); ap.add_argument("--bucket", type=int, default=64)
a = ap.parse_args()
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
s = starts[a.index]
hdr = rows[s]
print(rows[s - 1][1])
data = []
for r in rows[s + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    data.append(r)
col = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_barrier", "stall_selected", "stall_not_selected")}
tot = {k: sum(int(r[c]) for r in data) for k, c in col.items()}
print("instructions", len(data), {k: v for k, v in tot.items()})
print("bucket  first_instr                                   inst%  samp%  no_inst% wait% long% short% math% thr/inst")
for b in range(0, len(data), a.bucket):
    ch = data[b:b + a.bucket]
    g = {k: sum(int(r[c]) for r in ch) for k, c in col.items()}
    if g["# Samples"] * 200 < tot["# Samples"] and g["Instructions Executed"] * 200 < tot["Instructions Executed"]: continue
    print(f"{b:5d}  {ch[0][1].strip()[:44]:44s} {100*g['Instructions Executed']/tot['Instructions Executed']:5.1f} {100*g['# Samples']/tot['# Samples']:6.1f} "
          f"{100*g['stall_no_inst']/max(1,tot['stall_no_inst']):7.1f} {100*g['stall_wait']/max(1,tot['stall_wait']):6.1f} {100*g['stall_long_sb']/max(1,tot['stall_long_sb']):5.1f} "
          f"{100*g['stall_short_sb']/max(1,tot['stall_short_sb']):5.1f} {100*g['stall_math']/max(1,tot['stall_math']):5.1f} {g['Thread Instructions Executed']/max(1,g['Instructions Executed']):6.1f}")
