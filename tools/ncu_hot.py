#!/usr/bin/env python
"""Top stall-sample SASS lines of an ncu report: python tools/ncu_hot.py report.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; body = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ci['# Samples']]):6d} {100*int(r[ci['# Samples']])/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {st}")
