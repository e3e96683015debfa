#!/usr/bin/env python
"""Top stall-sample SASS lines of one launch in an ncu report: python tools/ncu_hot.py report.ncu-rep [n_lines] [launch_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page is a concatenation of per-launch tables: ["Kernel Name", name], header row ("Address", ...), body
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
i0 = starts[which]; i1 = starts[which + 1] if which + 1 < len(starts) else len(rows)
print("launch", which, "of", len(starts), ":", rows[i0][1][:120])
hdr = rows[i0 + 1]; body = [r for r in rows[i0 + 2:i1] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = "# Samples"
tot = sum(int(r[ci[S]] or 0) for r in body)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {c: sum(int(r[ci[c]] or 0) for r in body) for c in stall_cols}
print("stall totals:", sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True)[:8])
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci[S]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ci[S]]):6d} {100*int(r[ci[S]])/max(tot,1):5.1f}%  {r[ci['Source']].strip()[:80]:80s} {st}")
