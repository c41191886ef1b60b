#!/usr/bin/env python
"""Per-SASS-instruction view of an .ncu-rep source page: executed count, stall samples, top stall reason.
usage: ncu_hot.py <rep> [min_exec_fraction]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
body = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name":
            break
        continue
    body.append(r)
tot_exec = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"total warp-instr {tot_exec}, samples {tot_samp}")
for n, r in enumerate(body):
    ex = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    wf = r[ix["L1 Wavefronts Shared"]]
    print(f"{n:5d} {ex:10d} {100*ex/tot_exec:5.2f}% samp {sm:6d} {100*sm/max(1,tot_samp):5.2f}%  wf {wf:>9s}  {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {r[ix['Source']].strip()}")
