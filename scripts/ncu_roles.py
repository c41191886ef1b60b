#!/usr/bin/env python
"""Stall-sample totals per SASS address range of a kernel in an .ncu-rep (source page).
Ranges are split at USETMAXREG markers (role boundaries of the pipeline kernel).
usage: ncu_roles.py <rep>"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hi + 1:] if len(r) >= len(h)]
marks = [i for i, r in enumerate(body) if "USETMAXREG" in r[ix["Source"]]] + [len(body)]
names = ["prologue"] + [f"region{j}" for j in range(len(marks))]
lo = 0
for j, hi_ in enumerate(marks):
    seg = body[lo:hi_]
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    sm = sum(int(r[ix["# Samples"]] or 0) for r in seg)
    st = collections.Counter()
    for r in seg:
        for c in stall_cols:
            st[c[6:]] += int(r[ix[c]] or 0)
    wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in seg)
    print(f"{names[j]:9s} sass[{lo}:{hi_}] warp-instr {ex:>11d} samples {sm:>7d} smem-wavefronts {wf:>10d}  " +
          ", ".join(f"{k}={v}" for k, v in st.most_common(8)))
    lo = hi_
