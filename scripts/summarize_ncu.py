#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full ... --import-source on) into a small text file
for profiles/: per kernel the roofline-relevant counters, top stall reasons and the
executed-opcode mix.   usage: summarize_ncu.py <report.ncu-rep> [kernel-regex]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def run(page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = run("raw")
hdr, units = raw[0], raw[1]
ki = hdr.index("Kernel Name")
print(f"# {rep}")
for r in raw[2:]:
    if pat and not pat.search(r[ki]):
        continue
    print(f"\n== {r[ki][:110]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i]:>18s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try:
                stalls.append((float(r[i]), m.group(1)))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  top stalls (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))

src = run("source")
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
seen = set()
for k, st in enumerate(starts):
    name = src[st][1]
    if (pat and not pat.search(name)) or name in seen:
        continue
    seen.add(name)
    seg = src[st + 1: starts[k + 1]] if k + 1 < len(starts) else src[st + 1:]
    h = seg[0]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    iw, ii = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
    agg = collections.Counter()
    wf = wfi = 0
    for r in seg[1:]:
        if len(r) > ie and r[ie].isdigit():
            op = r[ia].strip().split()
            o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
            agg[o] += int(r[ie])
            if r[iw].isdigit():
                wf += int(r[iw]); wfi += int(r[ii]) if r[ii].isdigit() else 0
    tot = sum(agg.values())
    print(f"\n== executed warp-instructions by opcode: {name[:90]}  (total {tot})")
    print("  " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in agg.most_common(14)))
    if wf:
        print(f"  shared-memory wavefronts {wf} vs ideal {wfi}  ({wf / max(1, wfi):.2f}x)")
