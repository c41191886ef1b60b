#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.
usage: summarize_launches.py <launches.csv>"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("fb::", "").strip()
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'us/launch':>10s}")
for name, us in tot.most_common():
    print(f"{name[:70]:70s} {cnt[name]:8d} {us:12.1f} {100 * us / total:6.1f}% {us / cnt[name]:10.1f}")
