#!/usr/bin/env python
"""tools/ncu_top_lines.py <source-page.csv> [n]: the n most-sampled SASS lines of the first kernel with their top stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ends = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(ends) > 1:
    rows = rows[: ends[1]]
hdr = rows[1]
isrc, iex, ismp, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) > max(stall)]
tot = sum(int(r[ismp]) for r in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top = sorted(range(len(data)), key=lambda k: -int(data[k][ismp]))[:n]
for k in sorted(top):
    r = data[k]
    st = sorted(((int(r[j]), hdr[j][6:]) for j in stall if r[j] and int(r[j]) > 0), reverse=True)[:3]
    print(k, r[isrc].strip()[:64], r[iex], "lanes", round(int(r[ith]) / max(int(r[iex]), 1), 1), "smp%", round(100 * int(r[ismp]) / tot, 2), st)
