#!/usr/bin/env python
"""Per-instruction stall picture of a SASS line range from `ncu --page source --csv`.
usage: ncu_sass_lines.py file.csv first last [kernel_index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ends = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
ki = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = rows[ends[ki]: ends[ki + 1]]
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
a, b = int(sys.argv[2]), int(sys.argv[3])
tot = sum(int(r[ismp]) for r in rows[2:] if len(r) > ismp)
for n, r in enumerate(rows[2:]):
    if n < a or n > b or len(r) <= ismp: continue
    s = sorted(((int(r[k] or 0), h[6:]) for k, h in stall), reverse=True)[:3]
    print(f"{n:4d} {int(r[iex])/1e6:6.2f}M smp={int(r[ismp]):5d} ({100*int(r[ismp])/tot:4.1f}%) {r[isrc].strip():70s} " + " ".join(f"{h}={v}" for v, h in s if v))
