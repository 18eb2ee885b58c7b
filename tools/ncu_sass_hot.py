#!/usr/bin/env python
"""Print the hottest SASS regions of a kernel from `ncu --page source --csv` output.
usage: ncu_sass_hot.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# keep the first kernel instance only
ends = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(ends) > 1:
    rows = rows[: ends[1]]
hdr = rows[1]
ia, isrc, iex, ith, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = [(r[isrc].strip(), int(r[iex]), int(r[ith]), int(r[ismp])) for r in rows[2:] if len(r) > ismp]
tot = sum(d[1] for d in data); tots = sum(d[3] for d in data)
print(f"total warp-instr {tot/1e6:.1f}M, samples {tots}, sass lines {len(data)}")
# group consecutive instructions with equal exec count into blocks
blocks = []; cur = None
for k, d in enumerate(data):
    if cur and abs(d[1] - cur["ex"]) <= 0.02 * max(cur["ex"], 1):
        cur["n"] += 1; cur["sum"] += d[1]; cur["smp"] += d[3]; cur["thr"] += d[2]; cur["end"] = k
    else:
        cur = {"start": k, "end": k, "ex": d[1], "n": 1, "sum": d[1], "smp": d[3], "thr": d[2]}; blocks.append(cur)
blocks.sort(key=lambda b: -b["sum"])
for b in blocks[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    ops = {}
    for d in data[b["start"]: b["end"] + 1]:
        op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
        op = op.split(".")[0]; ops[op] = ops.get(op, 0) + 1
    top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10])
    print(f"lines {b['start']:4d}-{b['end']:4d} n={b['n']:3d} exec/instr={b['ex']/1e6:8.2f}M share={100*b['sum']/tot:5.1f}% samples={100*b['smp']/max(tots,1):5.1f}% lanes={b['thr']/max(b['sum'],1):4.1f} | {top}")
