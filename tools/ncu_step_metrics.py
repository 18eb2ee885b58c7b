#!/usr/bin/env python
"""tools/ncu_step_metrics.py <scene> <metrics.csv> [<scene> <metrics.csv> ...] : folds per-launch ncu metrics of ONE step
(`ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum`) into
profiles/traffic.json, the file bench.py reads roofline.traffic / roofline_issue from."""
import collections, csv, json, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ALG = {"k_density": 24.0, "k_force": 48.0}
out_path = ROOT / "profiles" / "traffic.json"
out = {"_comment": "per-launch ncu metrics of one step after 60 steps of pre-roll (tools/gpu_jobs/r2t.sh, tools/ncu_step_metrics.py): "
                   "dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum, gpu__time_duration.sum (cold, serialised). "
                   "Read by bench.py for roofline.traffic and roofline_issue when the workload matches.  The pair kernels' traffic exceeds "
                   "their algorithmic bytes on purpose: the neighbour masks travel from the density to the force pass."}
args = sys.argv[1:]
for scene, path in zip(args[0::2], args[1::2]):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui].strip().lower()
        scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "inst": 1.0, "": 1.0}.get(u, 1.0)
        launches.setdefault(r[ii], {"name": r[ki]})[r[mi]] = v * scale
    per = collections.OrderedDict()
    step = {"dram_bytes": 0.0, "warp_instructions": 0.0, "time_us_serialised": 0.0, "launches": 0}
    n = None
    for L in launches.values():
        name = L["name"].split("(")[0].split("::")[-1].split("<")[0]
        d = L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
        key = "k_density" if name.startswith("k_density") else "k_force" if name.startswith("k_force") else name
        e = per.setdefault(key, {"kernel": L["name"].split("(")[0].replace("sphb::<unnamed>::", "").replace("void ", ""), "dram_bytes_per_launch": 0.0,
                                 "warp_instructions_per_launch": 0.0, "time_us": 0.0})
        e["dram_bytes_per_launch"] += d
        e["warp_instructions_per_launch"] += L.get("smsp__inst_executed.sum", 0.0)
        e["time_us"] += L.get("gpu__time_duration.sum", 0.0)
        step["dram_bytes"] += d
        step["warp_instructions"] += L.get("smsp__inst_executed.sum", 0.0)
        step["time_us_serialised"] += L.get("gpu__time_duration.sum", 0.0)
        step["launches"] += 1
    for k, e in per.items():
        for f in ("dram_bytes_per_launch", "warp_instructions_per_launch"):
            e[f] = int(round(e[f]))
        e["time_us"] = round(e["time_us"], 1)
        e["share_of_step"] = round(e["time_us"] / step["time_us_serialised"], 4)
    step = {k: (int(round(v)) if k != "time_us_serialised" else round(v, 1)) for k, v in step.items()}
    out[scene] = {"step": step, **per}
out_path.write_text(json.dumps(out, indent=1) + "\n")
print("wrote", out_path)
