#!/usr/bin/env python
"""Turn raw ncu outputs (gpurun_out/) into the small text summaries committed under profiles/.

  ncu_summary.py launches <launches.csv> <out.md> [title]
  ncu_summary.py full <file.ncu-rep> <out.md> [title]
"""
import collections, csv, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip(), 1e-3)
        name = r[ki].replace("sphb::<unnamed>::", "").replace("void ", "")
        name = name.split("(")[0]
        agg.setdefault(name, []).append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and\nserialised: compare SHARES, not absolutes).  {sum(len(v) for v in agg.values())} launches, {tot/1e3:.3f} ms total.\n\n")
        f.write("| kernel | launches | avg µs | total µs | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |\n")
    print("wrote", out)


def full(path, out, title):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({path.split('/')[-1]}); one column per captured launch.\n\n")
        names = [r[hdr.index("Kernel Name")].replace("sphb::<unnamed>::", "").split("(")[0].replace("void ", "") for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|---|" + "---:|" * len(names) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")
    print("wrote", out)


if __name__ == "__main__":
    mode, src, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, out, title)
