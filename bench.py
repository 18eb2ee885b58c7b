#!/usr/bin/env python
"""bench.py — particle-updates/s of the SPH per-step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scene NAME] [--impl ours|reference]

A "step" is one call of the hot path (SPHEngine::step: neighbour build, density + EOS, force,
integrate + clamp) over every particle of the scene.  Default workload: BASELINE.json configs[1],
the 1M-particle dam break (N = 1 130 000 from the reference's own lattice generator at dx = 0.004)
with the bounded "P-tame" parameter set (the reference's defaults explode in one step).

One JSON line on stdout (rank 0).  `value` = particles x steps / device time with the state resident in
HBM; `e2e` = the same metric through the C ABI with HOST buffers (pinned H2D upload + step + D2H
download every step); `roofline` = the dominant kernel against the measured HBM peak; `cpu_baseline`
= the reference's own CPU/OpenMP code (oracle/_ref, compiled from the unmodified reference sources
with its own flags) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

METRIC = "particle_updates_per_sec"
UNIT = "M particle-updates/s"
B_ALG_STEP = 276.0     # algorithmic HBM bytes per particle-step
B_ALG = {"density": 24.0, "force": 48.0, "integrate": 60.0, "neighbor": 144.0}


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU code on the host cores
# ------------------------------------------------------------------------------------------------
def _use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU reference must use the box's cores.
    Has to run before libgomp initialises (i.e. before the oracle library is loaded)."""
    if "TORCHELASTIC_RUN_ID" in os.environ or os.environ.get("OMP_NUM_THREADS") in (None, "", "1"):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)


def cpu_reference_run(scene_name: str, steps: int, warmup: int):
    """Time the reference CPU implementation (oracle/_ref fast build; else the C port) on `scene_name`."""
    _use_all_host_cores()
    po = graft.load_oracle()
    pkg = graft.load_package()
    from sph_b200 import scenes
    kind, label = ("fast", "reference") if po.available("fast") else ("port", "port")
    if kind == "port" and not po.available("port"):
        subprocess.run(["make", "-s", "-f", str(ROOT / "oracle" / "Makefile"), "port"], check=True)
    pos, mass, params, dt = scenes.make_scene(scene_name)
    n = pos.shape[0]
    eng = po.Engine(kind, n)
    eng.initialize(params)
    eng.add_particles(pos, None, mass)
    cores = eng.L.num_threads()
    if warmup > 0:
        eng.timed_steps(warmup, dt)
    secs = eng.timed_steps(steps, dt)
    eng.close()
    return dict(value=n * steps / secs / 1e6, secs=secs, n=n, cores=cores, kind=label,
                sample=f"{scene_name} (N={n}), {warmup} warm-up + {steps} timed steps, fixed dt={dt:.3g}, OMP threads={cores}")


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # bounded sample of the same workload family: the reference's throughput is flat in N,
    # so a 150k-particle dam break at the same h/dx keeps every step ~1 s on 8 cores
    sample_scene = args.cpu_scene or "dam_break_150k"
    r = cpu_reference_run(sample_scene, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 6), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * r["secs"] / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.scene, "sample": r["sample"], "note": "reference CPU/OpenMP path on host cores; "
                   "bounded sample of the workload family (throughput is flat in N for the reference)"},
        "cpu_baseline": {"value": round(r["value"], 6), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": round(r["value"], 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def scaled_scene(scenes, name: str, world: int, mode: str):
    """N = 1: the named scene.  N > 1, weak scaling: the same geometry with dx shrunk so that the scene holds N times
    the particles of the single-GPU scene (per-GPU work fixed); strong scaling: the named scene split N ways."""
    family, dx = scenes.SCENES[name]
    if world > 1 and mode == "weak":
        if family == "dam":
            dx = scenes.dam_break_dx_for(world * scenes.dam_break_count(dx), dx / world ** (1.0 / 3.0))
        else:
            dx = dx / world ** (1.0 / 3.0)
    pos, mass, params, dt = scenes.dam_break_scene(dx) if family == "dam" else scenes.fluid_drop_scene(dx)
    return pos, mass, params, dt, dx


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    pkg = graft.load_package()
    from sph_b200 import scenes, slab
    capi = pkg.capi

    pos, mass, params, dt, dx = scaled_scene(scenes, args.scene, world, args.scaling)
    n_total = pos.shape[0]
    stream = torch.cuda.current_stream(dev)
    strict = args.math == "strict"
    if args.refine <= 0:
        # scene-level tuning: internal cells of about one lattice spacing (~1 particle per cell): nsr / dx = 4 for the
        # dam-break scenes (h = 2 dx), 5 for the fluid drop (h = 2.5 dx)
        args.refine = int(max(1, min(6, round(float(params["neighbor_search_radius"]) / dx))))
    opts = {capi.OPT_PAIR_KERNEL: args.pair_kernel, capi.OPT_GRID_REFINE: args.refine}

    if world == 1:
        n_local = n_total
        ctx = pkg.Context(n_total, local)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT if strict else capi.MATH_FAST)
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_params(params)
        mine = None
        parallelism = "single"
    else:
        nsr = float(params["neighbor_search_radius"])
        axis = 2
        cells = slab.axis_cells(pos, axis, nsr)
        cuts = slab.plan_cuts(cells, world, 2)
        owner = slab.rank_of_cells(cuts, cells)
        n_local = int((owner == rank).sum())
        lay = int(((cells >= cuts[rank + 1]) & (cells < cuts[rank + 1] + 2)).sum() + ((cells < cuts[rank]) & (cells >= cuts[rank] - 2)).sum())
        cap = int(1.25 * n_local + 2 * lay + 65536)
        store = slab.GpuStore(pkg, cap, local, params, strict=strict, stream=stream.cuda_stream, options=opts)
        box_min = np.minimum(pos.min(0), [params["xmin"], params["ymin"], params["zmin"]])
        box_max = np.maximum(pos.max(0), [params["xmax"], params["ymax"], params["zmax"]])
        sr = slab.SlabRank(store, rank, cuts, axis, 2, n_total, box_min, box_max, max(4 * lay + 65536, n_local // 2))
        ctx = store.ctx
        mine = np.flatnonzero(owner == rank)
        parallelism = f"slab{world}(z), 2-layer halo, NCCL all_to_all migration + isend/irecv halo"

    # pinned host buffers (the user's arrays at the API boundary)
    sel = slice(None) if mine is None else mine
    h_pos = torch.from_numpy(np.ascontiguousarray(pos[sel])).pin_memory()
    h_mass = torch.from_numpy(np.ascontiguousarray(mass[sel])).pin_memory()
    h_vel = torch.zeros((n_local, 3), dtype=torch.float32).pin_memory()
    h_ids = None if mine is None else torch.from_numpy(mine.astype(np.int32)).pin_memory()
    o_cap = n_local if world == 1 else store.capacity
    o_pos = torch.empty((o_cap, 3), dtype=torch.float32).pin_memory()
    o_vel = torch.empty((o_cap, 3), dtype=torch.float32).pin_memory()
    o_rho = torch.empty((o_cap,), dtype=torch.float32).pin_memory()
    o_ids = torch.empty((o_cap,), dtype=torch.int32).pin_memory()

    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    import ctypes

    def upload():
        if world == 1:
            ctx.upload_raw(n_local, h_pos.data_ptr(), h_vel.data_ptr(), h_mass.data_ptr())
        else:
            ctx._ck(ctx.L.sphb_upload_ids(ctx.h, n_local, ctypes.c_void_p(h_pos.data_ptr()), ctypes.c_void_p(h_vel.data_ptr()),
                                          ctypes.c_void_p(h_mass.data_ptr()), ctypes.c_void_p(h_ids.data_ptr())))

    def step():
        if world == 1:
            ctx.step(dt)
        else:
            slab.step_distributed(sr, dt)

    def download():
        if world == 1:
            ctx.download_raw(o_pos.data_ptr(), o_vel.data_ptr(), o_rho.data_ptr(), None, None)
            return n_local
        cnt = ctypes.c_size_t()
        ctx._ck(ctx.L.sphb_slab_download(ctx.h, o_cap, ctypes.c_void_p(o_ids.data_ptr()), ctypes.c_void_p(o_pos.data_ptr()),
                                         ctypes.c_void_p(o_vel.data_ptr()), ctypes.c_void_p(o_rho.data_ptr()), None, None,
                                         ctypes.byref(cnt)))
        return cnt.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -------------------------------------------------------------
    upload()
    for _ in range(args.warmup):
        step()
    ctx.set_option(capi.OPT_STAGE_TIMING, 1)
    ctx.reset_stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # evict L2 between timed steps (outside the event bracket)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    st = ctx.stats()
    ctx.set_option(capi.OPT_STAGE_TIMING, 0)
    launches = int(st["kernel_launches"])

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n_total * args.steps / (total_ms_max * 1e-3) / 1e6

    # ---- end-to-end through the host-buffer API --------------------------------------------------
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        upload(); step(); download()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        upload()
        step()
        download()                         # synchronises
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(t.item()) / 1e6
    h2d = n_local * (12 + 12 + 4 + (4 if world > 1 else 0))
    d2h = n_local * (12 + 12 + 4 + (4 if world > 1 else 0))

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage times from CUDA events inside the timed region) ---
    peak, peak_src = read_peaks()
    stages = {"neighbor": st["neighbor_search_time"], "density": st["density_computation_time"],
              "force": st["force_computation_time"], "integrate": st["integration_time"]}
    dom = max(("density", "force"), key=lambda k: stages[k])
    dom_s = stages[dom] / max(1, st["steps"])
    n_kernel = n_local     # particles the profiled kernel (rank 0's) processes per launch
    achieved = B_ALG[dom] * n_kernel / dom_s / 1e9 if dom_s > 0 else 0.0
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            rec = json.loads(tp.read_text()).get(args.scene, {}).get(f"k_{dom}")
            if rec and world == 1:
                traffic = rec["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"k_{dom}", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": B_ALG[dom], "kernel_ms": round(1e3 * dom_s, 4),
                "note": "pair kernels are instruction-issue bound at h = 2 dx (hundreds of candidate pairs per particle; ncu: 85 % / 74 % "
                        "issue-active, 6-9 % DRAM utilisation, profiles/r1d_pair_kernels_mask_full.md); traffic exceeds the algorithmic "
                        "bytes on purpose (neighbour bitmasks handed from the density to the force pass); HBM fraction reported as mandated"}
    step_s = total_ms_max * 1e-3 / args.steps
    extra = {
        "step_hbm_frac": round(B_ALG_STEP * n_total / world / step_s / 1e9 / peak, 5),
        "stage_ms_rank0": {k: round(1e3 * v / max(1, st["steps"]), 4) for k, v in stages.items()},
        "max_neighbors": int(st["max_neighbors"]),
        "ms_per_step_min": round(min(step_ms), 4), "ms_per_step_max": round(max(step_ms), 4),
        "wall_ms_per_step_incl_l2_flush": round(1e3 * wall / args.steps, 3),
    }
    if world > 1:
        extra["slab"] = {"cuts": [int(c) for c in cuts[1:-1]], "owned_rank0": n_local, **sr.stats}

    # ---- CPU baseline: the reference's own code on this box's host cores, bounded sample ---------
    cpu = None
    if not args.no_cpu and world == 1:      # contract: CPU baseline on rank 0 at N = 1 only
        try:
            r = cpu_reference_run(args.cpu_scene or "dam_break_347k", 3, 1)
            cpu = {"value": round(r["value"], 6), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total_ms_max / args.steps, 4), "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.scene, "dx": dx, "particles_total": n_total, "particles_rank0": n_local, "h_over_dx": 2.0,
                   "params": "P-tame", "dt": dt, "math": args.math, "pair_kernel": args.pair_kernel, "grid_refine": args.refine,
                   "parallelism": parallelism, "l2": "flushed between timed steps (512 MiB write outside the event bracket)"},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "what": "sphb_upload(host pos,vel,mass) + sphb_step + sphb_download(host pos,vel,rho) per step"
                                            + (" per rank, slab exchange included" if world > 1 else "")},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="dam_break_1M")
    ap.add_argument("--cpu-scene", default=None)
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--pair-kernel", type=int, default=2)
    ap.add_argument("--refine", type=int, default=0, help="internal grid refine 1..6; 0 = neighbor_search_radius / dx of the scene")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
