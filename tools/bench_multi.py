#!/usr/bin/env python
"""Side measurement (not the driver's bench line): ONE engine over G GPUs inside ONE process — sphb_create_multi behind
the C ABI, peer-to-peer copies instead of NCCL (csrc/multi.cu) — on the weak-scaled dam break of bench.py.

    python tools/bench_multi.py --gpus 2 [--scene dam_break_10M] [--steps 30] [--warmup 60] [--check]

Timed region: `steps` calls of sphb_multi_step bracketed by sphb_multi_synchronize (all devices) and a host clock —
several devices have no common CUDA event, and the call under test is a host call that drives all of them.
--check: after the timed region the same scene is stepped on one device and compared (owned counts, sum rho, KE)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--devices", default=None, help="comma-separated ordinals (default 0..gpus-1; repeat an ordinal for slabs on one GPU)")
    ap.add_argument("--scene", default="dam_break_10M")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    pkg = graft.load_package()
    from sph_b200 import scenes
    capi = pkg.capi
    devices = [int(d) for d in args.devices.split(",")] if args.devices else list(range(args.gpus))
    G = len(devices)
    family, dx = scenes.SCENES[args.scene]
    if G > 1 and args.scaling == "weak":
        dx = scenes.dam_break_dx_for(G * scenes.dam_break_count(dx), dx / G ** (1.0 / 3.0)) if family == "dam" else dx / G ** (1.0 / 3.0)
    pos, mass, params, dt = scenes.dam_break_scene(dx) if family == "dam" else scenes.fluid_drop_scene(dx)
    n = pos.shape[0]
    refine = int(max(1, min(6, round(float(params["neighbor_search_radius"]) / dx))))
    m = pkg.MultiContext(n, devices)
    m.set_option(capi.OPT_GRID_REFINE, refine)
    m.set_params(params)
    t0 = time.perf_counter()
    m.upload(pos, None, mass)
    m.synchronize()
    upload_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        m.step(dt)
    m.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.step(dt)
    m.synchronize()
    secs = time.perf_counter() - t0
    lay = m.layout()
    out = {"metric": "particle-updates/sec (M/s)", "value": n * args.steps / secs / 1e6, "unit": "M particle-updates/s", "n_gpus": G,
           "devices": devices, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
           "scaling": args.scaling, "timing": "host clock around sphb_multi_step x steps, sphb_multi_synchronize on both sides",
           "config": {"workload": args.scene, "particles": int(n), "dx": dx, "engine": "sphb_create_multi (one process, peer copies)",
                      "slab_axis": int(lay["axis"]), "owned": [int(x) for x in lay["owned"]], "halo_copies": [int(x) for x in lay["ghosts"]]},
           "upload_s": upload_s, "max_neighbors": int(m.stats()["max_neighbors"]), "error_flags": int(m.stats()["error_flags"])}
    if args.check:
        sr, ke, vmax = m.diagnostics()
        total = args.warmup + args.steps
        one = pkg.Context(n, devices[0])
        one.set_option(capi.OPT_GRID_REFINE, refine)
        one.set_option(capi.OPT_LAYOUT_MAJOR, int(lay["axis"]))
        one.set_params(params)
        one.upload(pos, None, mass)
        for _ in range(total):
            one.step(dt)
        sr1, ke1, vmax1 = one.diagnostics()
        a, b = m.download(pos=True, vel=False, rho=True, pressure=False, acc=False), one.download(pos=True, vel=False, rho=True, pressure=False, acc=False)
        one.close()
        out["validation"] = {"steps": total, "owned_total": int(lay["owned"].sum()), "sum_rho_rel": abs(sr - sr1) / abs(sr1),
                             "ke_rel": abs(ke - ke1) / max(abs(ke1), 1e-300), "max_speed_equal": bool(vmax == vmax1),
                             "pos_bit_equal": bool(np.array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32))),
                             "rho_bit_equal": bool(np.array_equal(a["rho"].view(np.uint32), b["rho"].view(np.uint32)))}
    m.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
