#!/usr/bin/env python
"""Where does an N-slab fast-mode run first differ from the single-context run?  (debug aid)
usage: python tools/debug/multi_bisect.py [scene] [devices] [max_steps]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
pkg = graft.load_package()
from sph_b200 import scenes
capi = pkg.capi
scene = sys.argv[1] if len(sys.argv) > 1 else "dam_break_10M"
devices = [int(d) for d in (sys.argv[2] if len(sys.argv) > 2 else "0,0").split(",")]
max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 90
G = len(devices)
family, dx = scenes.SCENES[scene]
dx = scenes.dam_break_dx_for(G * scenes.dam_break_count(dx), dx / G ** (1.0 / 3.0))
pos, mass, params, dt = scenes.dam_break_scene(dx)
n = len(pos)
refine = int(max(1, min(6, round(float(params["neighbor_search_radius"]) / dx))))
import os
m = pkg.MultiContext(n, devices); m.set_option(capi.OPT_GRID_REFINE, refine); m.set_option(capi.OPT_MULTI_HALO_LAYERS, int(os.environ.get("LAYERS", "2"))); m.set_params(params); m.upload(pos, None, mass)
one = pkg.Context(n, devices[0]); one.set_option(capi.OPT_GRID_REFINE, refine); one.set_option(capi.OPT_LAYOUT_MAJOR, 2); one.set_params(params); one.upload(pos, None, mass)
nsr = float(params["neighbor_search_radius"])
done = 0
for target in (1, 3, 10, 20, 30, 45, 60, 90):
    if target > max_steps:
        break
    while done < target:
        m.step(dt); one.step(dt); done += 1
    a, b = m.download(), one.download()
    lay = m.layout()
    bad = {}
    for f in ("rho", "acc", "pos", "vel"):
        x, y = a[f].view(np.uint32).reshape(n, -1), b[f].view(np.uint32).reshape(n, -1)
        bad[f] = np.flatnonzero((x != y).any(1))
    print(f"step {done}: cuts {lay['cuts'].tolist()} owned {lay['owned'].tolist()} differing rho {bad['rho'].size} acc {bad['acc'].size} pos {bad['pos'].size} vel {bad['vel'].size}", flush=True)
    ids = bad["rho"] if bad["rho"].size else bad["acc"]
    if ids.size:
        z = b["pos"][ids, 2] / nsr
        print("   first ids", ids[:8].tolist(), "z/nsr", np.round(z[:8], 4).tolist(), "z/nsr range", float(z.min()), float(z.max()))
        print("   rho multi/single", a["rho"][ids[:4]].tolist(), b["rho"][ids[:4]].tolist())
        print("   |d rho| max rel", float(np.abs(a["rho"][ids].astype(np.float64) - b["rho"][ids]).max() / np.abs(b["rho"]).max()))
        break
