#!/usr/bin/env python
"""Step-by-step comparison of a 2-slab fast-mode run with the single-context run from step START on (debug aid)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
pkg = graft.load_package()
from sph_b200 import scenes
capi = pkg.capi
start, stop = int(sys.argv[1]), int(sys.argv[2])
G = 2
family, dx = scenes.SCENES["dam_break_10M"]
dx = scenes.dam_break_dx_for(G * scenes.dam_break_count(dx), dx / G ** (1.0 / 3.0))
pos, mass, params, dt = scenes.dam_break_scene(dx)
n = len(pos)
m = pkg.MultiContext(n, [0, 0]); m.set_option(capi.OPT_GRID_REFINE, 4); m.set_params(params); m.upload(pos, None, mass)
one = pkg.Context(n, 0); one.set_option(capi.OPT_GRID_REFINE, 4); one.set_option(capi.OPT_LAYOUT_MAJOR, 2); one.set_params(params); one.upload(pos, None, mass)
nsr = float(params["neighbor_search_radius"])
for k in range(1, stop + 1):
    m.step(dt); one.step(dt)
    if k < start:
        continue
    a, b = m.download(), one.download()
    bad = {f: np.flatnonzero((a[f].view(np.uint32).reshape(n, -1) != b[f].view(np.uint32).reshape(n, -1)).any(1)) for f in ("rho", "P", "acc", "pos", "vel")}
    print(f"step {k}: owned {m.layout()['owned'].tolist()} differing " + " ".join(f"{f} {v.size}" for f, v in bad.items()), flush=True)
    ids = np.unique(np.concatenate(list(bad.values())))
    if ids.size:
        for i in ids[:12]:
            print(f"   id {i}: pos/nsr {np.round(b['pos'][i] / nsr, 4).tolist()} mass {mass[i]:.3e} rho {a['rho'][i]!r} {b['rho'][i]!r} acc {a['acc'][i].tolist()} {b['acc'][i].tolist()}")
        # everything within 2.01 nsr below / above the cut that sits within 0.003 nsr past a layer boundary
        z = b["pos"][:, 2] / nsr
        for lo, hi in ((-2.003, -2.0), (2.0, 2.003), (-1.003, -1.0), (1.0, 1.003)):
            print(f"   particles with z/nsr in [{lo}, {hi}): {int(((z >= lo) & (z < hi)).sum())}")
        break
