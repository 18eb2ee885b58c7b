#!/usr/bin/env python
"""Debug aid (needs a build with -DSPHB_DEBUG_EXPORT_GHOSTS): after STEP steps of 2 slabs on one GPU, compare EVERY particle a
slab holds — halo copies included — with the single-context run: which densities differ, and where do they sit?"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
pkg = graft.load_package()
from sph_b200 import scenes, slab
capi = pkg.capi
steps = int(sys.argv[1])
G = 2
family, dx = scenes.SCENES["dam_break_10M"]
dx = scenes.dam_break_dx_for(G * scenes.dam_break_count(dx), dx / G ** (1.0 / 3.0))
pos, mass, params, dt = scenes.dam_break_scene(dx)
n = len(pos)
nsr = float(params["neighbor_search_radius"])
cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), G, 2)
box_min = np.minimum(pos.min(0), [params["xmin"], params["ymin"], params["zmin"]])
box_max = np.maximum(pos.max(0), [params["xmax"], params["ymax"], params["zmax"]])
ranks = []
for d in range(G):
    store = slab.GpuStore(pkg, int(0.6 * n), 0, params, strict=False, options={capi.OPT_GRID_REFINE: 4})
    ranks.append(slab.SlabRank(store, d, cuts, 2, 2, n, box_min, box_max, n // 8))
    ranks[-1].load_initial(pos, None, mass, nsr)
one = pkg.Context(n, 0); one.set_option(capi.OPT_GRID_REFINE, 4); one.set_option(capi.OPT_LAYOUT_MAJOR, 2); one.set_params(params); one.upload(pos, None, mass)
for k in range(steps):
    slab.step_local(ranks, dt); one.step(dt)
ref = one.download()
print("cuts", cuts.tolist())
for d, r in enumerate(ranks):
    got = r.store.ctx.slab_download()
    ids = got["ids"]
    ghost = (ids & np.uint32(0x80000000)) != 0
    gid = (ids & np.uint32(0x7FFFFFFF)).astype(np.int64)
    z = ref["pos"][gid, 2] / nsr
    for f in ("pos", "vel"):
        same = (got[f].view(np.uint32) == ref[f][gid].view(np.uint32)).all(1)
        print(f"rank {d}: {f} of {len(ids)} held particles ({int(ghost.sum())} halo copies): {int((~same).sum())} differ")
    drho = got["rho"].view(np.uint32) != ref["rho"][gid].view(np.uint32)
    dacc = (got["acc"].view(np.uint32) != ref["acc"][gid].view(np.uint32)).any(1)
    for name, sel in (("owned", ~ghost), ("halo", ghost)):
        zz = z[sel & drho]
        print(f"rank {d} {name}: rho differs for {zz.size}" + (f", z/nsr in [{zz.min():.4f}, {zz.max():.4f}]" if zz.size else ""))
        if name == "halo" and zz.size:
            cellz = np.floor(z[sel & drho]).astype(int)
            print("      by reference cell:", {int(c): int((cellz == c).sum()) for c in np.unique(cellz)})
    za = z[~ghost & dacc]
    print(f"rank {d} owned: acc differs for {za.size}" + (f", z/nsr {np.round(za[:8], 4).tolist()}" if za.size else ""))
    # first-layer halo copies whose density differs: the ones that can reach owned particles
    lay1 = ghost & drho & ((z >= cuts[d] - 1) & (z < cuts[d + 1] + 1))
    for i in np.flatnonzero(lay1)[:6]:
        print(f"      layer-1 halo id {gid[i]} pos/nsr {np.round(ref['pos'][gid[i]] / nsr, 4).tolist()} rho {got['rho'][i]!r} vs {ref['rho'][gid[i]]!r}")
