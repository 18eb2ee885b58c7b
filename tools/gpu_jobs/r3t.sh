# round 2 (session 3), job t: stage times of small scenes with 1 / 4 lanes per particle (stage timing on: direct launches)
set -x
python - <<'PY'
import time, json, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from sph_b200 import scenes
capi = pkg.capi
for name, dx in (("dam_break_13k", 0.02), ("dam_break_85k", 0.0105)):
    pos, mass, prm, dt = scenes.dam_break_scene(dx)
    for lanes, mode, graphs in ((1, 2, 0), (1, 0, 0), (4, 0, 0), (8, 0, 0), (1, 2, 1), (4, 0, 1)):
        ctx = pkg.Context(len(pos), 0)
        ctx.set_option(capi.OPT_LANES_PER_PARTICLE, lanes)
        ctx.set_option(capi.OPT_PAIR_MODE, mode)
        ctx.set_option(capi.OPT_STEP_GRAPHS, graphs)
        ctx.set_option(capi.OPT_GRID_REFINE, 4)
        ctx.set_params(prm)
        ctx.upload(pos, None, mass)
        for _ in range(68):
            ctx.step(dt)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            ctx.step(dt)
        ctx.synchronize()
        wall = (time.perf_counter() - t0) / 300 * 1e3
        ctx.set_option(capi.OPT_STAGE_TIMING, 1)
        ctx.reset_stats()
        for _ in range(50):
            ctx.step(dt)
        ctx.synchronize()
        st = ctx.stats()
        print(json.dumps({"scene": name, "lanes": lanes, "mode": mode, "graphs": graphs, "wall_ms_per_step": round(wall, 4),
                          "stage_ms": {k: round(1e3 * st[k] / st["steps"], 4) for k in ("neighbor_search_time", "density_computation_time", "force_computation_time", "integration_time")},
                          "launches_per_step": st["kernel_launches"] / st["steps"]}))
        ctx.close()
PY
