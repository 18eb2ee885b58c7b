# round 2 (session 3), job f: CUDA-graph replay of repeating steps — parity with direct launches, small-scene step time
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graphs" 2>&1 | tail -15
python - <<'PY'
import time, json, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from sph_b200 import scenes
capi = pkg.capi
import torch
for name, dx in (("dam_break_13k", 0.02), ("dam_break_85k", 0.0105), ("dam_break_1M", None)):
    if dx is None:
        fam, dx = scenes.SCENES[name]
    pos, mass, prm, dt = scenes.dam_break_scene(dx)
    for graphs in (0, 1):
        ctx = pkg.Context(len(pos), 0)
        ctx.set_option(capi.OPT_STEP_GRAPHS, graphs)
        ctx.set_option(capi.OPT_GRID_REFINE, 4)
        ctx.set_params(prm)
        ctx.upload(pos, None, mass)
        for _ in range(60):
            ctx.step(dt)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 400
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(8):
            ctx.step(dt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            ctx.step(dt)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        print(json.dumps({"scene": name, "n": len(pos), "graphs": graphs, "event_ms_per_step": e0.elapsed_time(e1) / K, "wall_ms_per_step": 1e3 * wall / K,
                          "M_upd_s": len(pos) * K / (e0.elapsed_time(e1) * 1e-3) / 1e6}))
        ctx.close()
PY
