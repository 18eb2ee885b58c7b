# round 2 (session 3), job s: lanes-per-particle kernels — parity, whole suite, small-scene step times
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -k "lanes" 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -12
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head -12
python - <<'PY'
import time, json, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package()
from sph_b200 import scenes
capi = pkg.capi
import torch
for name, dx in (("dam_break_3k", 0.035), ("dam_break_13k", 0.02), ("dam_break_85k", 0.0105), ("dam_break_347k", None)):
    if dx is None:
        fam, dx = scenes.SCENES[name]
    pos, mass, prm, dt = scenes.dam_break_scene(dx)
    for lanes, mode in ((1, 2), (1, 0), (2, 0), (4, 0), (8, 0)):
        ctx = pkg.Context(len(pos), 0)
        ctx.set_option(capi.OPT_LANES_PER_PARTICLE, lanes)
        ctx.set_option(capi.OPT_PAIR_MODE, mode)
        ctx.set_option(capi.OPT_GRID_REFINE, 4)
        ctx.set_params(prm)
        ctx.upload(pos, None, mass)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(68):
            ctx.step(dt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 300
        e0.record()
        for _ in range(K):
            ctx.step(dt)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"scene": name, "n": len(pos), "lanes": lanes, "pair_mode": mode, "ms_per_step": round(e0.elapsed_time(e1) / K, 4),
                          "M_upd_s": round(len(pos) * K / (e0.elapsed_time(e1) * 1e-3) / 1e6, 1)}))
        ctx.close()
PY
