#!/usr/bin/env python
"""Which transfers of the end-to-end cycle hide behind the step?  (debug aid: cycle times of partial cycles, dam_break_10M)"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
pkg = graft.load_package()
from sph_b200 import scenes
capi = pkg.capi
pos, mass, prm, dt = scenes.make_scene("dam_break_10M")
n = len(pos)
h_pos = torch.from_numpy(pos).pin_memory(); h_vel = torch.zeros((n, 3), dtype=torch.float32).pin_memory(); h_mass = torch.from_numpy(mass).pin_memory()
o = [torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty((n,), dtype=torch.float32).pin_memory()]
ctx = pkg.Context(n, 0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_params(prm)
up = lambda: ctx.upload_raw(n, h_pos.data_ptr(), h_vel.data_ptr(), h_mass.data_ptr())
st = lambda: ctx.step(dt)
db = lambda: ctx.download_begin_raw(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None, None)
ds = lambda: ctx.download_raw(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None, None)
up(); st(); torch.cuda.synchronize()
def run(name, fns, reps=8):
    for f in fns: f()
    ctx.download_end(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns: f()
    ctx.download_end(); torch.cuda.synchronize()
    print(f"{name:42s} {1e3 * (time.perf_counter() - t0) / reps:7.2f} ms per cycle", flush=True)
run("step", [st]); run("upload", [up]); run("download (sync)", [ds]); run("download_begin", [db])
run("upload + step", [up, st]); run("step + download_begin", [st, db]); run("upload + step + download (sync)", [up, st, ds])
run("upload + step + download_begin", [up, st, db])
run("upload + download_begin (no step)", [up, db])
run("download_begin + upload (no step)", [db, up])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
# where does the time of the full cycle go?  host timestamps of each call
def traced(reps=6):
    for f in (up, st, db): f()
    ctx.download_end(); torch.cuda.synchronize()
    t0 = time.perf_counter(); marks = []
    for _ in range(reps):
        a = time.perf_counter(); up(); b = time.perf_counter(); st(); c = time.perf_counter(); db(); d = time.perf_counter()
        marks.append((1e3 * (a - t0), 1e3 * (b - a), 1e3 * (c - b), 1e3 * (d - c)))
    ctx.download_end(); torch.cuda.synchronize()
    for m in marks:
        print("   cycle starts at %7.2f ms: upload call %6.2f, step call %6.2f, download_begin call %6.2f" % m)
traced()
