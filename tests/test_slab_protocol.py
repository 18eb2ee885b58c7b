"""CPU: the multi-GPU slab protocol (planning, migration, halo selection) with a numpy test double as the
particle store — in-process with 3 ranks, and over torch.distributed/gloo with world_size 2."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

BOUNDS = (-0.5, 0.5, -0.5, 0.5, -1.0, 1.0)
NSR = 0.04


def make_cloud(n=4000, seed=3):
    rng = np.random.default_rng(seed)
    pos = rng.uniform([-0.5, -0.5, -1.0], [0.5, 0.5, 1.0], size=(n, 3)).astype(np.float32)
    vel = rng.normal(0, 1.0, size=(n, 3)).astype(np.float32)
    vel[:, 2] *= 40.0          # fast along the slab axis: several cells (and sometimes whole slabs) per step
    return pos, vel


def reference_drift(pos, vel, dt):
    lo = np.array(BOUNDS[0::2], np.float32); hi = np.array(BOUNDS[1::2], np.float32)
    pos = pos + vel * np.float32(dt)
    out = (pos < lo) | (pos > hi)
    vel = np.where(out, vel * np.float32(-0.8), vel)
    return np.minimum(np.maximum(pos, lo), hi), vel


def test_plan_cuts(pkg):
    from sph_b200 import slab
    pos, _ = make_cloud()
    cells = slab.axis_cells(pos, 2, NSR)
    for g in (1, 2, 3, 8):
        cuts = slab.plan_cuts(cells, g, min_width=2)
        assert len(cuts) == g + 1 and cuts[0] == slab.OPEN_LO and cuts[-1] == slab.OPEN_HI
        assert (np.diff(cuts) >= 2).all()
        counts = np.bincount(slab.rank_of_cells(cuts, cells), minlength=g)
        assert counts.sum() == len(pos) and counts.min() > 0.7 * len(pos) / g, counts
    with pytest.raises(ValueError):
        slab.plan_cuts(np.array([0, 1, 2, 3]), 4, min_width=2)
    # dam-break z extent at the 1M resolution: 50 cells → 8 slabs of >= 2 cells
    z = np.linspace(-0.4, 0.4, 1000, dtype=np.float32)
    cuts = slab.plan_cuts(slab.axis_cells(np.stack([z * 0, z * 0, z], 1), 2, 0.016), 8)
    assert (np.diff(cuts[1:-1]) >= 2).all()


def test_protocol_in_process(pkg):
    from sph_b200 import slab
    from slab_doubles import NumpyStore, expected_sets
    pos, vel = make_cloud()
    n = len(pos)
    G, L = 3, 2
    cuts = slab.plan_cuts(slab.axis_cells(pos, 2, NSR), G, L)
    ranks = [slab.SlabRank(NumpyStore(NSR, BOUNDS), d, cuts, 2, L, n, BOUNDS[0::2], BOUNDS[1::2], 3 * n) for d in range(G)]
    for r in ranks:
        r.load_initial(pos, vel, None, NSR)
    gp, gv = pos.copy(), vel.copy()
    dt = 0.004
    for step in range(8):
        slab.step_local(ranks, dt)
        want = expected_sets(gp, cuts, 2, NSR, L)
        for d, r in enumerate(ranks):
            own, ghost = r.store.snapshots[-1]
            assert np.array_equal(own, np.sort(want[d][0])), f"step {step} rank {d}: owned set"
            assert np.array_equal(ghost, np.sort(want[d][1])), f"step {step} rank {d}: ghost set"
        gp, gv = reference_drift(gp, gv, dt)
    assert sum(r.stats["migrants_sent"] for r in ranks) > 100          # the scene really migrates
    merged = slab.gather_by_id([r.store.download() for r in ranks], n)
    assert (merged["owners"] == 1).all()
    assert np.array_equal(merged["pos"], gp) and np.array_equal(merged["vel"], gv)


WORKER = r'''
import sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import __graft_entry__ as g
g.load_package()
from sph_b200 import slab
from slab_doubles import NumpyStore, expected_sets
import test_slab_protocol as T
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pos, vel = T.make_cloud()
n = len(pos)
cuts = slab.plan_cuts(slab.axis_cells(pos, 2, T.NSR), world, 2)
r = slab.SlabRank(NumpyStore(T.NSR, T.BOUNDS), rank, cuts, 2, 2, n, T.BOUNDS[0::2], T.BOUNDS[1::2], 3 * n)
r.load_initial(pos, vel, None, T.NSR)
gp, gv = pos.copy(), vel.copy()
for step in range(6):
    slab.step_distributed(r, 0.004)
    want = expected_sets(gp, cuts, 2, T.NSR, 2)[rank]
    own, ghost = r.store.snapshots[-1]
    assert np.array_equal(own, np.sort(want[0])), (step, rank, "owned")
    assert np.array_equal(ghost, np.sort(want[1])), (step, rank, "ghost")
    gp, gv = T.reference_drift(gp, gv, 0.004)
parts = [None] * world
dist.all_gather_object(parts, r.store.download())
merged = slab.gather_by_id(parts, n)
assert (merged["owners"] == 1).all() and np.array_equal(merged["pos"], gp)
print("RANK_OK", rank, r.stats)
dist.destroy_process_group()
'''


def test_protocol_gloo_world2(pkg, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29541", str(script), str(ROOT)], capture_output=True, text=True,
                         env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("RANK_OK") == 2
