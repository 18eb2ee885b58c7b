"""GPU: slab decomposition parity — G slabs must reproduce the single-context run bit for bit
(every local step re-sorts owned + ghost particles by (cell, global id), so layouts and summation
orders do not depend on G).  Runs G contexts on ONE GPU inside one process (device copies instead of
NCCL); the NCCL path itself needs >= 2 GPUs and is skipped otherwise."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_bits, load_golden, params_from

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def single_run(pkg, params, pos, vel, mass, dt, steps, strict, layout_major=2):
    capi = pkg.capi
    ctx = pkg.Context(len(pos), 0)
    ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT if strict else capi.MATH_FAST)
    # fast mode: a slab context orders its cells with the slab axis most significant; the single context is given
    # the same layout so that the fp32 summation orders (and therefore the bits) agree.  Strict mode ignores it.
    ctx.set_option(capi.OPT_LAYOUT_MAJOR, layout_major)
    ctx.set_params(params)
    ctx.upload(pos, vel, mass)
    for _ in range(steps):
        ctx.step(dt)
    out = ctx.download()
    ctx.close()
    return out


def slab_run(pkg, params, pos, vel, mass, dt, steps, strict, G, axis=2, cuts=None):
    from sph_b200 import slab
    n = len(pos)
    nsr = float(params["neighbor_search_radius"])
    if cuts is None:
        cuts = slab.plan_cuts(slab.axis_cells(pos, axis, nsr), G, 2)
    box_min = np.minimum(pos.min(0), [params["xmin"], params["ymin"], params["zmin"]])
    box_max = np.maximum(pos.max(0), [params["xmax"], params["ymax"], params["zmax"]])
    ranks = []
    for d in range(G):
        store = slab.GpuStore(pkg, n, 0, params, strict=strict)
        ranks.append(slab.SlabRank(store, d, cuts, axis, 2, n, box_min, box_max, 3 * n))   # <= 3 records per owned particle
        ranks[-1].load_initial(pos, vel, mass, nsr)
    for _ in range(steps):
        slab.step_local(ranks, dt)
    merged = slab.gather_by_id([r.store.download() for r in ranks], n)
    stats = [dict(r.stats) for r in ranks]
    for r in ranks:
        r.store.close()
    return merged, stats


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("G", [2, 3])
def test_dam_break_slabs_equal_single(pkg, G, strict):
    from sph_b200 import scenes
    pos, mass, params, dt = scenes.dam_break_scene(0.02)
    want = single_run(pkg, params, pos, None, mass, dt, 6, strict)
    got, stats = slab_run(pkg, params, pos, None, mass, dt, 6, strict, G)
    assert (got["owners"] == 1).all()
    for f in ("pos", "vel", "rho", "P", "acc"):
        assert_bits(got[f], want[f], f"G={G} strict={strict} {f}")
    assert all(st["halo_sent"] > 0 for st in stats)


def test_migration_heavy_cloud(pkg):
    """Particles streaming along the slab axis (several cells per step, bouncing off the walls)."""
    g = load_golden("cloud600")
    params = params_from(g["params"])
    params = {k: float(v) for k, v in params.items()}
    vel = g["vel"].copy()
    vel[:, 2] = np.where(np.arange(600) % 2 == 0, 45.0, -45.0).astype(np.float32)
    dt = 1e-3
    want = single_run(pkg, params, g["pos"], vel, g["mass"], dt, 12, True)
    from sph_b200 import slab
    for G, cuts in ((2, None), (4, np.array([slab.OPEN_LO, -2, 0, 2, slab.OPEN_HI], np.int32))):
        got, stats = slab_run(pkg, params, g["pos"], vel, g["mass"], dt, 12, True, G, cuts=cuts)
        assert (got["owners"] == 1).all()
        for f in ("pos", "vel", "rho", "acc"):
            assert_bits(got[f], want[f], f"G={G} {f}")
        assert sum(s["migrants_sent"] for s in stats) > 200


def test_slab_along_x(pkg):
    from sph_b200 import scenes
    pos, mass, params, dt = scenes.dam_break_scene(0.02)
    want = single_run(pkg, params, pos, None, mass, dt, 3, True)
    got, _ = slab_run(pkg, params, pos, None, mass, dt, 3, True, 2, axis=0)
    for f in ("pos", "rho", "acc"):
        assert_bits(got[f], want[f], f"x-slabs {f}")


NCCL_WORKER = r'''
import sys, numpy as np, torch, torch.distributed as dist, os
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import __graft_entry__ as g
pkg = g.load_package()
from sph_b200 import slab, scenes
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
pos, mass, params, dt = scenes.dam_break_scene(0.02)
n = len(pos); nsr = float(params["neighbor_search_radius"])
cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), world, 2)
store = slab.GpuStore(pkg, n, local, params, strict=True, stream=torch.cuda.current_stream().cuda_stream)
r = slab.SlabRank(store, rank, cuts, 2, 2, n, [-0.2, 0.0, -0.4], [0.2, 0.6, 0.4], 3 * n)
r.load_initial(pos, None, mass, nsr)
for _ in range(6):
    slab.step_distributed(r, dt)
parts = [None] * world
dist.all_gather_object(parts, store.download())
if rank == 0:
    merged = slab.gather_by_id(parts, n)
    ctx = pkg.Context(n, local); ctx.set_option(pkg.capi.OPT_MATH_MODE, 0); ctx.set_params(params); ctx.upload(pos, None, mass)
    for _ in range(6): ctx.step(dt)
    want = ctx.download()
    for f in ("pos", "vel", "rho", "acc"):
        assert np.array_equal(merged[f].view(np.uint32), want[f].view(np.uint32)), f
    print("NCCL_SLAB_OK")
dist.destroy_process_group()
'''


def test_nccl_two_gpus(pkg, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (NCCL refuses two ranks on one device)")
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29547", str(script), str(ROOT)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "NCCL_SLAB_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_adaptive_timestep_across_slabs(pkg):
    """dt <= 0: the CFL inputs (max |v|^2, acceleration of particle 0) are reduced across slabs; every step's
    dt and the final state equal the single-context adaptive run bit for bit."""
    from sph_b200 import scenes, slab
    pos, mass, params, _ = scenes.dam_break_scene(0.02)
    n = len(pos)
    capi = pkg.capi
    ctx = pkg.Context(n, 0)
    ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT)
    ctx.set_params(params); ctx.upload(pos, None, mass)
    want_dts = []
    for _ in range(5):
        want_dts.append(np.float32(ctx.cfl_timestep())); ctx.step(0.0)
    want = ctx.download(); want_t = ctx.get_time()[0]; ctx.close()

    nsr = float(params["neighbor_search_radius"])
    cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), 3, 2)
    ranks = []
    for d in range(3):
        store = slab.GpuStore(pkg, n, 0, params, strict=True)
        ranks.append(slab.SlabRank(store, d, cuts, 2, 2, n, [-0.2, 0.0, -0.4], [0.2, 0.6, 0.4], 3 * n))
        ranks[-1].load_initial(pos, None, mass, nsr)
    got_dts = [np.float32(slab.step_local(ranks, 0.0)) for _ in range(5)]
    assert [float(x) for x in got_dts] == [float(x) for x in want_dts]
    merged = slab.gather_by_id([r.store.download() for r in ranks], n)
    for f in ("pos", "vel", "rho", "acc"):
        assert_bits(merged[f], want[f], f"adaptive slabs {f}")
    for r in ranks:
        assert np.float32(r.store.ctx.get_time()[0]) == np.float32(want_t)
        r.store.close()


def test_slab_mode_guards_and_diagnostics(pkg):
    """Dense-by-id getters are refused in slab mode; diagnostics sum owned particles only, so the slabs' sums add
    up to the single-context values."""
    from sph_b200 import scenes, slab
    pos, mass, params, dt = scenes.dam_break_scene(0.02)
    n = len(pos)
    ctx = pkg.Context(n, 0); ctx.set_params(params); ctx.upload(pos, None, mass)
    ctx.set_option(pkg.capi.OPT_LAYOUT_MAJOR, 2)      # the slab contexts below order their cells z-major (fast mode)
    for _ in range(2):
        ctx.step(dt)
    want = ctx.diagnostics(); ctx.close()
    nsr = float(params["neighbor_search_radius"])
    cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), 2, 2)
    ranks = []
    for d in range(2):
        store = slab.GpuStore(pkg, n, 0, params)
        ranks.append(slab.SlabRank(store, d, cuts, 2, 2, n, [-0.2, 0.0, -0.4], [0.2, 0.6, 0.4], 3 * n))
        ranks[-1].load_initial(pos, None, mass, nsr)
    for _ in range(2):
        slab.step_local(ranks, dt)
    with pytest.raises(pkg.SphbError):
        ranks[0].store.ctx.download()
    got = [r.store.ctx.diagnostics() for r in ranks]
    assert abs(sum(g[0] for g in got) - want[0]) <= 1e-9 * abs(want[0])
    assert abs(sum(g[1] for g in got) - want[1]) <= 1e-9 * abs(want[1]) + 1e-15
    assert max(g[2] for g in got) == want[2]
    for r in ranks:
        r.store.close()


def test_sparse_drop_slabs_track_bbox(pkg):
    """Fluid drop in a much larger AABB, cut into slabs: each context sizes its cell table from the tracked
    bounding box of owned + arriving particles; results still equal the single-context run bit for bit."""
    from sph_b200 import scenes
    pos, mass, params, dt = scenes.fluid_drop_scene(0.004)
    want = single_run(pkg, params, pos, None, mass, dt, 5, True)
    for G in (2, 3):
        got, stats = slab_run(pkg, params, pos, None, mass, dt, 5, True, G)
        assert (got["owners"] == 1).all()
        for f in ("pos", "vel", "rho", "acc"):
            assert_bits(got[f], want[f], f"drop G={G} {f}")


@pytest.mark.parametrize("name", ["fast_cloud", "light_zero_mixed"])
def test_cfl_limited_adaptive_timestep_across_slabs(pkg, po, name):
    """The CFL-limited cases of tests/golden/scalars.json on 2 slabs: max |v|^2 is reduced across the slabs and the
    acceleration of particle 0 is handed over by the rank that advanced it, so both branches of the rule (dt_cfl,
    dt_force) give the reference's dt, times and final state bit for bit."""
    import hashlib, json
    from helpers import GOLDEN, cfl_case_inputs
    from sph_b200 import scenes, slab
    case = json.loads((GOLDEN / "scalars.json").read_text())["cfl_cases"][name]
    prm, pos, vel, mass, n = cfl_case_inputs(po, scenes, name, case)
    nsr = float(prm["neighbor_search_radius"])
    cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), 2, 2)
    ranks = []
    for d in range(2):
        store = slab.GpuStore(pkg, n, 0, prm, strict=True)
        ranks.append(slab.SlabRank(store, d, cuts, 2, 2, n, [-0.2, 0.0, -0.4], [0.2, 0.6, 0.4], 3 * n))
        ranks[-1].load_initial(pos, vel, mass, nsr)
    for k, want_dt in enumerate(case["dts"]):
        assert np.float32(slab.step_local(ranks, 0.0)) == np.float32(want_dt), f"{name} step {k} ({case['branch'][k]} branch)"
    merged = slab.gather_by_id([r.store.download() for r in ranks], n)
    for f in ("pos", "vel", "rho"):
        assert hashlib.sha256(np.ascontiguousarray(merged[f]).tobytes()).hexdigest() == case[f"final_{f}_sha256"], f"{name} final {f}"
    for r in ranks:
        assert np.float32(r.store.ctx.get_time()[0]) == np.float32(case["times"][-1])
        r.store.close()


def test_slab_download_begin_end(pkg):
    """sphb_slab_download_begin / _end (the read-back that overlaps the next upload): same owned set and fields as the
    synchronous sphb_slab_download, count reported by _end; an output capacity below the owned count is an error."""
    import torch
    from sph_b200 import scenes, slab
    pos, mass, params, dt = scenes.dam_break_scene(0.02)
    n = len(pos)
    nsr = float(params["neighbor_search_radius"])
    cuts = slab.plan_cuts(slab.axis_cells(pos, 2, nsr), 2, 2)
    box_min = np.minimum(pos.min(0), [params["xmin"], params["ymin"], params["zmin"]])
    box_max = np.maximum(pos.max(0), [params["xmax"], params["ymax"], params["zmax"]])
    ranks = []
    for d in range(2):
        store = slab.GpuStore(pkg, n, 0, params, strict=False)
        ranks.append(slab.SlabRank(store, d, cuts, 2, 2, n, box_min, box_max, 3 * n))
        ranks[-1].load_initial(pos, None, mass, nsr)
    for _ in range(3):
        slab.step_local(ranks, dt)
    for r in ranks:
        ctx = r.store.ctx
        want = ctx.slab_download()
        cap = ctx.size
        ids = torch.empty((cap,), dtype=torch.int32).pin_memory()
        p = torch.empty((cap, 3), dtype=torch.float32).pin_memory()
        v = torch.empty((cap, 3), dtype=torch.float32).pin_memory()
        rho = torch.empty((cap,), dtype=torch.float32).pin_memory()
        ctx.slab_download_begin_raw(cap, ids.data_ptr(), p.data_ptr(), v.data_ptr(), rho.data_ptr(), None, None)
        k = ctx.slab_download_end()
        assert k == len(want["ids"]) and 0 < k < cap                       # the slab also holds halo copies
        order_a, order_b = np.argsort(ids.numpy()[:k].view(np.uint32)), np.argsort(want["ids"])
        assert_bits(ids.numpy()[:k].view(np.uint32)[order_a], want["ids"][order_b], "ids")
        assert_bits(p.numpy()[:k][order_a], want["pos"][order_b], "pos")
        assert_bits(v.numpy()[:k][order_a], want["vel"][order_b], "vel")
        assert_bits(rho.numpy()[:k][order_a], want["rho"][order_b], "rho")
        ctx.slab_download_begin_raw(k - 1, ids.data_ptr(), None, None, None, None, None)
        with pytest.raises(pkg.SphbError):
            ctx.slab_download_end()
        r.store.close()


def test_read_small_kernel_written_readback(pkg):
    """sphb_read_small: the exchange driver's read-back of the gathered group-size table — a kernel writes it into pinned
    host memory (a device-to-host copy would queue on the copy engine behind a bulk read-back in flight)."""
    import torch
    ctx = pkg.Context(16, 0)
    d = torch.arange(1, 129, dtype=torch.int32, device="cuda:0") * 7
    h = torch.zeros(128, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    ctx.read_small(d.data_ptr(), h.data_ptr(), 128 * 4)
    ctx.synchronize()
    assert torch.equal(h, d.cpu())
    with pytest.raises(pkg.SphbError):
        ctx.read_small(d.data_ptr(), h.data_ptr(), 6)            # not a multiple of 4
    with pytest.raises(pkg.SphbError):
        ctx.read_small(d.data_ptr(), h.data_ptr(), 4 << 20)      # not small
    ctx.close()
