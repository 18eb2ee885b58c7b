"""CPU: the plain-C oracle (oracle/sph_oracle.c) reproduces, bit for bit, every golden vector that was
generated from the unmodified reference compiled IEEE-strict (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest

from helpers import GOLDEN, KERNEL_FIXTURES, STEP_FIXTURES, assert_bits, golden_steps, load_golden, params_from


@pytest.mark.parametrize("name", STEP_FIXTURES + KERNEL_FIXTURES)
def test_port_matches_reference_golden(po, name):
    """KERNEL_FIXTURES: the reference engine with its Wendland C2 / Gaussian class in the kernel slot (kernels.cpp:166-236)."""
    g = load_golden(name)
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    e = po.Engine("port", n)
    e.initialize(prm)
    if "kernel_type" in g:
        e.set_kernel(int(g["kernel_type"]))
    e.add_particles(g["pos"], g["vel"], g["mass"])
    keep = golden_steps(g)
    for k, dt in enumerate(g["dts"]):
        keys = e.keys()
        e.step(float(dt))
        if k in keep:
            s = e.state()
            assert_bits(keys, g[f"s{k}_keys"], f"{name} step {k} keys")
            assert_bits(e.neighbor_counts().astype(np.uint32), g[f"s{k}_counts"], f"{name} step {k} neighbour counts")
            for f in ("rho", "P", "acc", "pos", "vel"):
                if f"s{k}_{f}" in g:
                    assert_bits(s[f], g[f"s{k}_{f}"], f"{name} step {k} {f}")
            assert np.float32(e.time) == g[f"s{k}_time"]
    assert np.float32(e.total_mass()) == g["final_total_mass"] or (np.isnan(e.total_mass()) and np.isnan(g["final_total_mass"]))
    assert np.float32(e.total_energy()) == g["final_total_energy"] or (np.isnan(e.total_energy()) and np.isnan(g["final_total_energy"]))
    assert int(e.stats()["max_neighbors"]) == int(g["final_max_neighbors"])
    e.close()


def test_kernel_known_answers(po):
    g = load_golden("kat_kernel_keys")
    e = po.Engine("port", 16)
    base = dict(po.default_params("port"))
    for h in (0.02, 0.025, 0.008):
        prm = dict(base); prm["smoothing_length"] = h
        e.initialize(prm)
        r = g[f"h{h}_r"]
        assert_bits(np.array([e.kernel_W(x) for x in r], np.float32), g[f"h{h}_W"], f"W h={h}")
        assert_bits(np.array([e.kernel_gradW(x) for x in r], np.float32), g[f"h{h}_gradW"], f"gradW h={h}")
        assert_bits(np.array([e.kernel_lapW(x) for x in r], np.float32), g[f"h{h}_lapW"], f"lapW h={h}")
    # hand-checkable values: W(0) = sigma * 2/3, W(q=1) = sigma/6, W(q=2) = 0 with sigma = 1/(pi h^3)
    h = np.float32(0.02)
    prm = dict(base); prm["smoothing_length"] = h
    e.initialize(prm)
    sigma = np.float32(1.0) / (np.float32(np.pi) * h * h * h)
    assert np.float32(e.kernel_W((0, 0, 0))) == sigma * np.float32(2.0 / 3.0)
    assert abs(e.kernel_W((float(h), 0, 0)) / float(sigma) - 1.0 / 6.0) < 1e-6
    assert e.kernel_W((float(2 * h), 0, 0)) == 0.0
    assert e.kernel_lapW((0, 0, 0)) == np.float32(sigma * np.float32(-2.0)) / np.float32(h * h)
    e.close()


def test_cell_key_known_answers(po):
    g = load_golden("kat_kernel_keys")
    base = dict(po.default_params("port"))
    for cell in (0.04, 0.016, 0.05):
        prm = dict(base); prm["neighbor_search_radius"] = cell
        pts = g[f"cell{cell}_pts"]
        e = po.Engine("port", pts.shape[0]); e.initialize(prm); e.add_particles(pts)
        assert_bits(e.keys(), g[f"cell{cell}_keys"], f"keys cell={cell}")
        e.close()
    # cell (0,0,0) → 0 ; cell (-1,-1,-1) → 0x7FFFFFFFFFFFFFFF ; cell (25,-25,3)
    prm = dict(base)
    e = po.Engine("port", 3); e.initialize(prm)
    e.add_particles(np.array([[0.01, 0.01, 0.01], [-0.01, -0.01, -0.01], [1.01, -0.99, 0.13]], np.float32))
    k = e.keys()
    assert int(k[0]) == 0
    assert int(k[1]) == 0x7FFFFFFFFFFFFFFF
    assert int(k[2]) == (25 << 42) | (((-25) & 0x1FFFFF) << 21) | 3
    e.close()


def test_adaptive_timestep_and_scalars(po, graft):
    meta = json.loads((GOLDEN / "scalars.json").read_text())
    graft.load_package()
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    e = po.Engine("port", pos.shape[0]); e.initialize(prm); e.add_particles(pos, None, mass)
    for want_dt, want_t in zip(meta["adaptive_dts"], meta["adaptive_times"]):
        assert np.float32(e.cfl_timestep()) == np.float32(want_dt)
        e.step(0.0)
        assert np.float32(e.time) == np.float32(want_t)
    assert np.float32(e.total_mass()) == np.float32(meta["adaptive_total_mass"])
    assert np.float32(e.total_energy()) == np.float32(meta["adaptive_total_energy"])
    e.close()


@pytest.mark.parametrize("name", ["perf_test", "fast_cloud", "light_zero", "light_zero_mixed"])
def test_cfl_limited_adaptive_timestep(po, graft, name):
    """compute_cfl_timestep where it bites (sph_engine.cpp:312-333): dt_cfl = CFL h / max|v| and dt_force =
    CFL sqrt(h / |a_0|) (particle 0 only, previous step's acceleration), incl. the reference benchmark's own set-up
    (benchmarks/performance_test.cpp:85-125).  The C restatement reproduces the reference's dt sequence bit for bit."""
    from helpers import cfl_case_inputs
    import hashlib
    graft.load_package()
    from sph_b200 import scenes
    case = json.loads((GOLDEN / "scalars.json").read_text())["cfl_cases"][name]
    assert set(case["branch"]) - {"timestep"}, "the case must leave dt = params.timestep"
    prm, pos, vel, mass, cap = cfl_case_inputs(po, scenes, name, case)
    e = po.Engine("port", cap); e.initialize(prm); e.add_particles(pos, vel, mass)
    for want_dt, want_t in zip(case["dts"], case["times"]):
        assert np.float32(e.cfl_timestep()) == np.float32(want_dt)
        e.step(0.0)
        assert np.float32(e.time) == np.float32(want_t)
    st = e.state()
    for f in ("pos", "vel", "rho"):
        assert hashlib.sha256(np.ascontiguousarray(st[f]).tobytes()).hexdigest() == case[f"final_{f}_sha256"], f
    e.close()


def test_engine_quirks(po):
    """Behavioural quirks of SPHEngine the drop-in must keep."""
    meta = json.loads((GOLDEN / "scalars.json").read_text())
    e = po.Engine("port", 100000)
    assert not e.L.is_initialized(e.h)
    e.step(0.001)                                   # not initialised → no-op
    assert e.step_count == 0
    e.initialize_dam_break()                        # auto-initialises with defaults
    assert e.size == meta["initialize_dam_break_n"] == 84800
    p = e.get_parameters()
    assert p["neighbor_search_radius"] == np.float32(0.04) and p["smoothing_length"] == np.float32(0.02)
    e.set_smoothing_length(0.0125)                  # Q1: only this setter ties nsr to 2h
    assert e.get_parameters()["neighbor_search_radius"] == np.float32(2.0) * np.float32(0.0125)
    assert e.densities_raw().shape[0] == 100000     # Q13: capacity-length buffer
    e.close()
    e = po.Engine("port", 20000)                    # Q14 / §0.5: capacity truncation keeps the first 20 000 (all wall)
    e.initialize_dam_break()
    assert e.size == 20000
    e.step(0.001); e.step(0.001)
    assert e.step_count == 2
    e.initialize_fluid_drop()                       # Q15: re-initialising keeps time and step count ...
    assert e.step_count == 2 and e.time > 0
    e.clear_particles()                             # ... clear_particles resets them
    assert e.step_count == 0 and e.time == 0.0 and e.size == 0
    e.step(0.001)                                   # empty system → no-op
    assert e.step_count == 0
    e.close()
