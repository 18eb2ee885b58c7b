"""GPU: the CUDA path through the C ABI against the oracle and the committed golden vectors.

Bars (BASELINE.json north star):
  * cell keys, sorted permutation, neighbour counts: bit-exact in both math modes
  * strict math mode: density, pressure, acceleration, position, velocity, time bit-exact
  * fast math mode (default): single step rho rel <= 2e-5, acc <= 1e-4 of max|acc|, pos abs <= 1e-6 * L
    (SURVEY.md §8c gates; the reference's own fast-math vs strict noise floor is 4.4e-7 / 1.5e-8)
"""
import json

import numpy as np
import pytest

from helpers import GOLDEN, KERNEL_FIXTURES, STEP_FIXTURES, assert_bits, golden_steps, load_golden, params_from, rel_err, stable_perm

pytestmark = pytest.mark.gpu

TOL_RHO = 2e-5
TOL_ACC = 1e-4
TOL_POS = 1e-6


def make_ctx(pkg, n, params, strict=True, **opts):
    capi = pkg.capi
    ctx = pkg.Context(max(n, 1), 0)
    ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT if strict else capi.MATH_FAST)
    ctx.set_option(capi.OPT_DEBUG_CAPTURE, 1)
    for k, v in opts.items():
        ctx.set_option(getattr(capi, k), v)
    ctx.set_params(params)
    return ctx


def check_fast(got, want, L, what):
    assert rel_err(got["rho"], want["rho"]) <= TOL_RHO, f"{what}: rho rel {rel_err(got['rho'], want['rho'])}"
    amax = np.abs(want["acc"]).max()
    if np.isfinite(amax) and amax > 0:
        assert np.abs(got["acc"].astype(np.float64) - want["acc"]).max() <= TOL_ACC * amax, f"{what}: acc"
    assert np.abs(got["pos"].astype(np.float64) - want["pos"]).max() <= TOL_POS * L, f"{what}: pos"


@pytest.mark.parametrize("variant", [0, 2])   # strict mode always runs the reference-order tested walk, whatever is selected
@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_strict_bit_exact_vs_reference_golden(pkg, name, variant):
    g = load_golden(name)
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    ctx = make_ctx(pkg, n, prm, strict=True, OPT_PAIR_KERNEL=variant)
    ctx.upload(g["pos"], g["vel"], g["mass"])
    keep = golden_steps(g)
    for k, dt in enumerate(g["dts"]):
        ctx.step(float(dt))
        if k in keep:
            d = ctx.debug_dump()
            assert_bits(d["keys"], g[f"s{k}_keys"], f"{name} step {k} keys")
            assert_bits(d["perm"], stable_perm(g[f"s{k}_keys"]), f"{name} step {k} sorted permutation")
            assert_bits(d["nbr_count"], g[f"s{k}_counts"], f"{name} step {k} neighbour counts")
            s = ctx.download()
            for f in ("rho", "P", "acc", "pos", "vel"):
                if f"s{k}_{f}" in g:
                    assert_bits(s[f], g[f"s{k}_{f}"], f"{name} step {k} {f}")
            t, sc = ctx.get_time()
            assert np.float32(t) == g[f"s{k}_time"] and sc == k + 1
    assert ctx.stats()["max_neighbors"] == int(g["final_max_neighbors"])
    ctx.close()


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("name", ["micro_pair", "micro_coincident", "micro_lattice27", "cloud600", "cloud600_truncated_support",
                                  "cloud600_wide_cell", "dam_break_13k_tame"])
def test_fast_mode_single_step_tolerance(pkg, name, variant):
    g = load_golden(name)
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    ctx = make_ctx(pkg, n, prm, strict=False, OPT_PAIR_KERNEL=variant)
    ctx.upload(g["pos"], g["vel"], g["mass"])
    k0 = golden_steps(g)[0]                     # first step the fixture holds (0 or 1)
    for k in range(k0 + 1):
        ctx.step(float(g["dts"][k]))
    d = ctx.debug_dump()
    if k0 == 0:                                 # after >1 fast steps positions differ in the last bits
        assert_bits(d["keys"], g["s0_keys"], f"{name} keys")
        assert_bits(d["perm"], stable_perm(g["s0_keys"]), f"{name} permutation")
        assert_bits(d["nbr_count"], g["s0_counts"], f"{name} counts (fast mode must keep neighbour sets)")
    got = ctx.download()
    want = {f: g[f"s{k0}_{f}"] for f in ("rho", "P", "acc", "pos", "vel")}
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    check_fast(got, want, L, name)
    ctx.close()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", KERNEL_FIXTURES)
def test_other_kernel_classes_vs_reference_golden(pkg, name, strict):
    """Wendland C2 / Gaussian (reference kernels.cpp:166-236) as template parameters of the tested-walk
    pair kernels (SPHB_OPT_KERNEL_TYPE), against golden vectors of the UNMODIFIED reference engine stepping with those
    classes in its kernel slot.  Keys, permutation and neighbour counts: bit-exact.  Strict Wendland C2: every field
    bit-exact.  Strict Gaussian: the classes call expf, whose glibc and CUDA implementations differ in the last bit —
    held to 2e-6 relative.  Fast mode: the fast-mode gates."""
    g = load_golden(name)
    kt = int(g["kernel_type"])
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    ctx = make_ctx(pkg, n, prm, strict=strict, OPT_KERNEL_TYPE=kt)
    assert ctx.get_option(pkg.capi.OPT_KERNEL_TYPE) == kt
    ctx.upload(g["pos"], g["vel"], g["mass"])
    keep = golden_steps(g)
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    for k, dt in enumerate(g["dts"]):
        ctx.step(float(dt))
        if k not in keep:
            continue
        d = ctx.debug_dump()
        s = ctx.download()
        want = {f: g[f"s{k}_{f}"] for f in ("rho", "P", "acc", "pos", "vel")}
        exact = strict and kt == pkg.capi.KERNEL_WENDLAND_C2
        if exact or k == 0:
            assert_bits(d["keys"], g[f"s{k}_keys"], f"{name} step {k} keys")
            assert_bits(d["perm"], stable_perm(g[f"s{k}_keys"]), f"{name} step {k} sorted permutation")
            assert_bits(d["nbr_count"], g[f"s{k}_counts"], f"{name} step {k} neighbour counts")
        if exact:
            for f in want:
                assert_bits(s[f], want[f], f"{name} step {k} {f}")
            assert np.float32(ctx.get_time()[0]) == g[f"s{k}_time"]
        elif strict:
            assert rel_err(s["rho"], want["rho"]) <= 2e-6 * (k + 1), f"{name} step {k} rho {rel_err(s['rho'], want['rho'])}"
            assert rel_err(s["acc"], want["acc"]) <= 2e-5 * (k + 1), f"{name} step {k} acc {rel_err(s['acc'], want['acc'])}"
            assert np.abs(s["pos"].astype(np.float64) - want["pos"]).max() <= 1e-7 * L
        else:
            scale = 1 if k == 0 else 10          # N steps 10x looser
            assert rel_err(s["rho"], want["rho"]) <= scale * TOL_RHO, f"{name} step {k} rho {rel_err(s['rho'], want['rho'])}"
            assert np.abs(s["acc"].astype(np.float64) - want["acc"]).max() <= scale * TOL_ACC * np.abs(want["acc"]).max(), f"{name} step {k} acc"
            assert np.abs(s["pos"].astype(np.float64) - want["pos"]).max() <= scale * TOL_POS * L
    ctx.close()


def test_kernel_type_option_guards(pkg):
    """Unknown kernel types are refused; the cubic spline stays the default and the bitmask kernels stay its path."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.025)
    ctx = make_ctx(pkg, len(pos), prm, strict=False)
    assert ctx.get_option(pkg.capi.OPT_KERNEL_TYPE) == pkg.capi.KERNEL_CUBIC_SPLINE
    with pytest.raises(pkg.SphbError):
        ctx.set_option(pkg.capi.OPT_KERNEL_TYPE, 3)
    ctx.upload(pos, None, mass)
    ctx.step(dt)
    cubic = ctx.download()
    ctx.set_option(pkg.capi.OPT_KERNEL_TYPE, pkg.capi.KERNEL_WENDLAND_C2)
    ctx.upload(pos, None, mass)
    ctx.step(dt)
    wend = ctx.download()
    # Wendland C2 carries 21/2 : 2/3 of the cubic spline's (2/3-normalised) weight at the origin: densities must differ
    assert rel_err(wend["rho"], cubic["rho"]) > 0.1
    ctx.set_option(pkg.capi.OPT_KERNEL_TYPE, pkg.capi.KERNEL_CUBIC_SPLINE)
    ctx.upload(pos, None, mass)
    ctx.step(dt)
    assert_bits(ctx.download()["rho"], cubic["rho"], "back to the cubic spline")
    ctx.close()


@pytest.mark.parametrize("strict", [True, False])
def test_step_graphs_replay_identical_steps(pkg, strict):
    """SPHB_OPT_STEP_GRAPHS (default on): repeating steps are replayed from a CUDA graph.  Same bits as direct launches —
    fixed dt, adaptive dt (the dt of a replayed step is still the device's CFL rule of that step), across a parameter
    change, a re-upload and a switch of stage timing (which suspends the graphs) — and the launch count keeps counting
    the kernels inside the graph."""
    from sph_b200 import scenes
    capi = pkg.capi
    pos, mass, prm, dt = scenes.dam_break_scene(0.025)
    n = pos.shape[0]

    def run(graphs):
        ctx = pkg.Context(n, 0)
        ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT if strict else capi.MATH_FAST)
        ctx.set_option(capi.OPT_STEP_GRAPHS, graphs)
        assert ctx.get_option(capi.OPT_STEP_GRAPHS) == graphs
        ctx.set_params(prm)
        ctx.upload(pos, None, mass)
        snaps = []
        for _ in range(9):
            ctx.step(dt)
        snaps.append(ctx.download())
        for _ in range(7):
            ctx.step(0.0)                                  # adaptive
        snaps.append(ctx.download()); snaps[-1]["time"] = ctx.get_time()
        p2 = dict(prm); p2["viscosity"] = float(prm["viscosity"]) * 3.0; p2["gravity"] = -4.0
        ctx.set_params(p2)
        for _ in range(6):
            ctx.step(dt)
        snaps.append(ctx.download())
        ctx.set_option(capi.OPT_STAGE_TIMING, 1)
        for _ in range(3):
            ctx.step(dt)
        ctx.set_option(capi.OPT_STAGE_TIMING, 0)
        for _ in range(5):
            ctx.step(dt * 0.5)
        snaps.append(ctx.download())
        ctx.upload(pos[: n // 2], None, mass[: n // 2])   # another particle count: new graphs
        for _ in range(6):
            ctx.step(dt)
        snaps.append(ctx.download())
        st = ctx.stats()
        snaps.append({"launches": st["kernel_launches"], "steps": st["steps"], "time": ctx.get_time()})
        ctx.close()
        return snaps

    a, b = run(1), run(0)
    for k, (x, y) in enumerate(zip(a[:-1], b[:-1])):
        for f in ("pos", "vel", "rho", "P", "acc"):
            assert_bits(x[f], y[f], f"graphs on/off, snapshot {k} {f}")
    assert a[1]["time"] == b[1]["time"]
    assert a[-1] == b[-1]


def test_download_begin_end_overlaps_the_next_upload(pkg):
    """sphb_download_begin / _end: the read-back of one step may still be in flight while the next upload and step are
    enqueued; every read-back equals the synchronous download of the same state, and a second _begin (or destroy) waits."""
    import torch
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    n = len(pos)
    h_pos = torch.from_numpy(pos.copy()).pin_memory()
    h_vel = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
    h_mass = torch.from_numpy(mass.copy()).pin_memory()
    outs = [[torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty((n, 3), dtype=torch.float32).pin_memory(),
             torch.empty((n,), dtype=torch.float32).pin_memory()] for _ in range(2)]
    ctx = make_ctx(pkg, n, prm, strict=False)
    ref = make_ctx(pkg, n, prm, strict=False)
    for k in range(4):
        h_pos[:, 1] += 1e-4                                   # a different input every cycle
        ctx.upload_raw(n, h_pos.data_ptr(), h_vel.data_ptr(), h_mass.data_ptr())
        ctx.step(dt)
        o = outs[k % 2]
        ctx.download_begin_raw(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None, None)
        ref.upload_raw(n, h_pos.data_ptr(), h_vel.data_ptr(), h_mass.data_ptr())   # meanwhile: other work
        ref.step(dt)
        want = ref.download()
        ctx.download_end()
        assert_bits(o[0].numpy(), want["pos"], f"cycle {k} pos")
        assert_bits(o[1].numpy(), want["vel"], f"cycle {k} vel")
        assert_bits(o[2].numpy(), want["rho"], f"cycle {k} rho")
    ctx.download_begin_raw(outs[0][0].data_ptr(), None, None, None, None)
    ctx.download_begin_raw(outs[1][0].data_ptr(), None, None, None, None)       # waits for the first
    assert_bits(outs[0][0].numpy(), ref.download()["pos"], "back-to-back begins")
    ctx.close()                                               # waits for the second
    assert_bits(outs[1][0].numpy(), ref.download()["pos"], "destroy waits for the transfer")
    ref.close()


def test_ten_steps_vs_oracle(pkg, po):
    """N-step parity on the 13k tame dam break: strict stays bit-exact; fast stays inside 10x the
    single-step gates (10 steps 10x looser)."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    n = pos.shape[0]
    ora = po.Engine("port", n); ora.initialize(prm); ora.add_particles(pos, None, mass)
    cs = make_ctx(pkg, n, prm, strict=True); cs.upload(pos, None, mass)
    cf = make_ctx(pkg, n, prm, strict=False); cf.upload(pos, None, mass)
    for k in range(10):
        keys = ora.keys()
        ora.step(dt); cs.step(dt); cf.step(dt)
        d = cs.debug_dump()
        assert_bits(d["keys"], keys, f"step {k} keys")
        assert_bits(d["perm"], stable_perm(keys), f"step {k} permutation")
        assert_bits(d["nbr_count"], ora.neighbor_counts(), f"step {k} counts")
    want = ora.state(); got = cs.download(); fast = cf.download()
    for f in ("rho", "P", "acc", "pos", "vel"):
        assert_bits(got[f], want[f], f"10 steps strict {f}")
    assert rel_err(fast["rho"], want["rho"]) <= 10 * TOL_RHO
    assert np.abs(fast["pos"].astype(np.float64) - want["pos"]).max() <= 10 * TOL_POS * 0.8
    assert np.abs(fast["acc"].astype(np.float64) - want["acc"]).max() <= 10 * TOL_ACC * np.abs(want["acc"]).max()
    # diagnostics agree (mass = sum rho * h^3, kinetic energy)
    sum_rho, ke, vmax = cs.diagnostics()
    h = np.float32(prm["smoothing_length"])
    assert abs(sum_rho * float(h * h * h) - ora.total_mass()) <= 1e-4 * abs(ora.total_mass())
    assert abs(ke - ora.total_energy()) <= 1e-4 * abs(ora.total_energy()) + 1e-12
    assert abs(vmax - np.sqrt((want["vel"].astype(np.float64) ** 2).sum(1)).max()) <= 1e-6 * max(vmax, 1e-9)
    ora.close(); cs.close(); cf.close()


def test_adaptive_timestep(pkg):
    meta = json.loads((GOLDEN / "scalars.json").read_text())
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    ctx = make_ctx(pkg, pos.shape[0], prm, strict=True)
    ctx.upload(pos, None, mass)
    for want_dt, want_t in zip(meta["adaptive_dts"], meta["adaptive_times"]):
        assert np.float32(ctx.cfl_timestep()) == np.float32(want_dt)
        ctx.step(0.0)
        assert np.float32(ctx.get_time()[0]) == np.float32(want_t)
    import hashlib
    s = ctx.download()
    assert hashlib.sha256(s["pos"].tobytes()).hexdigest() == meta["adaptive_final_pos_sha256"]
    assert hashlib.sha256(s["rho"].tobytes()).hexdigest() == meta["adaptive_final_rho_sha256"]
    ctx.close()


@pytest.mark.parametrize("name", ["perf_test", "fast_cloud", "light_zero", "light_zero_mixed"])
def test_cfl_limited_adaptive_timestep(pkg, po, name):
    """k_cfl_dt on the branches the tame scenes never take: dt_cfl = CFL h / max|v| and dt_force = CFL sqrt(h / |a_0|)
    (particle 0 only, previous step's a), incl. the set-up of the reference's own benchmarks/performance_test.cpp
    (exploding default parameters, step() with dt = 0).  Strict mode: every dt, every time and the final state are the
    reference's bits (golden vectors from the unmodified reference, tests/golden/make_golden.py section 5)."""
    import hashlib
    from helpers import cfl_case_inputs
    from sph_b200 import scenes
    case = json.loads((GOLDEN / "scalars.json").read_text())["cfl_cases"][name]
    prm, pos, vel, mass, cap = cfl_case_inputs(po, scenes, name, case)
    ctx = make_ctx(pkg, cap, prm, strict=True)
    ctx.upload(pos, vel, mass)
    for k, (want_dt, want_t) in enumerate(zip(case["dts"], case["times"])):
        assert np.float32(ctx.cfl_timestep()) == np.float32(want_dt), f"{name} step {k} ({case['branch'][k]} branch)"
        ctx.step(0.0)
        assert np.float32(ctx.get_time()[0]) == np.float32(want_t)
    s = ctx.download()
    for f in ("pos", "vel", "rho"):
        assert hashlib.sha256(np.ascontiguousarray(s[f]).tobytes()).hexdigest() == case[f"final_{f}_sha256"], f"{name} final {f}"
    # fast mode walks the same branches: same dt up to the fast-mode tolerances of the velocities behind it
    if name != "perf_test":        # (exploding parameters: fast and strict diverge in the last bits of 1e8-sized velocities)
        cf = make_ctx(pkg, cap, prm, strict=False)
        cf.upload(pos, vel, mass)
        for want_dt in case["dts"]:
            assert abs(cf.cfl_timestep() - want_dt) <= 1e-3 * want_dt
            cf.step(0.0)
        cf.close()
    ctx.close()


def test_walk_radius_two_equals_one(pkg):
    """27-cell walk vs the reference's 125-cell walk (spatial_hash.cpp:35): identical sets and sums."""
    g = load_golden("cloud600")
    prm = params_from(g["params"])
    outs = []
    for r in (1, 2):
        ctx = make_ctx(pkg, 600, prm, strict=True, OPT_WALK_RADIUS=r)
        ctx.upload(g["pos"], g["vel"], g["mass"]); ctx.step(float(g["dts"][0]))
        outs.append((ctx.download(), ctx.debug_dump())); ctx.close()
    for f in ("rho", "acc", "pos"):
        assert_bits(outs[0][0][f], outs[1][0][f], f"walk radius {f}")
    assert_bits(outs[0][1]["nbr_count"], outs[1][1]["nbr_count"], "walk radius counts")


def test_edge_cases(pkg, po):
    capi = pkg.capi
    prm = dict(pkg.DEFAULT_PARAMS)
    # empty system: step is a no-op (sph_engine.cpp:94)
    ctx = make_ctx(pkg, 16, prm)
    ctx.step(0.001)
    assert ctx.get_time() == (0.0, 0) and ctx.size == 0
    assert ctx.download()["pos"].shape == (0, 3)
    # over capacity → error, state untouched
    with pytest.raises(pkg.SphbError) as ei:
        ctx.upload(np.zeros((17, 3), np.float32))
    assert ei.value.code == -3
    # step before set_params
    c2 = pkg.Context(4, 0)
    c2.upload(np.zeros((2, 3), np.float32))
    with pytest.raises(pkg.SphbError):
        c2.step(0.001)
    c2.close()
    # default mass = params.particle_mass, default velocity 0; exactly-full capacity
    pts = np.random.default_rng(7).uniform(-0.05, 0.05, size=(16, 3)).astype(np.float32)
    ctx.upload(pts)
    ctx.step(0.001)
    ora = po.Engine("port", 16); ora.initialize(prm); ora.add_particles(pts, None, np.full(16, prm["particle_mass"], np.float32)); ora.step(0.001)
    for f in ("rho", "pos", "vel", "acc"):
        assert_bits(ctx.download()[f], ora.state()[f], f"edge {f}")
    # NaN / inf positions must not crash (undefined in the reference: (int)floorf(NaN))
    bad = pts.copy(); bad[3] = np.nan; bad[5, 0] = np.inf
    ctx.upload(bad); ctx.step(0.001); ctx.step(0.001)
    out = ctx.download()
    assert np.isfinite(out["pos"][0]).all()
    ctx.close(); ora.close()
    # a grid too large for the dense cell table is refused, not mis-handled
    big = dict(prm); big["neighbor_search_radius"] = 1e-4; big["smoothing_length"] = 5e-5
    c3 = make_ctx(pkg, 4, big)
    c3.upload(np.array([[-1, -1, -1], [1, 1, 1]], np.float32))
    with pytest.raises(pkg.SphbError) as ei:
        c3.step(0.001)
    assert ei.value.code == -4
    c3.close()


def test_strided_aos_roundtrip(pkg, po):
    """Upload/download through the 76-byte sph::Particle layout (particle.h:17-49)."""
    g = load_golden("cloud600")
    prm = params_from(g["params"])
    n = 600
    rec = np.zeros((n, 19), np.float32)      # 76 bytes: pos 0, vel 12, acc 24, density 36, pressure 40, mass 44, ...
    rec[:, 0:3] = g["pos"]; rec[:, 3:6] = g["vel"]; rec[:, 11] = g["mass"]
    ctx = make_ctx(pkg, n, prm, strict=True)
    ctx.upload_strided(n, rec.ctypes.data, 76, 0, 12, 44)
    ctx.step(float(g["dts"][0]))
    ctx.download_strided(rec.ctypes.data, 76, 0, 12, 36, 40)
    assert_bits(np.ascontiguousarray(rec[:, 0:3]), g["s0_pos"], "aos pos")
    assert_bits(np.ascontiguousarray(rec[:, 3:6]), g["s0_vel"], "aos vel")
    assert_bits(np.ascontiguousarray(rec[:, 9]), g["s0_rho"], "aos rho")
    assert_bits(np.ascontiguousarray(rec[:, 10]), g["s0_P"], "aos P")
    assert_bits(np.ascontiguousarray(rec[:, 11]), g["mass"], "aos mass untouched")
    ctx.close()


def test_reupload_and_parameter_change(pkg, po):
    """initialize()/set_smoothing_length between steps (cell size follows nsr), re-upload resets ids."""
    g = load_golden("cloud600")
    prm = params_from(g["params"])
    ctx = make_ctx(pkg, 1000, prm, strict=True)
    ora = po.Engine("port", 1000); ora.initialize(prm); ora.add_particles(g["pos"], g["vel"], g["mass"])
    ctx.upload(g["pos"], g["vel"], g["mass"])
    dt = float(g["dts"][0])
    ora.step(dt); ctx.step(dt)
    ora.set_smoothing_length(0.03); p2 = ora.get_parameters(); ctx.set_params(p2)
    ora.step(dt); ctx.step(dt)
    ora.set_boundaries(-0.1, 0.1, 0.2, 0.4, -0.1, 0.1); ctx.set_params(ora.get_parameters())
    ora.step(dt); ctx.step(dt)
    ora.step(dt); ctx.step(dt)
    want, got = ora.state(), ctx.download()
    for f in ("rho", "P", "acc", "pos", "vel"):
        assert_bits(got[f], want[f], f"param change {f}")
    assert_bits(ctx.debug_dump()["nbr_count"], ora.neighbor_counts(), "param change counts")
    ctx.close(); ora.close()


def test_full_size_properties_1M(pkg):
    """BASELINE.json configs[1] at full size (N = 1 130 000) through size-independent properties:
    the permutation is a bijection that sorts the keys stably; neighbour relation is symmetric (even
    total of off-diagonal pairs); strict and fast (tested walk = f0, bitmask hand-off = f2) agree within the
    fast-mode gates."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.make_scene("dam_break_1M")
    n = pos.shape[0]
    res = {}
    for name, strict, variant in (("s1", True, 0), ("f1", False, 0), ("f2", False, 2)):
        ctx = make_ctx(pkg, n, prm, strict=strict, OPT_PAIR_KERNEL=variant)
        ctx.upload(pos, None, mass)
        ctx.step(dt)
        first = (ctx.download(), ctx.debug_dump())      # identical inputs → discrete outputs must agree
        ctx.step(dt)
        res[name] = (first, ctx.download(), ctx.debug_dump())
        ctx.close()
    for which in (0, 2):                                  # step 1 and step 2 dumps of the strict run
        d = res["s1"][which][1] if which == 0 else res["s1"][2]
        perm, keys, cnt = d["perm"], d["keys"], d["nbr_count"]
        assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32)), "permutation is not a bijection"
        sk = keys[perm]
        assert (np.diff(sk.astype(np.int64)) >= 0).all(), "keys not sorted"
        same = sk[1:] == sk[:-1]
        assert (perm[1:][same] > perm[:-1][same]).all(), "sort not stable (ids not ascending inside a cell)"
        assert (cnt >= 1).all() and int((cnt.astype(np.int64) - 1).sum()) % 2 == 0, "neighbour relation not symmetric"
    for fv in ("f1", "f2"):   # f2 = the default fast path (bitmask hand-off, monotone layout)
        assert_bits(res[fv][0][1]["nbr_count"], res["s1"][0][1]["nbr_count"], f"{fv} vs strict counts (step 1)")
        assert_bits(res[fv][0][1]["perm"], res["s1"][0][1]["perm"], f"{fv} vs strict perm (step 1)")
        assert_bits(res[fv][0][1]["keys"], res["s1"][0][1]["keys"], f"{fv} vs strict keys (step 1)")
    check_fast(res["f2"][0][0], res["s1"][0][0], 0.8, "1M fast (variant 2) vs strict, step 1")
    check_fast(res["f1"][0][0], res["s1"][0][0], 0.8, "1M fast (tested walk) vs strict, step 1")
    # after two steps ulp-level position differences may move lattice particles across cell faces, so only
    # the continuous fields are compared (2x the single-step gates)
    a, b = res["f1"][1], res["s1"][1]
    assert rel_err(a["rho"], b["rho"]) <= 2 * TOL_RHO
    assert np.abs(a["pos"].astype(np.float64) - b["pos"]).max() <= 2 * TOL_POS * 0.8


@pytest.mark.parametrize("scene", ["dam_break_1M", "fluid_drop_1M", "dam_break_10M"])
def test_full_size_step_vs_oracle(pkg, po, scene):
    """BASELINE.json configs[1], configs[2] and the headline configs[3] workload (10.7 M particles, one GPU) at FULL size
    against the oracle itself (the C restatement, OpenMP: ~1 s per step at 1 M, ~15 s at 10.7 M on the box's host cores):
    keys, sorted permutation and neighbour counts bit-exact in both math modes, strict fields bit-exact, the default fast
    path (fluid drop: the R = 5 kernels at size) inside the single-step gates."""
    from sph_b200 import scenes
    if scene == "dam_break_10M":
        import psutil
        if psutil.virtual_memory().available < 48 * 2 ** 30:
            pytest.skip("the oracle's neighbour lists of 10.7 M particles need ~12 GB, the scene arrays and dumps some more")
    pos, mass, prm, dt = scenes.make_scene(scene)
    n = pos.shape[0]
    ora = po.Engine("port", n); ora.initialize(prm); ora.add_particles(pos, None, mass)
    keys = ora.keys()
    ora.step(dt)
    want, counts = ora.state(), ora.neighbor_counts()
    ora.close()
    refine = int(max(1, min(6, round(float(prm["neighbor_search_radius"]) / scenes.SCENES[scene][1]))))
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    for strict in (True, False):
        ctx = make_ctx(pkg, n, prm, strict=strict, OPT_GRID_REFINE=refine)
        ctx.upload(pos, None, mass)
        ctx.step(dt)
        d, got = ctx.debug_dump(), ctx.download()
        what = f"{scene} {'strict' if strict else 'fast'}"
        assert_bits(d["keys"], keys, f"{what} keys")
        assert_bits(d["perm"], stable_perm(keys), f"{what} sorted permutation")
        assert_bits(d["nbr_count"], counts, f"{what} neighbour counts")
        assert ctx.stats()["max_neighbors"] == int(counts.max())
        if strict:
            for f in ("rho", "P", "acc", "pos", "vel"):
                assert_bits(got[f], want[f], f"{what} {f}")
        else:
            check_fast(got, want, L, what)
        ctx.close()


def test_sparse_fluid_drop_multi_step(pkg, po):
    """S3 family (drop in a much larger AABB): the cell table is sized from the particles' bounding box, which
    the integrate kernel tracks on the device; 6 steps strict stay bit-exact, fast stays inside 6x the gates."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.fluid_drop_scene(0.004)
    n = pos.shape[0]
    ora = po.Engine("port", n); ora.initialize(prm); ora.add_particles(pos, None, mass)
    cs = make_ctx(pkg, n, prm, strict=True); cs.upload(pos, None, mass)
    cf = make_ctx(pkg, n, prm, strict=False); cf.upload(pos, None, mass)
    for k in range(6):
        keys = ora.keys()
        ora.step(dt); cs.step(dt); cf.step(dt)
        d = cs.debug_dump()
        assert_bits(d["keys"], keys, f"step {k} keys")
        assert_bits(d["nbr_count"], ora.neighbor_counts(), f"step {k} counts")
        if k == 0:
            df = cf.debug_dump()
            assert_bits(df["perm"], stable_perm(keys), "fast-mode reference-order permutation")
            assert_bits(df["nbr_count"], ora.neighbor_counts(), "fast-mode counts")
    want, got, fast = ora.state(), cs.download(), cf.download()
    for f in ("rho", "P", "acc", "pos", "vel"):
        assert_bits(got[f], want[f], f"drop strict {f}")
    assert rel_err(fast["rho"], want["rho"]) <= 6 * TOL_RHO
    assert np.abs(fast["pos"].astype(np.float64) - want["pos"]).max() <= 6 * TOL_POS * 2.0
    ora.close(); cs.close(); cf.close()


@pytest.mark.parametrize("refine", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("name", ["cloud600", "cloud600_truncated_support", "dam_break_13k_tame"])
def test_fast_mode_grid_refinements(pkg, name, refine):
    """Default fast path (bitmask hand-off) on every internal grid: refine 2/3 use 64-bit column masks (pair_mask_wide.cu),
    refine 4..6 paired 16-bit ones (pair_mask.cu), refine 1 falls back to the tested walk.  Neighbour sets must not depend on the grid; fields stay in the gates."""
    g = load_golden(name)
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    ctx = make_ctx(pkg, n, prm, strict=False, OPT_PAIR_KERNEL=2, OPT_GRID_REFINE=refine)
    ctx.upload(g["pos"], g["vel"], g["mass"])
    k0 = golden_steps(g)[0]
    for k in range(k0 + 1):
        ctx.step(float(g["dts"][k]))
    d = ctx.debug_dump()
    if k0 == 0:
        assert_bits(d["perm"], stable_perm(g["s0_keys"]), f"{name} refine {refine} permutation")
        assert_bits(d["nbr_count"], g["s0_counts"], f"{name} refine {refine} counts")
    got = ctx.download()
    want = {f: g[f"s{k0}_{f}"] for f in ("rho", "P", "acc", "pos", "vel")}
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    check_fast(got, want, L, f"{name} refine {refine}")
    ctx.close()


@pytest.mark.parametrize("refine", [2, 4])
def test_fast_mode_mask_overflow_falls_back(pkg, po, refine):
    """A collapsed state: far more candidates per cell column than a column mask holds (64 / 16 bits).  Those particles
    take the tested walk in the force pass; counts stay bit-exact and the fields stay inside the gates.  Half of the
    cloud is dilute, so both paths run inside the same warps."""
    rng = np.random.default_rng(11)
    prm = dict(pkg.DEFAULT_PARAMS)
    prm.update(smoothing_length=0.05, neighbor_search_radius=0.1, gas_constant=1e-3, viscosity=1e-6, particle_mass=1e-4,
               xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0)
    dense = rng.uniform(-0.06, 0.06, size=(3000, 3))
    dilute = rng.uniform(-0.9, 0.9, size=(3000, 3))
    pos = np.concatenate([dense, dilute]).astype(np.float32)
    vel = rng.normal(0, 0.1, size=pos.shape).astype(np.float32)
    mass = np.full(len(pos), prm["particle_mass"], np.float32)
    ora = po.Engine("port", len(pos)); ora.initialize(prm); ora.add_particles(pos, vel, mass)
    ctx = make_ctx(pkg, len(pos), prm, strict=False, OPT_PAIR_KERNEL=2, OPT_GRID_REFINE=refine)
    ctx.upload(pos, vel, mass)
    ora.step(1e-4); ctx.step(1e-4)
    assert_bits(ctx.debug_dump()["nbr_count"], ora.neighbor_counts(), "overflow counts")
    assert ora.neighbor_counts().max() > 2000
    check_fast(ctx.download(), ora.state(), 2.0, f"overflow refine {refine}")
    ctx.close(); ora.close()


@pytest.mark.parametrize("lanes", [2, 4, 8])
@pytest.mark.parametrize("name", ["micro_pair", "micro_coincident", "micro_lattice27", "cloud600", "cloud600_truncated_support",
                                  "dam_break_13k_tame"])
def test_lanes_per_particle(pkg, name, lanes):
    """SPHB_OPT_LANES_PER_PARTICLE (pair_split.cu, for small scenes): the column groups of a particle dealt to L lanes,
    partial sums joined by shuffles.  Neighbour sets are those of the one-lane kernels (counts bit-exact against the
    reference golden), fields inside the fast-mode gates against the reference and within summation-order noise of the
    one-lane result; several steps stay inside the N-step gates."""
    g = load_golden(name)
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    k0 = golden_steps(g)[0]
    outs = {}
    for ln in (1, lanes):
        ctx = make_ctx(pkg, n, prm, strict=False, OPT_LANES_PER_PARTICLE=ln, OPT_PAIR_MODE=0)
        assert ctx.get_option(pkg.capi.OPT_LANES_PER_PARTICLE) == ln
        ctx.upload(g["pos"], g["vel"], g["mass"])
        for k in range(k0 + 1):
            ctx.step(float(g["dts"][k]))
        outs[ln] = (ctx.download(), ctx.debug_dump(), ctx.stats()["max_neighbors"])
        ctx.close()
    got, dbg, mx = outs[lanes]
    if k0 == 0:
        assert_bits(dbg["nbr_count"], g["s0_counts"], f"{name} counts with {lanes} lanes per particle")
    assert_bits(dbg["nbr_count"], outs[1][1]["nbr_count"], "counts vs one lane")
    assert mx == outs[1][2]
    check_fast(got, {f: g[f"s{k0}_{f}"] for f in ("rho", "P", "acc", "pos", "vel")}, L, f"{name} lanes={lanes}")
    assert rel_err(got["rho"], outs[1][0]["rho"]) <= 2e-6
    if n > 100:
        assert not np.array_equal(got["acc"], outs[1][0]["acc"]), "the split kernels did not run (identical bits)"
    amax = np.abs(outs[1][0]["acc"]).max()
    if amax > 0:
        assert np.abs(got["acc"].astype(np.float64) - outs[1][0]["acc"]).max() <= 2e-5 * amax


@pytest.mark.parametrize("lanes", [4])
def test_lanes_per_particle_overflow_and_steps(pkg, po, lanes):
    """The mask-overflow fallback (collapsed cloud, > 2 000 neighbours) and ten consecutive steps with 4 lanes per particle."""
    rng = np.random.default_rng(11)
    prm = dict(pkg.DEFAULT_PARAMS)
    prm.update(smoothing_length=0.05, neighbor_search_radius=0.1, gas_constant=1e-3, viscosity=1e-6, particle_mass=1e-4,
               xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0)
    pos = np.concatenate([rng.uniform(-0.06, 0.06, size=(3000, 3)), rng.uniform(-0.9, 0.9, size=(3000, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.1, size=pos.shape).astype(np.float32)
    mass = np.full(len(pos), prm["particle_mass"], np.float32)
    ora = po.Engine("port", len(pos)); ora.initialize(prm); ora.add_particles(pos, vel, mass)
    ctx = make_ctx(pkg, len(pos), prm, strict=False, OPT_LANES_PER_PARTICLE=lanes)
    ctx.upload(pos, vel, mass)
    ora.step(1e-4); ctx.step(1e-4)
    assert_bits(ctx.debug_dump()["nbr_count"], ora.neighbor_counts(), "overflow counts")
    check_fast(ctx.download(), ora.state(), 2.0, "overflow with 4 lanes per particle")
    ctx.close(); ora.close()
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    ora = po.Engine("port", len(pos)); ora.initialize(prm); ora.add_particles(pos, None, mass)
    ctx = make_ctx(pkg, len(pos), prm, strict=False, OPT_LANES_PER_PARTICLE=lanes)
    ctx.upload(pos, None, mass)
    for _ in range(10):
        ora.step(dt); ctx.step(dt)
    want, fast = ora.state(), ctx.download()
    assert rel_err(fast["rho"], want["rho"]) <= 10 * TOL_RHO
    assert np.abs(fast["pos"].astype(np.float64) - want["pos"]).max() <= 10 * TOL_POS * 0.8
    assert np.abs(fast["acc"].astype(np.float64) - want["acc"]).max() <= 10 * TOL_ACC * np.abs(want["acc"]).max()
    ctx.close(); ora.close()


def test_fast_mode_layout_major_axis(pkg):
    """SPHB_OPT_LAYOUT_MAJOR only permutes the device's cell order: neighbour sets, keys and the reference-order
    permutation are identical for every major axis, fields agree within the fast-mode gates, and the same axis twice
    is bit-identical (what the slab == single-context parity relies on)."""
    g = load_golden("dam_break_13k_tame")
    prm = params_from(g["params"])
    n = g["pos"].shape[0]
    runs = []
    for major in (0, 1, 2, 2):
        ctx = make_ctx(pkg, n, prm, strict=False, OPT_LAYOUT_MAJOR=major)
        ctx.upload(g["pos"], g["vel"], g["mass"])
        ctx.step(float(g["dts"][0]))
        runs.append((ctx.download(), ctx.debug_dump()))
        ctx.close()
    L = float(max(prm["xmax"] - prm["xmin"], prm["ymax"] - prm["ymin"], prm["zmax"] - prm["zmin"]))
    for out, dbg in runs[1:]:
        for f in ("keys", "perm", "nbr_count"):
            assert_bits(dbg[f], runs[0][1][f], f"layout major: {f}")
        check_fast(out, runs[0][0], L, "layout major")
    for f in ("rho", "P", "acc", "pos", "vel"):
        assert_bits(runs[2][0][f], runs[3][0][f], f"same layout twice: {f}")


def test_sixty_steps_conservation_diagnostics_fast_vs_oracle(pkg, po):
    """SURVEY.md §8c, >= 50 steps: trajectories diverge chaotically (the reference's own fast-math build drifts 1e-3
    in density after 50 steps), so the gate is on the conservation diagnostics and the neighbour statistics: the
    default fast path and the oracle agree to 1 % after 60 steps of the 13k tame dam break."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    n = pos.shape[0]
    ora = po.Engine("port", n); ora.initialize(prm); ora.add_particles(pos, None, mass)
    cf = make_ctx(pkg, n, prm, strict=False); cf.upload(pos, None, mass)
    for _ in range(60):
        ora.step(dt); cf.step(dt)
    sum_rho, ke, vmax = cf.diagnostics()
    h = float(np.float32(prm["smoothing_length"]))
    want = ora.state()
    assert abs(sum_rho * h ** 3 - ora.total_mass()) <= 1e-2 * abs(ora.total_mass())
    assert abs(ke - ora.total_energy()) <= 1e-2 * abs(ora.total_energy())
    assert abs(vmax - np.sqrt((want["vel"].astype(np.float64) ** 2).sum(1)).max()) <= 1e-2 * vmax
    cnt = cf.debug_dump()["nbr_count"].astype(np.float64)
    ref = ora.neighbor_counts().astype(np.float64)
    assert abs(cnt.mean() - ref.mean()) <= 1e-2 * ref.mean()
    assert abs(cnt.max() - ref.max()) <= 0.05 * ref.max()
    t, steps = cf.get_time()
    assert steps == 60
    ora.close(); cf.close()


def _staged_vs_global(pkg, pos, vel, mass, prm, dts, refine, what):
    """Runs the same steps with the staged pair kernels (pair_stage.cu, default) and with the per-lane global-memory
    kernels (pair_mask.cu): every output must be the same bits."""
    outs = []
    for staged in (1, 0, 2):
        ctx = make_ctx(pkg, len(pos), prm, strict=False, OPT_PAIR_KERNEL=2, OPT_GRID_REFINE=refine, OPT_PAIR_MODE=staged)
        ctx.upload(pos, vel, mass)
        for dt in dts:
            ctx.step(float(dt))
        outs.append((ctx.download(), ctx.debug_dump(), ctx.stats()["max_neighbors"]))
        ctx.close()
    for other in outs[1:]:
        for f in ("rho", "P", "acc", "pos", "vel"):
            assert_bits(outs[0][0][f], other[0][f], f"{what}: staged vs global {f}")
        assert_bits(outs[0][1]["nbr_count"], other[1]["nbr_count"], f"{what}: staged vs global counts")
        assert outs[0][2] == other[2]
    return outs[0]


@pytest.mark.parametrize("refine", [4, 5, 6])
@pytest.mark.parametrize("name", ["micro_pair", "micro_coincident", "cloud600", "cloud600_truncated_support", "dam_break_13k_tame",
                                  "fluid_drop_default_pref"])
def test_staged_equals_global(pkg, name, refine):
    """The shared-memory staged kernels only change where a candidate is read from (a tile-wide stage filled by bulk
    asynchronous copies instead of a private global load): results are bit-identical to pair_mask.cu on every fixture,
    including the truncated-support kernels, tiny tiles and scenes whose tiles span sparse cells (unstaged groups)."""
    g = load_golden(name)
    prm = params_from(g["params"])
    dts = [float(d) for d in g["dts"][:3]]
    out, dbg, _ = _staged_vs_global(pkg, g["pos"], g["vel"], g["mass"], prm, dts[:1], refine, f"{name} refine {refine}")
    if "s0_counts" in g:
        assert_bits(dbg["nbr_count"], g["s0_counts"], f"{name} refine {refine}: staged counts vs reference")
    _staged_vs_global(pkg, g["pos"], g["vel"], g["mass"], prm, dts, refine, f"{name} refine {refine}, {len(dts)} steps")


def test_staged_equals_global_mask_overflow(pkg):
    """Collapsed cloud (> 2000 neighbours): most candidate ranges exceed a stage, so tiles mix staged and unstaged
    groups, and most columns overflow their 16-bit masks."""
    rng = np.random.default_rng(11)
    prm = dict(pkg.DEFAULT_PARAMS)
    prm.update(smoothing_length=0.05, neighbor_search_radius=0.1, gas_constant=1e-3, viscosity=1e-6, particle_mass=1e-4,
               xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0)
    pos = np.concatenate([rng.uniform(-0.06, 0.06, size=(3000, 3)), rng.uniform(-0.9, 0.9, size=(3000, 3))]).astype(np.float32)
    vel = rng.normal(0, 0.1, size=pos.shape).astype(np.float32)
    mass = np.full(len(pos), prm["particle_mass"], np.float32)
    _staged_vs_global(pkg, pos, vel, mass, prm, [1e-4, 1e-4], 4, "overflow cloud")


def test_staged_equals_global_1M(pkg):
    """BASELINE.json configs[1] at full size, three steps (the lattice order is gone after the first)."""
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.make_scene("dam_break_1M")
    _staged_vs_global(pkg, pos, None, mass, prm, [dt] * 3, 4, "dam_break_1M")


def test_download_between_upload_and_step(pkg):
    """Density, pressure and acceleration belong to the last step: after a new upload (another particle set, another
    order) they must not be un-permuted with the new ids.  Until the next step they read as zeros."""
    g = load_golden("cloud600")
    prm = params_from(g["params"])
    ctx = make_ctx(pkg, 600, prm, strict=True)
    ctx.upload(g["pos"], g["vel"], g["mass"])
    ctx.step(float(g["dts"][0]))
    assert np.abs(ctx.download()["rho"]).min() > 0
    ctx.upload(g["pos"][::-1].copy(), None, g["mass"][::-1].copy())    # same particles, reversed ids, no step yet
    s = ctx.download()
    assert_bits(s["pos"], g["pos"][::-1].copy(), "positions of the new upload")
    assert not s["rho"].any() and not s["P"].any() and not s["acc"].any()
    assert ctx.diagnostics()[0] == 0.0
    ctx.step(float(g["dts"][0]))
    assert rel_err(ctx.download()["rho"], g["s0_rho"][::-1]) <= 1e-6    # (summation order follows the ids: last bits may differ)
    ctx.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("strict", [True, False])
def test_two_cells_apart_tie_at_cell_faces(pkg, po, axis, strict):
    """fp32 edge of the 27-cell walk with cell = neighbor_search_radius: a = 0.029999996 lies in cell 0, b = 0.059999995
    in cell 2, yet (b - a)^2 <= fl(nsr nsr) — the reference's two-ring query (spatial_hash.cpp:35) finds the pair.  The
    same on the negative side, with bystanders around so that the cells are not trivially empty."""
    prm = dict(pkg.DEFAULT_PARAMS)
    prm.update(smoothing_length=0.015, neighbor_search_radius=0.03, gas_constant=1e-3, viscosity=1e-6, particle_mass=1e-4,
               xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0)
    a, b = np.float32(0.029999996), np.float32(0.059999995)
    base = np.zeros((6, 3), np.float32)
    base[:, (axis + 1) % 3] = 0.0151; base[:, (axis + 2) % 3] = 0.0449
    base[0, axis], base[1, axis] = a, b
    base[2, axis], base[3, axis] = -a, -b
    base[4, axis], base[5, axis] = 0.0451, 0.0149
    rng = np.random.default_rng(5)
    pos = np.concatenate([base, rng.uniform(-0.12, 0.12, size=(400, 3)).astype(np.float32)])
    mass = np.full(len(pos), prm["particle_mass"], np.float32)
    ora = po.Engine("port", len(pos)); ora.initialize(prm); ora.add_particles(pos, None, mass)
    ora.step(1e-4)
    cnt = ora.neighbor_counts()
    d = pos[1, axis] - pos[0, axis]
    assert np.float32(d) * np.float32(d) <= np.float32(0.03) * np.float32(0.03), "the constructed pair must be a neighbour pair"
    inv = np.float32(1.0) / np.float32(0.03)
    assert np.floor(a * inv) == 0.0 and np.floor(b * inv) == 2.0, "... whose cells are two apart"
    for refine in ((1,) if strict else (1, 2, 4)):
        ctx = make_ctx(pkg, len(pos), prm, strict=strict, OPT_GRID_REFINE=refine)
        ctx.upload(pos, None, mass)
        ctx.step(1e-4)
        assert_bits(ctx.debug_dump()["nbr_count"], cnt, f"axis {axis} refine {refine}: neighbour counts with the face tie")
        if strict:
            got, want = ctx.download(), ora.state()
            for f in ("rho", "acc", "pos"):
                assert_bits(got[f], want[f], f"face tie {f}")
        ctx.close()
    ora.close()
