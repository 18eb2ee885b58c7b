"""The drop-in boundary: pybind11 module `sph` (reference python/bindings.cpp surface) over the host shell.

CPU part: import, names, host-side generators/containers, loud failure without a GPU.
GPU part: sph.Simulator driven exactly as a reference user would, checked against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_bits, rel_err

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def sph(pkg):
    import subprocess
    pydir = ROOT / "sph-particle-simulator_b200" / "python"
    if not list(pydir.glob("sph.*.so")):
        subprocess.run(["bash", str(ROOT / "sph-particle-simulator_b200" / "host" / "build.sh")], check=True)
    sys.path.insert(0, str(pydir))
    import sph
    return sph


SIM_METHODS = ["initialize", "initialize_dam_break", "initialize_fluid_drop", "initialize_granular_flow", "add_particles",
               "clear_particles", "step", "run_steps", "get_particles", "get_parameters", "get_current_time", "get_step_count",
               "set_parameters", "set_gravity", "set_viscosity", "set_smoothing_length", "set_boundaries",
               "get_performance_stats", "reset_performance_stats", "compute_conservation_errors", "get_total_mass",
               "get_total_energy", "validate_simulation", "is_initialized", "get_positions", "get_velocities",
               "get_densities", "get_pressures"]


def test_module_surface(sph):
    # names of reference python/bindings.cpp:8-166
    for name in ("SPHParameters", "ParticleType", "Particle", "ParticleSystem", "PerformanceStats", "Simulator",
                 "create_fluid_block", "create_boundary_box"):
        assert hasattr(sph, name), name
    assert sph.__version__ == "1.0.0"
    for meth in SIM_METHODS:
        assert hasattr(sph.Simulator, meth), meth
    p = sph.SPHParameters()
    assert (p.rest_density, p.gas_constant, p.particle_mass) == (1000.0, 2000.0, pytest.approx(0.001))
    assert p.damping == pytest.approx(0.99) and p.gravity == pytest.approx(-9.81)
    assert not hasattr(p, "CFL_factor") and not hasattr(p, "bounds")        # Q20: 8 fields only
    assert [t.name for t in (sph.ParticleType.FLUID, sph.ParticleType.BOUNDARY, sph.ParticleType.SOLID)] == ["FLUID", "BOUNDARY", "SOLID"]
    q = sph.Particle()
    assert q.mass == 1.0 and q.id == -1 and q.temperature == pytest.approx(293.15) and q.position == (0.0, 0.0, 0.0)
    q2 = sph.Particle((1, 2, 3), mass=0.5, type=sph.ParticleType.SOLID)
    assert q2.position == (1.0, 2.0, 3.0) and q2.mass == 0.5


def test_generators_match_reference(sph, po):
    for args in [((0, 0.3, 0), (0.4, 0.6, 0.8), 0.02), ((0.1, -0.2, 0.3), (0.37, 0.61, 0.83), 0.03)]:
        c, s, dx = (np.array(args[0], np.float32), np.array(args[1], np.float32), args[2])
        blk = sph.create_fluid_block(c, s, dx, mass=0.25)
        ref, _ = po.gen_fluid_block(args[0], args[1], dx, 0.25, kind="port")
        assert_bits(np.array([p.position for p in blk], np.float32), ref, "create_fluid_block")
        assert blk[0].mass == 0.25 and blk[0].type == sph.ParticleType.FLUID
        box = sph.create_boundary_box(c, s, dx)
        ref, _ = po.gen_boundary_box(args[0], args[1], dx, 1.0, kind="port")
        assert_bits(np.array([p.position for p in box], np.float32), ref, "create_boundary_box")
        assert box[0].type == sph.ParticleType.BOUNDARY and box[0].mass == 1.0
    with pytest.raises(ValueError):
        sph.create_fluid_block(np.zeros(2, np.float32), np.ones(3, np.float32), 0.1)


def test_no_gpu_fails_loudly(sph, has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError) as ei:
        sph.Simulator(max_particles=100)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(TypeError):
        sph.Simulator(particles=100)          # Q20: the README's keyword is wrong in the reference too


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_simulator_default_dam_break_vs_oracle(sph, po):
    """A reference user's session: Simulator(), initialize_dam_break(), step(dt) — default (exploding)
    parameters, strict math: bit-exact against the oracle for two steps."""
    sim = sph.Simulator(max_particles=100000)
    sim.set_math_mode(0)
    assert not sim.is_initialized()
    sim.step(0.001)                                   # no-op before initialisation
    assert sim.get_step_count() == 0
    sim.initialize_dam_break()
    ora = po.Engine("port", 100000); ora.initialize_dam_break()
    assert sim.is_initialized() and sim.get_particles().size() == ora.size == 84800
    assert sim.get_particles().capacity() == 100000
    assert_bits(sim.get_positions(), ora.state()["pos"], "initial positions")
    for _ in range(2):
        sim.step(0.001); ora.step(0.001)
    want = ora.state()
    assert sim.get_step_count() == 2 and np.float32(sim.get_current_time()) == np.float32(ora.time)
    pos, vel = sim.get_positions(), sim.get_velocities()
    assert pos.shape == (84800, 3) and pos.dtype == np.float32
    assert_bits(pos, want["pos"], "positions"); assert_bits(vel, want["vel"], "velocities")
    rho, P = sim.get_densities(), sim.get_pressures()
    assert rho.shape == (100000,) and P.shape == (100000,)          # Q13: capacity-length
    assert_bits(rho[:84800], want["rho"], "densities"); assert_bits(P[:84800], want["P"], "pressures")
    assert not rho[84800:].any()
    assert_bits(sim.get_accelerations(), want["acc"], "accelerations")
    ps = sim.get_particles()                                         # live AoS, refreshed lazily
    assert_bits(ps.get_positions(), want["pos"], "ParticleSystem positions")
    assert_bits(ps.get_densities(), want["rho"], "ParticleSystem densities")
    st = sim.get_performance_stats()
    assert st.max_neighbors == ora.stats()["max_neighbors"] and st.total_neighbor_queries == 2 * 84800
    assert st.force_computation_time > 0 and st.total_time >= st.force_computation_time
    m_err, e_err = sim.compute_conservation_errors()
    assert e_err == 0.0 and rel_err(m_err, ora.conservation_errors()[0]) < 1e-4
    assert rel_err(sim.get_total_mass(), ora.total_mass()) < 1e-4
    assert rel_err(sim.get_total_energy(), ora.total_energy()) < 1e-4
    sim.reset_performance_stats()
    assert sim.get_performance_stats().max_neighbors == 0
    ora.close()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0", "0,0,0"])
def test_simulator_public_api_scene_and_quirks(sph, po, devices, monkeypatch):
    """Scene built through the public API (create_boundary_box + create_fluid_block + add_particles,
    set_smoothing_length, set_boundaries), adaptive stepping, capacity truncation, re-initialisation.
    devices: the same engine owning several slabs (SPHB_DEVICES → sphb_create_multi behind the host shell; ordinal 0
    repeated so that a one-GPU box runs it) must give the same bits."""
    if devices:
        monkeypatch.setenv("SPHB_DEVICES", devices)
    else:
        monkeypatch.delenv("SPHB_DEVICES", raising=False)
    dx = 0.02
    m = 1.5 * 1000 * dx ** 3
    prm = sph.SPHParameters()
    prm.gas_constant = 100.0 * m / 1000.0; prm.viscosity = 1e-3 * m; prm.particle_mass = m; prm.damping = 0.999
    prm.timestep = 0.25 * 2 * dx / 10.0
    sim = sph.Simulator(max_particles=13000)                         # 13 200 generated → 200 dropped (Q14)
    assert sim.device_count() == (len(devices.split(",")) if devices else 1)
    sim.set_math_mode(0)
    sim.initialize(prm)
    c, s = np.array([0, 0.3, 0], np.float32), np.array([0.4, 0.6, 0.8], np.float32)
    sim.add_particles(sph.create_boundary_box(c, s, dx, m))
    sim.add_particles(sph.create_fluid_block(np.array([-0.1, 0.2, 0], np.float32), np.array([0.2, 0.4, 0.8], np.float32), dx, m))
    sim.set_smoothing_length(2 * dx)
    sim.set_boundaries(-0.2, 0.2, 0.0, 0.6, -0.4, 0.4)
    assert sim.get_particles().size() == 13000

    ora = po.Engine("port", 13000)
    op = dict(po.default_params("port"))
    op.update(gas_constant=prm.gas_constant, viscosity=prm.viscosity, particle_mass=prm.particle_mass, damping=prm.damping,
              timestep=prm.timestep)
    ora.initialize(op)
    bp, bm = po.gen_boundary_box((0, 0.3, 0), (0.4, 0.6, 0.8), dx, m, kind="port")
    fp, fm = po.gen_fluid_block((-0.1, 0.2, 0), (0.2, 0.4, 0.8), dx, m, kind="port")
    ora.add_particles(bp, None, bm); ora.add_particles(fp, None, fm)
    ora.set_smoothing_length(2 * dx); ora.set_boundaries(-0.2, 0.2, 0.0, 0.6, -0.4, 0.4)
    assert ora.size == 13000
    assert np.float32(sim.compute_cfl_timestep()) == np.float32(ora.cfl_timestep())
    sim.run_steps(3)                                                  # adaptive by default (Q12)
    ora.run_steps(3, True)
    sim.run_steps(2, adaptive_timestep=False); ora.run_steps(2, False)
    assert sim.get_step_count() == 5 and np.float32(sim.get_current_time()) == np.float32(ora.time)
    want = ora.state()
    assert_bits(sim.get_positions(), want["pos"], "positions"); assert_bits(sim.get_velocities(), want["vel"], "velocities")
    assert_bits(sim.get_densities()[:13000], want["rho"], "densities")
    # appending after stepping: host AoS is refreshed first, ids continue (capacity reached → dropped)
    sim.add_particles([sph.Particle((0, 0.3, 0), m)])
    assert sim.get_particles().size() == 13000
    # re-initialising keeps the clock (Q15); clear_particles resets it
    t = sim.get_current_time()
    sim.initialize_fluid_drop()
    assert sim.get_step_count() == 5 and sim.get_current_time() == t and sim.get_particles().size() == 8144
    sim.step(0.0005)
    assert sim.get_step_count() == 6
    sim.clear_particles()
    assert sim.get_step_count() == 0 and sim.get_current_time() == 0.0 and sim.get_particles().size() == 0
    sim.step(0.001)
    assert sim.get_step_count() == 0
    ora.close()


@pytest.mark.gpu
def test_reference_benchmark_driver_runs_on_the_gpu_engine(tmp_path):
    """The reference's own benchmarks/performance_test.cpp, compiled unmodified against the host shell
    (dropin/build.sh, prebuilt in the build container), runs and writes its CSV schema."""
    import subprocess
    exe = ROOT / "dropin" / "_ref" / "performance_test"
    if not exe.exists():
        pytest.skip("dropin/_ref/performance_test not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "1000", "5000"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Results for 1000 particles" in out.stdout and "Results for 5000 particles" in out.stdout
    csv = (tmp_path / "benchmark_results.csv").read_text().splitlines()
    assert csv[0].startswith("particles,avg_fps,min_fps,max_fps,avg_ms_frame,neighbor_search_pct") and len(csv) == 3
    fps = float(csv[2].split(",")[1])
    assert fps > 100.0, f"5000-particle config ran at {fps} FPS"


@pytest.mark.gpu
def test_headless_dam_break_example(tmp_path):
    """examples/dam_break_headless.py: the reference example's parameters, loop and CSV columns on the GPU engine
    (BASELINE.json configs[0]: capacity 20 000 → 20 000 wall particles, 0 fluid)."""
    import subprocess
    out = subprocess.run([sys.executable, str(ROOT / "examples" / "dam_break_headless.py"), "10000", "0.012", str(tmp_path / "d.csv")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "Simulation initialized with 20000 particles" in out.stdout
    rows = (tmp_path / "d.csv").read_text().splitlines()
    assert rows[0] == "time,particles,mass_error,energy,total_energy,avg_density,max_velocity" and len(rows) >= 2
    assert rows[1].split(",")[1] == "20000"


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0"])
def test_headless_example_rows_match_the_reference(tmp_path, po, devices):
    """(devices = "0,0": the unchanged example program with SPHB_DEVICES exported — the engine owns two slabs.)
    The first five CSV rows of the example (one device reduction per report) against what the reference program
    computes from full arrays (examples/dam_break.cpp:132-166): compute_conservation_errors, get_total_energy, the mean
    of the CAPACITY-long density buffer (quirk Q13) and max |v| — on the reference engine itself (strict build or the C
    port), same parameters, same 50 steps.  The scene explodes (reference defaults): rows agree while finite and turn
    non-finite together."""
    import subprocess
    import os
    env = dict(os.environ)
    env.pop("SPHB_DEVICES", None)
    if devices:
        env["SPHB_DEVICES"] = devices
    out = subprocess.run([sys.executable, str(ROOT / "examples" / "dam_break_headless.py"), "10000", "0.05", str(tmp_path / "d.csv"), "--strict"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    rows = [r.split(",") for r in (tmp_path / "d.csv").read_text().splitlines()[1:]]
    assert len(rows) == 5
    prm = dict(po.DEFAULT_PARAMS) if hasattr(po, "DEFAULT_PARAMS") else None
    e = po.Engine("strict" if po.available("strict") else "port", 20000)
    p = e.get_parameters() if prm is None else prm
    p.update(rest_density=1000.0, gas_constant=2000.0, viscosity=0.001, smoothing_length=0.025, particle_mass=0.001, timestep=0.001,
             gravity=-9.81, damping=0.995)
    e.initialize(p)
    e.set_boundaries(-1.0, 1.0, -0.5, 1.5, -1.0, 1.0)
    e.initialize_dam_break()
    n = e.size
    k = 0
    max_steps = int(0.05 / float(np.float32(0.001)))   # 49: duration / (float)timestep truncated, as in the example and dam_break.cpp
    for step in range(max_steps):
        e.step(0.001)
        if step % 10 == 9 or step == max_steps - 1:
            st = e.state()
            want = [e.time, float(n), e.conservation_errors()[0], 0.0, e.total_energy(), float(e.densities_raw().astype(np.float64).mean()),
                    float(np.sqrt((st["vel"] * st["vel"]).sum(1, dtype=np.float32)).max())]   # glm::length in fp32, like dam_break.cpp:141
            got = [float(x) for x in rows[k]]
            for name, g_, w_ in zip("time particles mass_error energy total_energy avg_density max_velocity".split(), got, want):
                if np.isfinite(w_) and abs(w_) < 1e30:
                    assert abs(g_ - w_) <= 2e-3 * abs(w_) + 1e-12, f"row {k} {name}: {g_} vs reference {w_}"
                else:       # the exploded phase: fp32 sums overflow in the reference; huge or non-finite on both sides
                    assert not np.isfinite(g_) or abs(g_) >= 1e30, f"row {k} {name}: {g_} vs reference {w_}"
            k += 1
    e.close()


@pytest.mark.gpu
def test_reference_benchmark_dt_sequence_through_the_drop_in(sph):
    """benchmarks/performance_test.cpp:85-125 driven through the drop-in class exactly as the reference program does
    (SPHEngine(2 x 5000), its parameter block incl. h = (m / rho0)^(1/3), initialize, initialize_dam_break, step() with
    dt = 0): the adaptive dt sequence — params.timestep, then the CFL branch of the exploding defaults — equals the
    unmodified reference's (tests/golden/scalars.json cfl_cases.perf_test), time by time, bit for bit."""
    import hashlib, json
    case = json.loads((ROOT / "tests" / "golden" / "scalars.json").read_text())["cfl_cases"]["perf_test"]
    sim = sph.Simulator(max_particles=case["capacity"])
    sim.set_math_mode(0)
    p = sph.SPHParameters()
    p.rest_density = 1000.0; p.gas_constant = 2000.0; p.viscosity = 0.001; p.particle_mass = 0.001; p.timestep = 0.001; p.gravity = -9.81
    p.smoothing_length = float(np.float32(np.float32(np.float32(1.0) / np.float32(1000.0) * np.float32(0.001)) ** np.float32(1.0 / 3.0)))
    sim.initialize(p)
    sim.initialize_dam_break()
    assert sim.get_particles().size() == case["n"]
    for want_dt, want_t in zip(case["dts"], case["times"]):
        assert np.float32(sim.compute_cfl_timestep()) == np.float32(want_dt)
        sim.step()
        assert np.float32(sim.get_current_time()) == np.float32(want_t)
    assert hashlib.sha256(np.ascontiguousarray(sim.get_positions()).tobytes()).hexdigest() == case["final_pos_sha256"]
    d = sim.get_report_diagnostics()
    vel = sim.get_velocities().astype(np.float64)
    assert abs(d["max_velocity"] - np.sqrt((vel ** 2).sum(1)).max()) <= 1e-6 * d["max_velocity"]
    assert abs(d["average_density"] - sim.get_densities().astype(np.float64).mean()) <= 1e-5 * abs(d["average_density"])


@pytest.mark.gpu
def test_renderer_instance_records(sph, pkg):
    """The renderer feed (reference Renderer::update_particle_data, src/renderer.cpp:279-312): position, velocity and
    Particle::color of every particle in insertion order, from one export kernel — through the drop-in class (colours
    of the generators: wall vs fluid) and through the C ABI (default colour, device destination)."""
    import torch
    sim = sph.Simulator(max_particles=30000)
    sim.initialize(sph.SPHParameters())
    wall = sph.create_boundary_box(np.array([0, 0.3, 0], np.float32), np.array([0.4, 0.6, 0.8], np.float32), 0.04, 0.001)
    fluid = sph.create_fluid_block(np.array([-0.1, 0.2, 0], np.float32), np.array([0.2, 0.4, 0.8], np.float32), 0.04, 0.001)
    sim.add_particles(wall); sim.add_particles(fluid)
    sim.step(1e-4); sim.step(1e-4)
    rec = sim.get_instance_data()
    n = sim.get_particles().size()
    assert rec.shape == (n, 9)
    assert_bits(np.ascontiguousarray(rec[:, 0:3]), sim.get_positions(), "instance positions")
    assert_bits(np.ascontiguousarray(rec[:, 3:6]), sim.get_velocities(), "instance velocities")
    colors = np.array([[p.color[0], p.color[1], p.color[2]] for p in (list(wall) + list(fluid))], np.float32) if hasattr(wall[0], "color") and not isinstance(wall[0].color, (int, float)) else None
    if colors is not None:
        assert_bits(np.ascontiguousarray(rec[:, 6:9]), colors, "instance colours")
    assert len(np.unique(rec[:, 6:9], axis=0)) >= 1
    # C ABI: default colour, device destination (what a CUDA-mapped vertex buffer would be)
    g = np.random.default_rng(3)
    pos = g.uniform(-0.1, 0.1, size=(500, 3)).astype(np.float32)
    vel = g.normal(size=(500, 3)).astype(np.float32)
    ctx = pkg.Context(500, 0)
    prm = dict(pkg.DEFAULT_PARAMS); prm.update(xmin=-1, xmax=1, ymin=-1, ymax=1, zmin=-1, zmax=1)
    ctx.set_params(prm); ctx.upload(pos, vel, None)
    ctx.step(1e-4)
    host = ctx.export_instances()
    s = ctx.download()
    assert_bits(np.ascontiguousarray(host[:, 0:3]), s["pos"], "ABI instance positions")
    assert_bits(np.ascontiguousarray(host[:, 3:6]), s["vel"], "ABI instance velocities")
    assert (host[:, 6:9] == np.array([0.0, 0.5, 1.0], np.float32)).all()      # sph::Particle's default colour
    dst = torch.zeros((500, 9), dtype=torch.float32, device="cuda:0")
    rgb = g.uniform(size=(500, 3)).astype(np.float32)
    ctx.set_colors(rgb)
    ctx.export_instances(device_ptr=dst.data_ptr())
    ctx.synchronize() if hasattr(ctx, "synchronize") else torch.cuda.synchronize()
    dev = dst.cpu().numpy()
    assert_bits(np.ascontiguousarray(dev[:, 0:6]), np.ascontiguousarray(host[:, 0:6]), "device destination")
    assert_bits(np.ascontiguousarray(dev[:, 6:9]), rgb, "uploaded colours")
    ctx.close()
