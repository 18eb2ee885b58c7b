#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference → oracle/_ref/liboracle_strict.so, built by
`make -f oracle/Makefile ref`).  The reference ships no tests or fixtures of its own, so
these vectors — outputs of the reference compiled IEEE-strict (-O2 -ffp-contract=off) on small
deterministic inputs — are what pins the oracle and the CUDA path on the GPU box, where
/root/reference does not exist.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import pyoracle as po  # noqa: E402
import __graft_entry__ as graft  # noqa: E402

graft.load_package()
from sph_b200 import scenes  # noqa: E402

KIND = "strict"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_scene(params, pos, vel, mass, dts, capacity=None, keep_steps=None, kernel_type=0):
    """Step the reference; returns per-step dicts (state after the step + keys/counts of that step).
    kernel_type != 0: the reference's Wendland C2 (1) / Gaussian (2) class installed in the engine's kernel slot."""
    n = pos.shape[0]
    e = po.Engine(KIND, capacity or n)
    e.initialize(params)
    if kernel_type:
        e.set_kernel(kernel_type)
    e.add_particles(pos, vel, mass)
    out = []
    for k, dt in enumerate(dts):
        keys = e.keys()            # keys of the positions this step's neighbour build sees
        e.step(dt)
        s = e.state()
        if keep_steps is None or k in keep_steps:
            rec = {f"s{k}_{name}": s[name] for name in ("pos", "vel", "rho", "P", "acc")}
            rec[f"s{k}_keys"] = keys
            rec[f"s{k}_counts"] = e.neighbor_counts().astype(np.uint32)
            rec[f"s{k}_time"] = np.float32(e.time)
            out.append(rec)
    merged = {}
    for r in out:
        merged.update(r)
    merged["final_total_mass"] = np.float32(e.total_mass())
    merged["final_total_energy"] = np.float32(e.total_energy())
    merged["final_mass_error"] = np.float32(e.conservation_errors()[0])
    merged["final_max_neighbors"] = np.int64(e.stats()["max_neighbors"])
    e.close()
    return merged


def save(name, params, pos, vel, mass, dts, **arrays):
    p = HERE / f"{name}.npz"
    np.savez_compressed(p, params=po.pack_params(params), pos=pos.astype(np.float32),
                        vel=(np.zeros_like(pos) if vel is None else vel).astype(np.float32), mass=mass.astype(np.float32),
                        dts=np.asarray(dts, np.float32), **arrays)
    print(f"{p.name}: {p.stat().st_size / 1024:.0f} KiB, N={pos.shape[0]}, steps={len(dts)}")


def main():
    if not po.available(KIND):
        raise SystemExit("oracle/_ref/liboracle_strict.so missing: run `make -f oracle/Makefile ref` in the build container")
    rng = np.random.default_rng(1234)
    defaults = dict(po.default_params(KIND))

    # ---- 1. smoothing-kernel and cell-key known answers ------------------------------------------
    kat = {}
    e = po.Engine(KIND, 16)
    for h in (0.02, 0.025, 0.008):
        prm = dict(defaults); prm["smoothing_length"] = h
        e.initialize(prm)
        qs = np.array([0.0, 1e-7, 0.25, 0.5, 0.999999, 1.0, 1.000001, 1.5, 1.999, 2.0, 2.000001, 2.5], np.float32)
        dirs = np.array([[1, 0, 0], [0, 1, 0], [0.6, 0.0, 0.8], [-0.57735, 0.57735, -0.57735]], np.float32)
        r = (qs[:, None, None] * np.float32(h) * dirs[None, :, :]).reshape(-1, 3).astype(np.float32)
        kat[f"h{h}_r"] = r
        kat[f"h{h}_W"] = np.array([e.kernel_W(x) for x in r], np.float32)
        kat[f"h{h}_gradW"] = np.array([e.kernel_gradW(x) for x in r], np.float32)
        kat[f"h{h}_lapW"] = np.array([e.kernel_lapW(x) for x in r], np.float32)
    e.close()
    # keys: hand-checkable cells plus random positions at three cell sizes
    for cell in (0.04, 0.016, 0.05):
        prm = dict(defaults); prm["neighbor_search_radius"] = cell
        pts = np.concatenate([
            np.array([[0, 0, 0], [-1e-6, -1e-6, -1e-6], [1.0, -1.0, 0.12], [cell, 2 * cell, 3 * cell],
                      [-cell, -2 * cell, -3 * cell], [0.999999, 0.5, -0.999999]], np.float32),
            rng.uniform(-1.5, 1.5, size=(250, 3)).astype(np.float32)])
        e = po.Engine(KIND, pts.shape[0]); e.initialize(prm); e.add_particles(pts)
        kat[f"cell{cell}_pts"] = pts
        kat[f"cell{cell}_keys"] = e.keys()
        e.close()
    np.savez_compressed(HERE / "kat_kernel_keys.npz", **kat)
    print("kat_kernel_keys.npz", (HERE / "kat_kernel_keys.npz").stat().st_size // 1024, "KiB")

    # ---- 2. micro scenes (tame parameters at dx = 0.02) -------------------------------------------
    dx = 0.02
    tame = scenes.tame_params(dx, 2 * dx, scenes.DAM_BOUNDS)
    m = tame["particle_mass"]
    dt = tame["timestep"]
    # two particles inside one support radius
    pos = np.array([[0.0, 0.3, 0.0], [0.03, 0.31, -0.01]], np.float32)
    save("micro_pair", tame, pos, None, np.full(2, m, np.float32), [dt] * 3, **run_scene(tame, pos, None, np.full(2, m, np.float32), [dt] * 3))
    # three particles, two of them coincident (Q4: counted by density and viscosity, skipped by pressure)
    pos = np.array([[0.0, 0.3, 0.0], [0.0, 0.3, 0.0], [0.02, 0.3, 0.01]], np.float32)
    vel = np.array([[0.1, 0, 0], [-0.2, 0.05, 0], [0, 0, 0.3]], np.float32)
    mass = np.array([m, 2 * m, 0.5 * m], np.float32)
    save("micro_coincident", tame, pos, vel, mass, [dt] * 3, **run_scene(tame, pos, vel, mass, [dt] * 3))
    # single particle; particles that start outside the AABB (clamped by the first step)
    pos = np.array([[0.05, 0.2, 0.1]], np.float32)
    save("micro_single", tame, pos, None, np.full(1, m, np.float32), [dt] * 2, **run_scene(tame, pos, None, np.full(1, m, np.float32), [dt] * 2))
    pos = np.array([[-0.25, 0.3, 0.0], [0.3, -0.1, 0.5], [0.0, 0.7, -0.45], [0.0, 0.3, 0.0]], np.float32)
    vel = np.array([[1, 2, 3], [-1, -2, -3], [0.5, 0.5, 0.5], [0, 0, 0]], np.float32)
    save("micro_outside", tame, pos, vel, np.full(4, m, np.float32), [dt] * 2, **run_scene(tame, pos, vel, np.full(4, m, np.float32), [dt] * 2))
    # 3x3x3 lattice straddling the origin (sign change of the masked cell key on every axis)
    g = (np.arange(3, dtype=np.float32) - 1) * np.float32(dx)
    X, Y, Z = np.meshgrid(g, g + np.float32(0.3), g, indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
    save("micro_lattice27", tame, pos, None, np.full(27, m, np.float32), [dt] * 3, **run_scene(tame, pos, None, np.full(27, m, np.float32), [dt] * 3))
    # jittered cloud with per-particle masses and velocities, crossing x = 0 and z = 0
    n = 600
    pos = (rng.uniform(-1, 1, size=(n, 3)) * np.array([0.12, 0.12, 0.12]) + np.array([0.0, 0.3, 0.0])).astype(np.float32)
    vel = rng.normal(0, 0.5, size=(n, 3)).astype(np.float32)
    mass = (m * rng.uniform(0.5, 2.0, size=n)).astype(np.float32)
    save("cloud600", tame, pos, vel, mass, [dt] * 5, **run_scene(tame, pos, vel, mass, [dt] * 5, keep_steps={0, 4}))
    # nsr != 2h (Q1/Q2: examples/dam_break.cpp h = 0.025 with the default 0.04 cell truncates the support)
    prm = dict(tame); prm["smoothing_length"] = 0.025; prm["neighbor_search_radius"] = 0.04
    save("cloud600_truncated_support", prm, pos, vel, mass, [dt] * 2, **run_scene(prm, pos, vel, mass, [dt] * 2, keep_steps={1}))
    # cell twice the support (benchmarks/performance_test.cpp: h = 0.01 with the 0.04 default cell)
    prm = dict(tame); prm["smoothing_length"] = 0.01; prm["neighbor_search_radius"] = 0.04
    save("cloud600_wide_cell", prm, pos, vel, mass, [dt] * 2, **run_scene(prm, pos, vel, mass, [dt] * 2, keep_steps={1}))

    # ---- 3. scene-level vectors --------------------------------------------------------------------
    # S2-family dam break at dx = 0.02 (N = 13 200), tame parameters, 3 steps; keep steps 0 and 2
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    save("dam_break_13k_tame", prm, pos, None, mass, [dt] * 3, **run_scene(prm, pos, None, mass, [dt] * 3, keep_steps={0, 2}))
    # reference-default fluid drop (initialize_fluid_drop, N = 8144) with the reference's exploding defaults, 2 steps
    e = po.Engine(KIND, 10000); e.initialize_fluid_drop(); s0 = e.state(); prm = e.get_parameters(); e.close()
    save("fluid_drop_default_pref", prm, s0["pos"], None, s0["mass"], [0.001] * 2,
         **run_scene(prm, s0["pos"], None, s0["mass"], [0.001] * 2, keep_steps={1}))
    # S1: examples/dam_break 10000 — capacity 20 000 truncates the scene to wall particles only; example parameters
    s1 = dict(defaults)
    s1.update(smoothing_length=0.025, damping=0.995, xmin=-1.0, xmax=1.0, ymin=-0.5, ymax=1.5, zmin=-1.0, zmax=1.0)
    e = po.Engine(KIND, 20000); e.initialize(s1); e.initialize_dam_break(); s0 = e.state(); e.close()
    assert s0["pos"].shape[0] == 20000
    r = run_scene(s1, s0["pos"], None, s0["mass"], [0.001] * 2, capacity=20000, keep_steps={1})
    r = {k: v for k, v in r.items() if not k.endswith(("_vel", "_acc"))}     # keep the file small
    save("dam_break_example_10k_pref", s1, s0["pos"], None, s0["mass"], [0.001] * 2, **r)

    # ---- 4. adaptive timestep + scalar diagnostics --------------------------------------------------
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    e = po.Engine(KIND, pos.shape[0]); e.initialize(prm); e.add_particles(pos, None, mass)
    dts, times = [], []
    for _ in range(6):
        dts.append(e.cfl_timestep()); e.step(0.0); times.append(e.time)
    st = e.state()
    meta = {
        "adaptive_dts": [float(np.float32(x)) for x in dts], "adaptive_times": [float(np.float32(x)) for x in times],
        "adaptive_final_pos_sha256": sha(st["pos"]), "adaptive_final_rho_sha256": sha(st["rho"]),
        "adaptive_total_mass": float(np.float32(e.total_mass())), "adaptive_total_energy": float(np.float32(e.total_energy())),
    }
    e.close()

    # ---- 5. adaptive timestep where the CFL rule bites (compute_cfl_timestep, sph_engine.cpp:312-333) ----
    # The tame scene above never leaves dt = params.timestep.  Three runs whose dt comes from the other two branches:
    #   perf_test      the reference's own benchmark set-up (benchmarks/performance_test.cpp:85-125: capacity 2 x 5000,
    #                  h = (m / rho0)^(1/3) = 0.01 with the default 0.04 search radius, initialize_dam_break truncated to
    #                  10 000 wall particles, step() with dt = 0): the default parameters explode, so after the first step
    #                  both CFL terms are far below params.timestep
    #   fast_cloud     tame dam break with initial speeds up to 40: dt_cfl = CFL h / max|v| wins from the first step
    #   light_zero*    tame dam break with a light particle 0: dt_force = CFL sqrt(h / |a_0|) wins (every step / some steps)
    def cfl_terms(e, prm):
        st = e.state()
        vmax = float(np.sqrt((st["vel"].astype(np.float64) ** 2).sum(1)).max())
        a0 = float(np.sqrt((st["acc"][0].astype(np.float64) ** 2).sum()))
        h, cfl = float(prm["smoothing_length"]), float(prm["CFL_factor"])
        return cfl * h / (vmax + 1e-6), cfl * np.sqrt(h / (a0 + 1e-6))

    def adaptive_run(e, prm, steps):
        dts, times, branch = [], [], []
        for _ in range(steps):
            d_cfl, d_force = cfl_terms(e, prm)
            dt = e.cfl_timestep()
            branch.append("timestep" if np.float32(dt) == np.float32(prm["timestep"]) else ("cfl" if d_cfl < d_force else "force"))
            dts.append(dt); e.step(0.0); times.append(e.time)
        st = e.state()
        return {"dts": [float(np.float32(x)) for x in dts], "times": [float(np.float32(x)) for x in times], "branch": branch,
                "final_pos_sha256": sha(st["pos"]), "final_vel_sha256": sha(st["vel"]), "final_rho_sha256": sha(st["rho"])}

    cfl_cases = {}
    pt = dict(defaults)
    pt.update(rest_density=1000.0, gas_constant=2000.0, viscosity=0.001, particle_mass=0.001, timestep=0.001, gravity=-9.81)
    pt["smoothing_length"] = float(np.float32(np.float32(np.float32(1.0) / np.float32(1000.0) * np.float32(0.001)) ** np.float32(1.0 / 3.0)))
    e = po.Engine(KIND, 10000); e.initialize(pt); e.initialize_dam_break()
    cfl_cases["perf_test"] = {"capacity": 10000, "n": int(e.size), "params": {k: float(np.float32(v)) for k, v in e.get_parameters().items()},
                              **adaptive_run(e, e.get_parameters(), 5)}
    e.close()
    pos, mass, prm, dt = scenes.dam_break_scene(0.02)
    rng = np.random.default_rng(77)
    vel = (rng.normal(size=pos.shape) * 12.0).astype(np.float32)
    e = po.Engine(KIND, pos.shape[0]); e.initialize(prm); e.add_particles(pos, vel, mass)
    cfl_cases["fast_cloud"] = {"vel_seed": 77, "vel_sigma": 12.0, **adaptive_run(e, prm, 5)}
    e.close()
    # particle 0 made light: a = F / m_i (quirk Q6) gives it the largest acceleration while every speed stays moderate,
    # so dt_force = CFL sqrt(h / |a_0|) (particle 0 only, previous step's a: quirk Q10) undercuts dt_cfl
    for name, fac in (("light_zero", 1.0e-2), ("light_zero_mixed", 3.0e-3)):
        mass2 = mass.copy()
        mass2[0] *= np.float32(fac)
        e = po.Engine(KIND, pos.shape[0]); e.initialize(prm); e.add_particles(pos, None, mass2)
        cfl_cases[name] = {"mass0_factor": fac, **adaptive_run(e, prm, 5)}
        e.close()
    for name, c in cfl_cases.items():
        print(name, c["branch"], c["dts"])
    meta["cfl_cases"] = cfl_cases
    # generator hashes (scene parity without storing the arrays)
    gens = {}
    for name in ("dam_break_13k", "dam_break_85k", "dam_break_347k", "dam_break_1M", "fluid_drop_65k", "fluid_drop_1M"):
        fam, dxx = scenes.SCENES[name]
        if fam == "dam":
            p, _ = po.gen_dam_break((0.4, 0.6, 0.8), (0.2, 0.4, 0.8), dxx, 1.0, kind=KIND)
        else:
            p, _ = po.gen_fluid_drop((0.0, 0.5, 0.0), 0.1, dxx, 1.0, kind=KIND)
        gens[name] = {"n": int(p.shape[0]), "pos_sha256": sha(p)}
    meta["generators"] = gens
    e = po.Engine(KIND, 100000); e.initialize_dam_break(); meta["initialize_dam_break_n"] = e.size; meta["initialize_dam_break_sha256"] = sha(e.state()["pos"]); e.close()
    e = po.Engine(KIND, 100000); e.initialize_fluid_drop(); meta["initialize_fluid_drop_n"] = e.size; meta["initialize_fluid_drop_sha256"] = sha(e.state()["pos"]); e.close()
    e = po.Engine(KIND, 100000); e.initialize_granular_flow(); meta["initialize_granular_flow_n"] = e.size; meta["initialize_granular_flow_sha256"] = sha(e.state()["pos"]); e.close()
    (HERE / "scalars.json").write_text(json.dumps(meta, indent=1))
    print("scalars.json written")


def kernel_goldens():
    """Section 6: the unmodified reference engine stepping through its OTHER kernel classes
    (create_kernel: Wendland C2, Gaussian — kernels.cpp:166-236), on the inputs of two committed fixtures: the jittered
    cloud (per-particle masses, velocities) and the 13 k dam break (lattice + walls).  Run alone:
        python tests/golden/make_golden.py kernels"""
    if not po.available(KIND):
        raise SystemExit("oracle/_ref/liboracle_strict.so missing: make -f oracle/Makefile ref")
    for kt, tag in ((1, "wendland"), (2, "gaussian")):
        for src, steps, keep in (("cloud600", 3, {0, 2}), ("dam_break_13k_tame", 2, {1})):
            with np.load(HERE / f"{src}.npz") as z:
                prm = {k: np.float32(v) for k, v in zip(po.PARAM_NAMES, z["params"])}
                pos, vel, mass, dt = z["pos"], z["vel"], z["mass"], float(z["dts"][0])
            short = src.replace("_tame", "")
            save(f"{short}_{tag}", prm, pos, vel, mass, [dt] * steps, kernel_type=np.int32(kt),
                 **run_scene(prm, pos, vel, mass, [dt] * steps, keep_steps=keep, kernel_type=kt))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kernels":
        kernel_goldens()
    else:
        main()
        kernel_goldens()
