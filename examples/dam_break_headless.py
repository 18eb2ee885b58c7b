#!/usr/bin/env python
"""Headless counterpart of the reference's examples/dam_break.cpp (lines 40-61 parameters, 114-175 main loop,
153-166 CSV row) on the B200 engine, through the drop-in Python module `sph`.

    python examples/dam_break_headless.py [num_particles=10000] [duration_s=0.05] [csv_path] [--strict]

Like the reference program, `num_particles` only sets the capacity (2 x num_particles); the scene itself is the
hard-wired dam break of SPHEngine::initialize_dam_break, truncated at capacity.  The CSV has the
reference's columns: time,particles,mass_error,energy,total_energy,avg_density,max_velocity — avg_density is the
mean over the capacity-length density buffer, as in dam_break.cpp:146-151 (quirk Q13).  The report values come from
ONE device reduction per report (Simulator.get_report_diagnostics); the reference program pulls the full velocity and
density arrays for them.  --strict selects the bit-exact reference arithmetic (default: the fast path).
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "sph-particle-simulator_b200" / "python"))
import sph  # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    num_particles = int(argv[0]) if len(argv) > 0 else 10000
    duration = float(argv[1]) if len(argv) > 1 else 0.05
    csv_path = argv[2] if len(argv) > 2 else "dam_break_data.csv"

    engine = sph.Simulator(max_particles=num_particles * 2)
    if "--strict" in sys.argv:
        engine.set_math_mode(0)
    params = sph.SPHParameters()
    params.rest_density = 1000.0
    params.gas_constant = 2000.0
    params.viscosity = 0.001
    params.smoothing_length = 0.025
    params.particle_mass = 0.001
    params.timestep = 0.001
    params.gravity = -9.81
    params.damping = 0.995
    engine.initialize(params)
    engine.set_boundaries(-1.0, 1.0, -0.5, 1.5, -1.0, 1.0)
    engine.initialize_dam_break()
    n = engine.get_particles().size()
    print(f"Simulation initialized with {n} particles")

    dt = params.timestep
    max_steps = int(duration / dt)
    rows = ["time,particles,mass_error,energy,total_energy,avg_density,max_velocity"]
    t0 = time.perf_counter()
    for step in range(max_steps):
        engine.step(dt)
        if step % 10 == 9 or step == max_steps - 1:
            d = engine.get_report_diagnostics()    # one device reduction; energy_error is a constant 0 in the reference too
            rows.append(f"{engine.get_current_time():.6f},{n},{d['mass_error']:.6g},0,{d['kinetic_energy']:.6g},"
                        f"{d['average_density']:.6g},{d['max_velocity']:.6g}")
    wall = time.perf_counter() - t0
    Path(csv_path).write_text("\n".join(rows) + "\n")
    st = engine.get_performance_stats()
    print(f"{max_steps} steps in {wall:.3f} s wall ({max_steps * n / max(wall, 1e-9) / 1e6:.1f} M particle-updates/s incl. diagnostics)")
    print(f"GPU stage seconds: neighbour {st.neighbor_search_time:.4f} density {st.density_computation_time:.4f} "
          f"force {st.force_computation_time:.4f} integration {st.integration_time:.4f}; max neighbours {st.max_neighbors}")
    print(f"wrote {csv_path} ({len(rows) - 1} rows)")


if __name__ == "__main__":
    main()
