"""Host-side scene generators and parameter sets (numpy, fp32-exact mirrors of the reference's).

These are the Python counterparts of the lattice generators the reference keeps on the host
(``create_fluid_block`` / ``create_boundary_box``: reference src/particle.cpp:166-229;
``utils::create_dam_break_setup`` / ``create_fluid_drop_setup``: src/sph_engine.cpp:450-514).  Every
coordinate is produced by the same fp32 operations in the same order (``start + float(i) * spacing``),
so the arrays are bit-identical to what the C++ generators emit — tests/test_scenes.py checks that
against the compiled reference.  They are not on the per-step hot path.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _v3(a):
    return np.asarray(a, dtype=np.float32).reshape(3)


def create_fluid_block(center, size, spacing, mass=1.0):
    """reference src/particle.cpp:166-188 — returns (pos (n,3) f32, mass (n,) f32), i-major, k fastest."""
    c, s, dx = _v3(center), _v3(size), f32(spacing)
    n = [int(s[a] / dx) for a in range(3)]           # static_cast<int>(size.x / spacing), fp32 division
    start = c - s * f32(0.5)
    ax = [start[a] + np.arange(n[a], dtype=np.float32) * dx for a in range(3)]
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    return pos, np.full(pos.shape[0], mass, np.float32)


def create_boundary_box(center, size, spacing, mass=1.0):
    """reference src/particle.cpp:190-229 — six faces, int(size/spacing) points per axis, shared edges kept."""
    c, s, dx = _v3(center), _v3(size), f32(spacing)
    start = c - s * f32(0.5)
    faces = []

    def add_face(a1, a2, a3, value):
        n1, n2 = int(s[a1] / dx), int(s[a2] / dx)
        u = start[a1] + np.arange(n1, dtype=np.float32) * dx
        v = start[a2] + np.arange(n2, dtype=np.float32) * dx
        U, V = np.meshgrid(u, v, indexing="ij")
        p = np.empty((n1 * n2, 3), np.float32)
        p[:, a1] = U.ravel()
        p[:, a2] = V.ravel()
        p[:, a3] = f32(value)
        faces.append(p)

    add_face(0, 1, 2, start[2])
    add_face(0, 1, 2, start[2] + s[2])
    add_face(1, 2, 0, start[0])
    add_face(1, 2, 0, start[0] + s[0])
    add_face(0, 2, 1, start[1])
    add_face(0, 2, 1, start[1] + s[1])
    pos = np.concatenate(faces, axis=0)
    return pos, np.full(pos.shape[0], mass, np.float32)


def create_dam_break_setup(dam_size, fluid_size, spacing, mass):
    """reference src/sph_engine.cpp:450-487 — walls first, then the fluid block."""
    d, f = _v3(dam_size), _v3(fluid_size)
    bpos, bm = create_boundary_box((f32(0.0), d[1] / f32(2.0), f32(0.0)), d, spacing, mass)
    fpos, fm = create_fluid_block((-d[0] / f32(2.0) + f[0] / f32(2.0), f[1] / f32(2.0), f32(0.0)), f, spacing, mass)
    return np.concatenate([bpos, fpos]), np.concatenate([bm, fm])


def create_fluid_drop_setup(center, radius, spacing, mass):
    """reference src/sph_engine.cpp:489-514 — lattice points with dot(p-c, p-c) <= r*r."""
    c, r, dx = _v3(center), f32(radius), f32(spacing)
    n = int(f32(2.0) * r / dx)
    start = c - r
    ax = [start[a] + np.arange(n, dtype=np.float32) * dx for a in range(3)]
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    tx, ty, tz = X - c[0], Y - c[1], Z - c[2]
    keep = ((tx * tx + ty * ty) + tz * tz) <= r * r
    pos = np.stack([X[keep], Y[keep], Z[keep]], axis=1).astype(np.float32)
    return pos, np.full(pos.shape[0], mass, np.float32)


# ---- parameter sets -----------------------------------------------------------------

REFERENCE_DEFAULTS = dict(
    rest_density=1000.0, gas_constant=2000.0, viscosity=0.001, smoothing_length=0.02,
    particle_mass=0.001, timestep=0.001, gravity=-9.81, damping=0.99, CFL_factor=0.4,
    xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0, neighbor_search_radius=0.04,
)


def tame_params(dx: float, h: float, bounds) -> dict:
    """P-tame: bounded dynamics at lattice spacing dx (the reference's own defaults blow up within one
    step): m = 1.5 rho0 dx^3, k = c^2 m / rho0 with c^2 = 100, mu = 1e-3 m,
    dt = 0.25 h / c, damping 0.999; neighbor_search_radius = 2h as set_smoothing_length makes it."""
    rho0 = 1000.0
    m = 1.5 * rho0 * dx ** 3
    p = dict(REFERENCE_DEFAULTS)
    p.update(
        rest_density=rho0, gas_constant=100.0 * m / rho0, viscosity=1e-3 * m, smoothing_length=h,
        particle_mass=m, timestep=0.25 * h / 10.0, damping=0.999,
        neighbor_search_radius=float(f32(2.0) * f32(h)),   # SPHEngine::set_smoothing_length, sph_engine.cpp:160
        xmin=bounds[0], xmax=bounds[1], ymin=bounds[2], ymax=bounds[3], zmin=bounds[4], zmax=bounds[5],
    )
    return p


DAM_BOUNDS = (-0.2, 0.2, 0.0, 0.6, -0.4, 0.4)


def dam_break_scene(dx: float):
    """S2/S4/S5 family: the reference dam-break geometry (dam 0.4 x 0.6 x 0.8, fluid 0.2 x 0.4 x 0.8,
    sph_engine.cpp:40-45) at lattice spacing dx, h = 2 dx, P-tame parameters.
    Returns (pos, vel=None, mass, params, dt)."""
    params = tame_params(dx, 2.0 * dx, DAM_BOUNDS)
    pos, mass = create_dam_break_setup((0.4, 0.6, 0.8), (0.2, 0.4, 0.8), dx, params["particle_mass"])
    return pos, mass, params, params["timestep"]


def dam_break_count(dx: float) -> int:
    """Number of particles dam_break_scene(dx) yields (same fp32 int(size / spacing) arithmetic as the generators)."""
    dxf = f32(dx)
    n = lambda size: int(f32(size) / dxf)
    wx, wy, wz = n(0.4), n(0.6), n(0.8)
    return 2 * (wx * wy + wy * wz + wx * wz) + n(0.2) * n(0.4) * n(0.8)


def dam_break_dx_for(count: int, dx_guess: float) -> float:
    """Lattice spacing near dx_guess whose dam-break scene is closest to `count` particles (weak-scaling runs: N GPUs
    get N times the particles of the single-GPU scene, not just dx / N^(1/3), which falls ~6 % short at N = 8 because
    the wall layers grow with N^(2/3))."""
    cand = np.linspace(0.9 * dx_guess, 1.1 * dx_guess, 4001)
    err = [abs(dam_break_count(float(d)) - count) for d in cand]
    return float(cand[int(np.argmin(err))])


def fluid_drop_scene(dx: float):
    """S3 family: sphere of radius 0.1 centred (0, 0.5, 0) (sph_engine.cpp:56-61), h = 2.5 dx
    (the reference's 0.02 / 0.008 ratio), default [-1, 1]^3 bounds → sparse cell occupancy."""
    params = tame_params(dx, 2.5 * dx, (-1.0, 1.0, -1.0, 1.0, -1.0, 1.0))
    pos, mass = create_fluid_drop_setup((0.0, 0.5, 0.0), 0.1, dx, params["particle_mass"])
    return pos, mass, params, params["timestep"]


SCENES = {
    # name: (family, dx)  — sizes are what the generators yield
    "dam_break_13k": ("dam", 0.02),
    "dam_break_85k": ("dam", 0.01),
    "dam_break_150k": ("dam", 0.008),
    "dam_break_347k": ("dam", 0.006),
    "dam_break_1M": ("dam", 0.004),
    "dam_break_10M": ("dam", 0.00185),
    "dam_break_100M": ("dam", 0.00086),
    "fluid_drop_65k": ("drop", 0.004),
    "fluid_drop_1M": ("drop", 0.0016),
}


def make_scene(name: str):
    family, dx = SCENES[name]
    return dam_break_scene(dx) if family == "dam" else fluid_drop_scene(dx)
